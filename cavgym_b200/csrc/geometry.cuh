// geometry.cuh — oriented boxes, stopping zones, separating-axis test, containment and
// convex clipping on device.  Replaces the Shapely bridge of the reference
// (library/geometry.py:74-87 Shape.intersects / contains / percentage_intersects),
// make_rectangle + transform (geometry.py:241-251,117-126) and DynamicBody.stopping_zones
// (library/bodies.py:122-135).
//
// All quads are clockwise (rear_left, front_left, front_right, rear_right with x forward,
// y left), so a point is OUTSIDE edge i->j when cross(e, p - a) > 0.  Predicates are closed:
// touching counts as intersecting (Shapely semantics).  Every predicate also reports whether
// its decision margin is within `tau` pixels of zero (near-tangent flag, north star).
#pragma once
#include "dev_types.cuh"

namespace cav {

// a / b to ~2 ulp with a float reciprocal seed and two Newton steps (10 instructions instead of the ~60 of the
// IEEE division subroutine).  Used only for quantities that feed rewards and flagged predicates, never body state;
// |b| must be a normal float (here: pixel distances).
__device__ __forceinline__ double fast_div(double a, double b) {
  double r = (double)__frcp_rn((float)b);
  r = r * (2.0 - b * r);
  r = r * (2.0 - b * r);
  return a * r;
}
__device__ __forceinline__ float fast_div(float a, float b) { return a / b; }

template <typename R> __device__ __forceinline__ R rmin(R a, R b) { return b < a ? b : a; }
template <typename R> __device__ __forceinline__ R rmax(R a, R b) { return b > a ? b : a; }
__device__ __forceinline__ double rsqrt_(double v) { return sqrt(v); }
__device__ __forceinline__ float rsqrt_(float v) { return sqrtf(v); }
__device__ __forceinline__ double rabs(double v) { return fabs(v); }
__device__ __forceinline__ float rabs(float v) { return fabsf(v); }

template <typename R>
__device__ __forceinline__ Aabb<R> aabb_of(const Quad<R>& q) {
  Aabb<R> b;
  b.x0 = rmin(rmin(q.x[0], q.x[1]), rmin(q.x[2], q.x[3]));
  b.x1 = rmax(rmax(q.x[0], q.x[1]), rmax(q.x[2], q.x[3]));
  b.y0 = rmin(rmin(q.y[0], q.y[1]), rmin(q.y[2], q.y[3]));
  b.y1 = rmax(rmax(q.y[0], q.y[1]), rmax(q.y[2], q.y[3]));
  return b;
}

// Largest axis-aligned gap between two boxes (> 0 means disjoint along x or y).
template <typename R>
__device__ __forceinline__ R aabb_gap(const Aabb<R>& a, const Aabb<R>& b) {
  return rmax(rmax(a.x0 - b.x1, b.x0 - a.x1), rmax(a.y0 - b.y1, b.y0 - a.y1));
}

// aabb_gap(a, b) > tau without the max tree: four subtractions and compares.  True means the polygons inside the
// boxes are certainly disjoint, with a safety margin of tau >> rounding error (so no near-tangent flag is needed).
template <typename R>
__device__ __forceinline__ bool aabb_apart(const Aabb<R>& a, const Aabb<R>& b, R tau) {
  return (a.x0 - b.x1 > tau) || (b.x0 - a.x1 > tau) || (a.y0 - b.y1 > tau) || (b.y0 - a.y1 > tau);
}

// AABB of make_rectangle(length, width).transform(theta, p) from the half extents |c|hl + |s|hw, |s|hl + |c|hw
// (no corner min/max tree).  For theta == 0 it equals the corner AABB bit for bit; otherwise within a few ulp,
// far inside the tau margin every consumer applies.
template <typename R>
__device__ __forceinline__ Aabb<R> box_aabb(R length, R width, R theta, R c, R s, R px, R py) {
  const R hl = length * R(0.5), hw = width * R(0.5);
  R ex = hl, ey = hw;
  if (!(theta == R(0))) {
    const R ac = rabs(c), as = rabs(s);
    ex = ac * hl + as * hw;
    ey = as * hl + ac * hw;
  }
  return {px - ex, px + ex, py - ey, py + ey};
}

// make_rectangle(length, width) . transform(theta, (px, py)); c, s = cos/sin(theta).
// theta == 0 takes the reference's translate-only path (geometry.py:118-119) bit for bit.
template <typename R>
__device__ __forceinline__ void make_box(R length, R width, R theta, R c, R s, R px, R py, Quad<R>& q) {
  const R hl = length * R(0.5), hw = width * R(0.5);
  const R lx[4] = {-hl, hl, hl, -hl};
  const R ly[4] = {hw, hw, -hw, -hw};
  if (theta == R(0)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { q.x[i] = px + lx[i]; q.y[i] = py + ly[i]; }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      q.x[i] = px + ((c * lx[i]) - (s * ly[i]));
      q.y[i] = py + ((s * lx[i]) + (c * ly[i]));
    }
  }
}

// DynamicBody.stopping_zones + split_longitudinally (bodies.py:122-135, geometry.py:176-191).
// The braking/reaction split is built directly (rectangles of length bd and rd laid end to end from the
// front-centre anchor) instead of by interpolating with p = bd / td: algebraically identical, no division on the
// always-executed path, corners within a few ulp of the reference's.  The oracle keeps the reference's form.
template <typename R>
struct ZoneFrame {
  bool have;
  R ax, ay, bd, td, hw;  // anchor (front centre), braking and total distance, half width
};

template <typename R>
__device__ __forceinline__ ZoneFrame<R> zone_frame(const DevType<R>& k, R x, R y, R v, R theta, R c, R s, R steer) {
  ZoneFrame<R> z;
  z.bd = (v * v) * k.inv_2brake;
  z.td = z.bd + v * R(0.675);
  z.have = !(z.td == R(0)) && (steer == R(0));
  const R hx = k.length * R(0.5);
  z.hw = k.width * R(0.5);
  if (theta == R(0)) { z.ax = x + hx; z.ay = y; }
  else { z.ax = x + c * hx; z.ay = y + s * hx; }
  return z;
}

// Rectangle spanning local x in [x0, x1], y in [-hw, +hw] of the zone frame, corners RL, FL, FR, RR.
template <typename R>
__device__ __forceinline__ void zone_quad(const ZoneFrame<R>& z, R theta, R c, R s, R x0, R x1, Quad<R>& q) {
  const R lx[4] = {x0, x1, x1, x0};
  const R ly[4] = {z.hw, z.hw, -z.hw, -z.hw};
  if (theta == R(0)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { q.x[i] = z.ax + lx[i]; q.y[i] = z.ay + ly[i]; }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      q.x[i] = z.ax + ((c * lx[i]) - (s * ly[i]));
      q.y[i] = z.ay + ((s * lx[i]) + (c * ly[i]));
    }
  }
}

template <typename R>
__device__ __forceinline__ Aabb<R> zone_aabb(const ZoneFrame<R>& z, R theta, R c, R s, R x0, R x1) {
  if (theta == R(0)) return {z.ax + x0, z.ax + x1, z.ay - z.hw, z.ay + z.hw};
  Quad<R> q;
  zone_quad(z, theta, c, s, x0, x1, q);
  return aabb_of(q);
}

// Separating-axis bookkeeping without square roots or divisions.  For edge i of A let m_i be the smallest
// cross(e_i, p - a_i) over the vertices p of B (> 0 means all of B strictly outside that edge line) and
// near_i <=> |m_i| < tau * |e_i|  <=>  m_i^2 < tau^2 |e_i|^2.
//   bit 0 (SEP_CLEAR)   some edge separates with margin >= tau
//   bit 1 (SEP_ANY)     some edge separates (m_i > 0)
//   bit 2 (NEAR_ANY)    some edge has |margin| < tau
enum { SEP_CLEAR = 1, SEP_ANY = 2, NEAR_ANY = 4 };

template <typename R>
__device__ __forceinline__ int separation_bits(const Quad<R>& A, const Quad<R>& B, R tau2) {
  int bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    const R ax = A.x[i], ay = A.y[i], ex = A.x[j] - ax, ey = A.y[j] - ay;
    const R len2 = ex * ex + ey * ey;
    R m = INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) m = rmin(m, ex * (B.y[k] - ay) - ey * (B.x[k] - ax));
    if (len2 > R(0)) {
      const bool near = m * m < tau2 * len2;
      if (m > R(0)) bits |= near ? SEP_ANY : (SEP_ANY | SEP_CLEAR);
      if (near) bits |= NEAR_ANY;
    }
  }
  return bits;
}

// Result of a rare-path predicate: bit 0 = predicate holds, bit 1 = decision margin within tau (near-tangent).
enum { GEO_HIT = 1, GEO_TANGENT = 2 };

// Full separating-axis test (8 edge normals).  Rare path, called only when the AABBs overlap.  Quads are passed
// BY VALUE so that the caller's copies are never address-taken and stay in registers on the common path.
// Near-tangent <=> the largest normalised margin lies in (-tau, tau) <=> no edge separates clearly and
// some edge is within tau of touching.
template <typename R>
__device__ __noinline__ int sat_intersects(Quad<R> A, Quad<R> B, R tau) {
  const int bits = separation_bits(A, B, tau * tau) | separation_bits(B, A, tau * tau);
  const bool hit = !(bits & SEP_ANY);
  const bool tangent = !(bits & SEP_CLEAR) && (bits & NEAR_ANY);
  return (hit ? GEO_HIT : 0) | (tangent ? GEO_TANGENT : 0);
}

// Shape.intersects for two convex quads.  AABB rejection first: exact and conservative.
template <typename R>
__device__ __forceinline__ bool intersects(const Quad<R>& A, const Aabb<R>& a, const Quad<R>& B, const Aabb<R>& b, R tau,
                                           bool& tangent) {
  if (aabb_gap(a, b) > tau) return false;
  const int r = sat_intersects(A, B, tau);
  if (r & GEO_TANGENT) tangent = true;
  return (r & GEO_HIT) != 0;
}

// outer.contains(inner): every vertex of inner inside or on every edge of outer.  Near-tangent when the
// worst vertex is within tau of some edge line (compared without sqrt: m^2 < tau^2 |e|^2).
template <typename R>
__device__ __forceinline__ bool contains(const Quad<R>& outer, const Quad<R>& inner, R tau, bool& tangent) {
  bool inside = true, clear_out = false, near = false;
  const R tau2 = tau * tau;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    const R ax = outer.x[i], ay = outer.y[i], ex = outer.x[j] - ax, ey = outer.y[j] - ay;
    const R len2 = ex * ex + ey * ey;
    R m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) m = rmax(m, ex * (inner.y[k] - ay) - ey * (inner.x[k] - ax));
    if (len2 > R(0)) {
      const bool close = m * m < tau2 * len2;
      if (m > R(0)) { inside = false; if (!close) clear_out = true; }
      if (close) near = true;
    }
  }
  if (near && !clear_out) tangent = true;
  return inside;
}

template <typename R>
__device__ __forceinline__ R quad_area(const Quad<R>& q) {
  R a = R(0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    a += q.x[i] * q.y[j] - q.x[j] * q.y[i];
  }
  return rabs(a) * R(0.5);
}

// area(subject ∩ clip) by Sutherland–Hodgman; stands for Shapely's intersection(...).area.
// Rare path (a box straddling a road edge): vertex lists live in local memory.
template <typename R>
__device__ __noinline__ R clip_area(const Quad<R>& subject, const Quad<R>& clip) {
  R sx[8], sy[8], ox[8], oy[8];
  int n = 4;
  for (int i = 0; i < 4; ++i) { sx[i] = subject.x[i]; sy[i] = subject.y[i]; }
  for (int i = 0; i < 4 && n > 0; ++i) {
    const int j = (i + 1) & 3;
    const R ax = clip.x[i], ay = clip.y[i], ex = clip.x[j] - ax, ey = clip.y[j] - ay;
    if (ex == R(0) && ey == R(0)) continue;
    int m = 0;
    for (int k = 0; k < n; ++k) {
      const int l = (k + 1 == n) ? 0 : k + 1;
      // clockwise clip ring: inside is cross <= 0, so negate to keep "inside >= 0"
      const R sp = -(ex * (sy[k] - ay) - ey * (sx[k] - ax));
      const R sq = -(ex * (sy[l] - ay) - ey * (sx[l] - ax));
      if (sp >= R(0)) { ox[m] = sx[k]; oy[m] = sy[k]; ++m; }
      if ((sp > R(0) && sq < R(0)) || (sp < R(0) && sq > R(0))) {
        const R t = sp / (sp - sq);
        ox[m] = sx[k] + t * (sx[l] - sx[k]);
        oy[m] = sy[k] + t * (sy[l] - sy[k]);
        ++m;
      }
    }
    n = m;
    for (int k = 0; k < n; ++k) { sx[k] = ox[k]; sy[k] = oy[k]; }
  }
  if (n < 3) return R(0);
  R a = R(0);
  for (int k = 0; k < n; ++k) {
    const int l = (k + 1 == n) ? 0 : k + 1;
    a += sx[k] * sy[l] - sx[l] * sy[k];
  }
  return rabs(a) * R(0.5);
}

// Shape.percentage_intersects (geometry.py:80-87): share of `self` lying on `other`.
// Callers reject AABB-disjoint pairs first, so this is the rare path (by-value arguments, see sat_intersects).
template <typename R>
struct Share {
  R value;
  int tangent;
};

template <typename R>
__device__ __noinline__ Share<R> percentage_intersects(Quad<R> self, Quad<R> other, R tau) {
  Share<R> out;
  bool tangent = false;
  const int bits = separation_bits(self, other, tau * tau) | separation_bits(other, self, tau * tau);
  if (!(bits & SEP_CLEAR) && (bits & NEAR_ANY)) tangent = true;
  if (bits & SEP_ANY) out.value = R(0);
  else if (contains(other, self, tau, tangent)) out.value = R(1);
  else out.value = clip_area(self, other) / quad_area(self);
  out.tangent = tangent ? 1 : 0;
  return out;
}

// Area of the part of a convex quad inside ONE half-plane {p : nx*x + ny*y <= bound}, by a single
// Sutherland–Hodgman stage with a running shoelace sum (no vertex list, registers only).  This is the
// kerb-crossing case: a body box straddling exactly one edge of an axis-aligned road rectangle, where (nx, ny) is
// (+-1, 0) or (0, +-1) and the products are exact.
template <typename R>
__device__ __forceinline__ R halfplane_area(const Quad<R>& q, R nx, R ny, R bound) {
  R d[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i] = bound - (nx * q.x[i] + ny * q.y[i]);  // >= 0 inside
  R acc = R(0), fx = R(0), fy = R(0), px = R(0), py = R(0);
  bool have = false;
  auto emit = [&](R x, R y) {
    if (have) acc += px * y - x * py;
    else { fx = x; fy = y; have = true; }
    px = x; py = y;
  };
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    if (d[i] >= R(0)) emit(q.x[i], q.y[i]);
    if ((d[i] > R(0) && d[j] < R(0)) || (d[i] < R(0) && d[j] > R(0))) {
      const R t = fast_div(d[i], d[i] - d[j]);
      emit(q.x[i] + t * (q.x[j] - q.x[i]), q.y[i] + t * (q.y[j] - q.y[i]));
    }
  }
  if (have) acc += px * fy - fx * py;
  return rabs(acc) * R(0.5);
}

// How a body-box AABB sits on an EXACTLY axis-aligned road rectangle (road == road_bb); callers have already
// rejected boxes that are clearly apart.  Margins m >= tau mean "clearly inside that edge".
//   AXIS_INSIDE    clearly inside all four edges                       -> share 1
//   AXIS_ONE_EDGE  clearly across exactly one edge, inside the others  -> single half-plane clip (the kerb)
//   AXIS_GENERAL   road corner or within tau of an edge                -> general predicates
enum { AXIS_INSIDE = 0, AXIS_ONE_EDGE = 1, AXIS_GENERAL = 2 };

template <typename R>
__device__ __forceinline__ int axis_case(const Aabb<R>& bb, const Aabb<R>& road, R tau) {
  const R m0 = bb.x0 - road.x0, m1 = road.x1 - bb.x1, m2 = bb.y0 - road.y0, m3 = road.y1 - bb.y1;
  const int out = (m0 < tau) + (m1 < tau) + (m2 < tau) + (m3 < tau);
  if (out == 0) return AXIS_INSIDE;
  const int across = (m0 <= -tau) + (m1 <= -tau) + (m2 <= -tau) + (m3 <= -tau);
  return (out == 1 && across == 1) ? AXIS_ONE_EDGE : AXIS_GENERAL;
}

// ---------------------------------------------------------------- out-of-line rare paths
// Everything below is __noinline__ and takes small by-value arguments: the always-executed path of a step stays a
// few KB of straight-line code (it must fit the instruction caches), and the rare geometry lives once per kernel.

template <typename R>
struct Pose {  // what is needed to rebuild a body's corners: make_rectangle(length, width).transform(theta, (x, y))
  R x, y, theta, c, s, length, width;
};

template <typename R>
__device__ __forceinline__ void pose_quad(const Pose<R>& p, Quad<R>& q) {
  make_box(p.length, p.width, p.theta, p.c, p.s, p.x, p.y, q);
}

// body box vs body box
template <typename R>
__device__ __noinline__ int sat_pose_pose(Pose<R> a, Pose<R> b, R tau) {
  Quad<R> qa, qb;
  pose_quad(a, qa);
  pose_quad(b, qb);
  const int bits = separation_bits(qa, qb, tau * tau) | separation_bits(qb, qa, tau * tau);
  return (!(bits & SEP_ANY) ? GEO_HIT : 0) | ((!(bits & SEP_CLEAR) && (bits & NEAR_ANY)) ? GEO_TANGENT : 0);
}

// body box vs a static quad of the scenario tables (road, traffic light, obstacle, crossing box)
template <typename R>
__device__ __noinline__ int sat_pose_quad(Pose<R> a, const Quad<R>* other, R tau) {
  Quad<R> qa, qb = *other;
  pose_quad(a, qa);
  const int bits = separation_bits(qa, qb, tau * tau) | separation_bits(qb, qa, tau * tau);
  return (!(bits & SEP_ANY) ? GEO_HIT : 0) | ((!(bits & SEP_CLEAR) && (bits & NEAR_ANY)) ? GEO_TANGENT : 0);
}

// body box vs the part [x0, x1] of the ego's stopping-zone frame (braking or reaction zone)
template <typename R>
__device__ __noinline__ int sat_pose_zone(Pose<R> a, ZoneFrame<R> z, R theta, R c, R s, R x0, R x1, R tau) {
  Quad<R> qa, qz;
  pose_quad(a, qa);
  zone_quad(z, theta, c, s, x0, x1, qz);
  const int bits = separation_bits(qa, qz, tau * tau) | separation_bits(qz, qa, tau * tau);
  return (!(bits & SEP_ANY) ? GEO_HIT : 0) | ((!(bits & SEP_CLEAR) && (bits & NEAR_ANY)) ? GEO_TANGENT : 0);
}

// The kerb crossing: share of a body box lying on an axis-aligned road when its AABB is clearly across exactly
// one road edge (axis_case == AXIS_ONE_EDGE).  One compact out-of-line instance.
template <typename R>
__device__ __noinline__ R kerb_share(Pose<R> a, Aabb<R> bb, Aabb<R> road, R tau) {
  Quad<R> box;
  pose_quad(a, box);
  R nx = R(0), ny = R(0), bound;
  if (bb.x0 - road.x0 < tau) { nx = R(-1); bound = -road.x0; }        // inside: x >= x0
  else if (road.x1 - bb.x1 < tau) { nx = R(1); bound = road.x1; }     // inside: x <= x1
  else if (bb.y0 - road.y0 < tau) { ny = R(-1); bound = -road.y0; }
  else { ny = R(1); bound = road.y1; }
  return fast_div(halfplane_area(box, nx, ny, bound), a.length * a.width);
}

// Share of a body box lying on a road by the general predicates (Shape.percentage_intersects, geometry.py:80-87):
// rotated roads, road corners, near-tangent configurations.  Cold.
template <typename R>
__device__ __noinline__ Share<R> road_share_general(Pose<R> a, const Quad<R>* road, R tau) {
  Quad<R> box;
  pose_quad(a, box);
  const Quad<R> other = *road;
  Share<R> out;
  bool tangent = false;
  const int bits = separation_bits(box, other, tau * tau) | separation_bits(other, box, tau * tau);
  if (!(bits & SEP_CLEAR) && (bits & NEAR_ANY)) tangent = true;
  if (bits & SEP_ANY) out.value = R(0);
  else if (contains(other, box, tau, tangent)) out.value = R(1);
  else out.value = clip_area(box, other) / quad_area(box);
  out.tangent = tangent ? 1 : 0;
  return out;
}

}  // namespace cav
