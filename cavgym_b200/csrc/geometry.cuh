// geometry.cuh — oriented boxes, stopping zones, separating-axis tests, containment and
// convex clipping.  Replaces the Shapely bridge of the reference
// (library/geometry.py:74-87 Shape.intersects / contains / percentage_intersects),
// make_rectangle + transform (geometry.py:241-251,117-126) and DynamicBody.stopping_zones
// (library/bodies.py:122-135).
//
// Two representations:
//  * Box<R> (centre, unit heading, half extents): every rectangle-vs-rectangle test of a step.
//    The separating-axis test is four |T.L| - (ra + rb) margins, the share of a box lying on an
//    axis-aligned road across one kerb is the closed-form area of a rectangle cut by a half-plane.
//    Branch-free, registers only: this is the always-executed path.
//  * Quad<R> (corner lists, clockwise): general convex quads — non-rectangular statics, rotated
//    roads, road corners, the stand-alone geometry probe.  Out of line, rare.
// Predicates are closed: touching counts as intersecting (Shapely semantics).  Every predicate
// also reports whether its decision margin is within `tau` pixels of zero (near-tangent flag).
#pragma once
#include <cmath>

#include "dev_types.cuh"

#define CAV_HD __host__ __device__ __forceinline__
// Rare branches of the hot loops: laid out away from the straight-line path.  The long-transition kernels are bound by
// instruction fetch; measured with the hints on the branches below: 20-step replay launch 66.3 -> 64.7 us,
// two-body rollout 6.65 -> 6.48 ms per 100 steps of 1M envs, per-step kernel at 4M envs 0.158 -> 0.160 ms.  (More hints —
// near-tangent tests, road corners, crossing-agent transitions — gave the fused kernels nothing and cost the per-step
// kernel 3 %.)
#define CAV_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define CAV_LIKELY(x) __builtin_expect(!!(x), 1)

namespace cav {

// a / b to ~2 ulp with a float reciprocal seed and two Newton steps (10 instructions instead of the ~60 of the
// IEEE division subroutine).  Used only where |b| is a normal float (pixel distances, sines of steering angles).
__device__ __forceinline__ double fast_div(double a, double b) {
  double r = (double)__frcp_rn((float)b);
  r = fma(r, fma(-b, r, 1.0), r);
  r = fma(r, fma(-b, r, 1.0), r);
  const double q = a * r;
  return fma(r, fma(-b, q, a), q);  // one residual correction of the quotient: <= 1 ulp
}
__device__ __forceinline__ float fast_div(float a, float b) { return a / b; }

template <typename R> CAV_HD R rmin(R a, R b) { return b < a ? b : a; }
template <typename R> CAV_HD R rmax(R a, R b) { return b > a ? b : a; }
CAV_HD double rsqrt_(double v) { return sqrt(v); }
CAV_HD float rsqrt_(float v) { return sqrtf(v); }
CAV_HD double rabs(double v) { return fabs(v); }
CAV_HD float rabs(float v) { return fabsf(v); }

// ================================================================ centre-extent boxes (hot path)

// Half extents of the axis-aligned bounding box of a rotated rectangle: |c|hl + |s|hw, |s|hl + |c|hw.
template <typename R>
__device__ __forceinline__ void box_extents(R c, R s, R hl, R hw, R& ex, R& ey) {
  const R ac = rabs(c), as = rabs(s);
  ex = ac * hl + as * hw;
  ey = as * hl + ac * hw;
}

// Largest separating-axis margin of two oriented rectangles over their four edge normals:
//   margin_L = |T . L| - (radius_a(L) + radius_b(L)),   T = centre_b - centre_a.
// > 0: some axis separates them (disjoint).  <= 0: they intersect (closed).  |margin| < tau: near-tangent.
template <typename R>
__device__ __forceinline__ R box_margin(const Box<R>& a, const Box<R>& b) {
  const R tx = b.px - a.px, ty = b.py - a.py;
  const R cd = a.c * b.c + a.s * b.s, sd = a.c * b.s - a.s * b.c;  // cos / sin of the relative heading
  const R acd = rabs(cd), asd = rabs(sd);
  const R m1 = rabs(tx * a.c + ty * a.s) - (a.hl + (acd * b.hl + asd * b.hw));
  const R m2 = rabs(ty * a.c - tx * a.s) - (a.hw + (asd * b.hl + acd * b.hw));
  const R m3 = rabs(tx * b.c + ty * b.s) - (b.hl + (acd * a.hl + asd * a.hw));
  const R m4 = rabs(ty * b.c - tx * b.s) - (b.hw + (asd * a.hl + acd * a.hw));
  return rmax(rmax(m1, m2), rmax(m3, m4));
}

// A pedestrian box against the ego's box and its two stopping zones in ONE pass.  The zones are rectangles in the
// ego's frame laid end to end from the front-centre anchor (bodies.py:122-135, geometry.py:176-191): braking
// [hl0, hl0 + bd], reaction [hl0 + bd, hl0 + td] along the heading, the ego's width across.  All three tests share the
// relative pose; each costs three more margins.  (The reference interpolates the split corners with p = bd / td —
// identical up to a few ulp, far inside tau.)
template <typename R>
struct EgoFrame {
  R x, y, c, s, hl, hw;  // ego pose
  R bd, td;              // braking and total stopping distance
  bool have;             // zones exist: td != 0 and the ego is not steering (bodies.py:130-135)
};

template <typename R>
struct EgoMargins {
  R ego, braking, reaction;  // largest separating-axis margin against each rectangle
  bool all_clear;            // every rectangle is separated from the pedestrian by tau or more (the common case)
};

template <typename R>
__device__ __forceinline__ EgoMargins<R> ego_margins(const EgoFrame<R>& f, R px, R py, R cb, R sb, R hlb, R hwb, R tau) {
  const R tx = px - f.x, ty = py - f.y;
  const R cd = f.c * cb + f.s * sb, sd = f.c * sb - f.s * cb;
  const R acd = rabs(cd), asd = rabs(sd);
  const R u = tx * f.c + ty * f.s, w = ty * f.c - tx * f.s;    // pedestrian centre in the ego frame
  const R tp = tx * cb + ty * sb, tq = ty * cb - tx * sb;      // ego centre offset in the pedestrian frame
  const R exb = acd * hlb + asd * hwb, eyb = asd * hlb + acd * hwb;
  const R lateral = rabs(w) - (f.hw + eyb);                    // same for all three rectangles
  const R rp = hlb + asd * f.hw, rq = hwb + acd * f.hw;        // pedestrian-axis radii without the length term
  EgoMargins<R> out;
  // The strip [-hl, hl + td] x [-hw, hw] contains all three rectangles: being clear of it laterally, or beyond either
  // end along the heading, settles all three tests at once (f.td = 0 leaves the ego box alone).
  const R reach = f.have ? f.td : R(0);
  const R along = rmax(-f.hl - u, u - (f.hl + reach)) - exb;
  out.all_clear = (lateral >= tau) || (along >= tau);
  out.ego = out.braking = out.reaction = R(1);
  if (!out.all_clear) {
    auto margin = [&](R centre, R half) {
      const R mx = rabs(u - centre) - (half + exb);
      const R m3 = rabs(tp - centre * cd) - (rp + acd * half);
      const R m4 = rabs(tq + centre * sd) - (rq + asd * half);
      return rmax(rmax(lateral, mx), rmax(m3, m4));
    };
    out.ego = margin(R(0), f.hl);
    const R hb = f.bd * R(0.5), hr = (f.td - f.bd) * R(0.5);
    out.braking = margin(f.hl + hb, hb);
    out.reaction = margin(f.hl + f.bd + hr, hr);
  }
  return out;
}

// Fraction of a rectangle on the inner side of a line, as a function of the signed distance t of its centre from
// the line (t > 0 inside) and the projections u, w >= 0 of its two half-edge vectors on the line's normal.  The
// signed distance of a uniformly distributed point of the rectangle is t + U(-u, u) + U(-w, w), so the fraction is the
// CDF of a trapezoidal distribution: quadratic while only one corner has crossed, linear in between.
// This is Shape.percentage_intersects (geometry.py:80-87) for a body box across ONE edge of an axis-aligned road.
template <typename R>
__device__ __forceinline__ R kerb_share(R t, R u, R w) {
  const R A = rmax(u, w), B = rmin(u, w), D = A - B, S = A + B;
  const bool corner = (rabs(t) > D) && (B > R(0));
  const R q = rmax(R(0), S - rabs(t));
  const R num = corner ? q * q : rmax(R(0), rmin(A + A, t + A));
  const R den = corner ? R(8) * (A * B) : A + A;
  const R f = fast_div(num, den);
  return (corner && t > R(0)) ? R(1) - f : f;
}

// ================================================================ general convex quads (rare paths)

template <typename R>
CAV_HD Aabb<R> aabb_of(const Quad<R>& q) {
  Aabb<R> b;
  b.x0 = rmin(rmin(q.x[0], q.x[1]), rmin(q.x[2], q.x[3]));
  b.x1 = rmax(rmax(q.x[0], q.x[1]), rmax(q.x[2], q.x[3]));
  b.y0 = rmin(rmin(q.y[0], q.y[1]), rmin(q.y[2], q.y[3]));
  b.y1 = rmax(rmax(q.y[0], q.y[1]), rmax(q.y[2], q.y[3]));
  return b;
}

// Largest axis-aligned gap between two boxes (> 0 means disjoint along x or y).
template <typename R>
CAV_HD R aabb_gap(const Aabb<R>& a, const Aabb<R>& b) {
  return rmax(rmax(a.x0 - b.x1, b.x0 - a.x1), rmax(a.y0 - b.y1, b.y0 - a.y1));
}

// make_rectangle(length, width) . transform(theta, (px, py)); c, s = cos/sin(theta).
// theta == 0 takes the reference's translate-only path (geometry.py:118-119) bit for bit.
template <typename R>
CAV_HD void make_box(R length, R width, R theta, R c, R s, R px, R py, Quad<R>& q) {
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

// Separating-axis bookkeeping without square roots or divisions.  For edge i of A let m_i be the smallest
// cross(e_i, p - a_i) over the vertices p of B (> 0 means all of B strictly outside that edge line) and
// near_i <=> |m_i| < tau * |e_i|  <=>  m_i^2 < tau^2 |e_i|^2.
//   bit 0 (SEP_CLEAR)   some edge separates with margin >= tau
//   bit 1 (SEP_ANY)     some edge separates (m_i > 0)
//   bit 2 (NEAR_ANY)    some edge has |margin| < tau
enum { SEP_CLEAR = 1, SEP_ANY = 2, NEAR_ANY = 4 };

template <typename R>
CAV_HD int separation_bits(const Quad<R>& A, const Quad<R>& B, R tau2) {
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

// Result of a predicate: bit 0 = predicate holds, bit 1 = decision margin within tau (near-tangent).
enum { GEO_HIT = 1, GEO_TANGENT = 2 };

// Full separating-axis test on corner lists (8 edge normals).  Near-tangent <=> the largest normalised margin lies
// in (-tau, tau) <=> no edge separates clearly and some edge is within tau of touching.
template <typename R>
CAV_HD int sat_bits(const Quad<R>& A, const Quad<R>& B, R tau) {
  const int bits = separation_bits(A, B, tau * tau) | separation_bits(B, A, tau * tau);
  const bool hit = !(bits & SEP_ANY);
  const bool tangent = !(bits & SEP_CLEAR) && (bits & NEAR_ANY);
  return (hit ? GEO_HIT : 0) | (tangent ? GEO_TANGENT : 0);
}

// Shape.intersects for two convex quads.  AABB rejection first: exact and conservative.
template <typename R>
CAV_HD bool intersects(const Quad<R>& A, const Aabb<R>& a, const Quad<R>& B, const Aabb<R>& b, R tau, bool& tangent) {
  if (aabb_gap(a, b) > tau) return false;
  const int r = sat_bits(A, B, tau);
  if (r & GEO_TANGENT) tangent = true;
  return (r & GEO_HIT) != 0;
}

// outer.contains(inner): every vertex of inner inside or on every edge of outer.  Near-tangent when the
// worst vertex is within tau of some edge line (compared without sqrt: m^2 < tau^2 |e|^2).
template <typename R>
CAV_HD bool contains(const Quad<R>& outer, const Quad<R>& inner, R tau, bool& tangent) {
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
CAV_HD R quad_area(const Quad<R>& q) {
  R a = R(0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    a += q.x[i] * q.y[j] - q.x[j] * q.y[i];
  }
  return rabs(a) * R(0.5);
}

// area(subject ∩ clip) by Sutherland–Hodgman; stands for Shapely's intersection(...).area.
// Vertex lists live in local memory: rare path only.
template <typename R>
__host__ __device__ inline R clip_area(const Quad<R>& subject, const Quad<R>& clip) {
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
template <typename R>
struct Share {
  R value;
  int tangent;
};

template <typename R>
__host__ __device__ inline Share<R> percentage_of(const Quad<R>& self, const Quad<R>& other, R tau) {
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

// ---------------------------------------------------------------- out-of-line device entry points for the rare paths
// __noinline__ with small by-value arguments: the always-executed path of a step stays compact straight-line code and
// the general-polygon machinery exists once per kernel.

template <typename R>
struct Pose {  // what is needed to rebuild a body's corners: make_rectangle(length, width).transform(theta, (x, y))
  R x, y, theta, c, s, length, width;
};

// body box vs a quad of the scenario tables that is not a rectangle
template <typename R>
__device__ __noinline__ int sat_pose_quad(Pose<R> a, const Quad<R>* other, R tau) {
  Quad<R> qa;
  const Quad<R> qb = *other;
  make_box(a.length, a.width, a.theta, a.c, a.s, a.x, a.y, qa);
  return sat_bits(qa, qb, tau);
}

// Share of a body box lying on an axis-aligned road across one of its CORNERS: the box crosses one x-edge and one
// y-edge and is clear of the other two, so box ∩ road = box ∩ {sx x <= bx} ∩ {sy y <= by}.  Two Sutherland–Hodgman
// stages chained as a stream — every vertex the first stage emits is clipped by the second at once and goes straight
// into a running shoelace sum — so no vertex list exists and everything stays in registers.  (A pedestrian walking
// off the end of the road while it crosses the kerb stays in this case for dozens of steps; with the local-memory
// clipper below one such env made its whole warp several times slower.)  Disjoint boxes give area 0, which is what
// the reference returns too (intersects() false -> 0).
template <typename R>
__device__ __noinline__ R corner_share(Pose<R> a, R sx, R bx, R sy, R by) {
  Quad<R> q;
  make_box(a.length, a.width, a.theta, a.c, a.s, a.x, a.y, q);
  R acc = R(0), fx2 = R(0), fy2 = R(0), px2 = R(0), py2 = R(0);
  bool have2 = false;
  auto emit2 = [&](R x, R y) {   // vertex of the final polygon, in order
    if (have2) acc += px2 * y - x * py2;
    else { fx2 = x; fy2 = y; have2 = true; }
    px2 = x; py2 = y;
  };
  R f1x = R(0), f1y = R(0), f1d = R(0), q1x = R(0), q1y = R(0), q1d = R(0);
  bool have1 = false;
  auto crossing = [](R da, R db) { return (da > R(0) && db < R(0)) || (da < R(0) && db > R(0)); };
  auto emit1 = [&](R x, R y) {   // vertex of box ∩ first half-plane: clip the edge from the previous one by the second
    const R d = by - sy * y;
    if (!have1) {
      f1x = x; f1y = y; f1d = d; have1 = true;
    } else if (crossing(q1d, d)) {
      const R t = fast_div(q1d, q1d - d);
      emit2(q1x + t * (x - q1x), q1y + t * (y - q1y));
    }
    if (d >= R(0)) emit2(x, y);
    q1x = x; q1y = y; q1d = d;
  };
  R d1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) d1[i] = bx - sx * q.x[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    if (d1[i] >= R(0)) emit1(q.x[i], q.y[i]);
    if (crossing(d1[i], d1[j])) {
      const R t = fast_div(d1[i], d1[i] - d1[j]);
      emit1(q.x[i] + t * (q.x[j] - q.x[i]), q.y[i] + t * (q.y[j] - q.y[i]));
    }
  }
  if (have1 && crossing(q1d, f1d)) {   // closing edge of the first stage's polygon
    const R t = fast_div(q1d, q1d - f1d);
    emit2(q1x + t * (f1x - q1x), q1y + t * (f1y - q1y));
  }
  if (have2) acc += px2 * fy2 - fx2 * py2;
  return fast_div(rabs(acc) * R(0.5), a.length * a.width);
}

// area(box ∩ Q) for a rotated box that CONTAINS the road corner (the apex of the open quadrant Q outside both crossed edges),
// in box coordinates: p along the length, q across, the box is [-hl, hl] x [-hw, hw], the apex P = (pu, pv) lies inside it, and
// Q is the 90-degree wedge spanned from P by dX = (ux, vx) and dY = (uy, vy), the road's edge directions seen from the box.
// Each ray leaves the box through one side; between the two exits the wedge takes in at most two box vertices (from an
// interior point three consecutive vertices subtend more than 90 degrees), so the region is a fan of at most three triangles
// around P.  Two divisions per ray, no vertex list, no loop over polygon edges.
template <typename R>
__device__ __forceinline__ R wedge_in_box_area(R pu, R pv, R hl, R hw, R ux, R uy, R vx, R vy) {
  // sides counter-clockwise: 0 p = +hl, 1 q = +hw, 2 p = -hl, 3 q = -hw; vertex k follows side k counter-clockwise
  auto leave = [&](R dp, R dq, R& ep, R& eq) {
    const R bp = dp > R(0) ? hl : -hl, bq = dq > R(0) ? hw : -hw;
    const R tp = fast_div(bp - pu, dp), tq = fast_div(bq - pv, dq);   // both >= 0: P is inside; dp, dq != 0: the box is rotated
    if (tp <= tq) { ep = bp; eq = pv + tp * dq; return dp > R(0) ? 0 : 2; }
    ep = pu + tq * dp; eq = bq; return dq > R(0) ? 1 : 3;
  };
  R e1p, e1q, e2p, e2q;
  const int s1 = leave(ux, vx, e1p, e1q), s2 = leave(uy, vy, e2p, e2q);
  const int dir = (ux * vy - vx * uy) > R(0) ? 1 : 3;   // +1 / -1 mod 4: the sense in which dX turns towards dY
  R acc = R(0), cp = e1p - pu, cq = e1q - pv;
  int k = s1;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (k != s2) {
      const int vi = dir == 1 ? k : (k + 3) & 3;        // the vertex between side k and the next side in that sense
      const R wp = ((vi == 0 || vi == 3) ? hl : -hl) - pu, wq = (vi <= 1 ? hw : -hw) - pv;
      acc += rabs(cp * wq - cq * wp);
      cp = wp; cq = wq;
      k = (k + dir) & 3;
    }
  }
  acc += rabs(cp * (e2q - pv) - cq * (e2p - pu));
  return acc * R(0.5);
}

// The same share without clipping whenever possible.  With Hx, Hy the half-planes of the two crossed edges and Q the open
// quadrant outside both,  area(box ∩ Hx ∩ Hy) = area(box ∩ Hx) + area(box ∩ Hy) − area(box) + area(box ∩ Q)
// (inclusion-exclusion).  The first two terms are kerb shares (closed form); box ∩ Q is empty when a box axis separates the
// box from Q and a rectangle when the box is axis-aligned (a pedestrian that has turned onto ±pi/2, the case that actually
// occurs: it stays on a road corner for the ~65 steps of its kerb crossing, and the ~500 dependent instructions of the
// streaming clipper made its warp — and with it the whole CTA of a fused replay launch — run at half speed); a rotated box
// that contains the road corner gets the wedge area above.  Only a rotated box that reaches into Q without containing the
// corner still goes to corner_share.
//   tx, ty   signed distance of the box centre from the crossed x-edge / y-edge, positive inside the road
//   sx, sy   outward direction (+-1) of those edges; bx, by as for corner_share
template <typename R>
__device__ __noinline__ R corner_share_closed(Pose<R> a, R tx, R ty, R sx, R bx, R sy, R by, R tau) {
  const R hl = a.length * R(0.5), hw = a.width * R(0.5);
  const R ac = rabs(a.c), as = rabs(a.s);
  const R ex = ac * hl + as * hw, ey = as * hl + ac * hw;
  const R ox = ex - tx, oy = ey - ty;   // how far the box's bounding box reaches beyond each of the two edges
  R outside = R(0);                     // area(box ∩ Q)
  if (ox > R(0) && oy > R(0)) {
    if (rmin(ac, as) < (sizeof(R) == 8 ? R(1e-9) : R(1e-5))) {
      outside = rmin(ox, ex + ex) * rmin(oy, ey + ey);
    } else {
      // in the frame where Q is the first quadrant seen from the road corner: corner - centre = (tx, ty), box axes u, v
      const R ux = sx * a.c, uy = sy * a.s, vx = -(sx * a.s), vy = sy * a.c;
      const R pu = ux * tx + uy * ty, pv = vx * tx + vy * ty;
      bool apart = false;   // an axis pointing into Q along which the whole box lies before the corner
      if (ux >= R(0) && uy >= R(0)) apart |= pu >= hl + tau;
      if (ux <= R(0) && uy <= R(0)) apart |= -pu >= hl + tau;
      if (vx >= R(0) && vy >= R(0)) apart |= pv >= hw + tau;
      if (vx <= R(0) && vy <= R(0)) apart |= -pv >= hw + tau;
      if (!apart) {
        // the road corner inside the box (a car that swerves over the kerb while still astride the end of its road): the
        // wedge area in closed form; only a box that reaches into Q without containing the corner is clipped
        if (rabs(pu) < hl && rabs(pv) < hw) outside = wedge_in_box_area(pu, pv, hl, hw, ux, uy, vx, vy);
        else return corner_share(a, sx, bx, sy, by);
      }
    }
  }
  const R p = ((kerb_share(tx, ac * hl, as * hw) + kerb_share(ty, as * hl, ac * hw)) - R(1)) + fast_div(outside, a.length * a.width);
  return rmax(R(0), rmin(R(1), p));
}

// Share of a body box lying on a road by the general predicates: rotated roads, road corners, near-tangent
// configurations.
template <typename R>
__device__ __noinline__ Share<R> road_share_general(Pose<R> a, const Quad<R>* road, R tau) {
  Quad<R> box;
  make_box(a.length, a.width, a.theta, a.c, a.s, a.x, a.y, box);
  const Quad<R> other = *road;
  return percentage_of(box, other, tau);
}

}  // namespace cav
