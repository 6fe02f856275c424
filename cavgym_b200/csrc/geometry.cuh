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
template <typename R>
__device__ __forceinline__ bool stopping_zones(const DevType<R>& k, R x, R y, R v, R theta, R c, R s, R steer,
                                               Quad<R>& braking, Quad<R>& reaction) {
  const R bd = (v * v) / (R(2) * -k.amin);
  const R rd = v * R(0.675);
  const R td = bd + rd;
  if (td == R(0) || !(steer == R(0))) return false;
  const R hx = k.length * R(0.5), hw = k.width * R(0.5);
  R ax, ay;
  if (theta == R(0)) { ax = x + hx; ay = y + R(0); }
  else { ax = x + ((c * hx) - (s * R(0))); ay = y + ((s * hx) + (c * R(0))); }
  // make_rectangle(td, width, rear_offset=0): rear = 0, front = td, left = +hw, right = -hw
  const R lx[4] = {R(0), td, td, R(0)};
  const R ly[4] = {hw, hw, -hw, -hw};
  R zx[4], zy[4];
  if (theta == R(0)) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { zx[i] = ax + lx[i]; zy[i] = ay + ly[i]; }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      zx[i] = ax + ((c * lx[i]) - (s * ly[i]));
      zy[i] = ay + ((s * lx[i]) + (c * ly[i]));
    }
  }
  const R p = bd / td, q = R(1) - p;
  const R lsx = (zx[0] * q) + (zx[1] * p), lsy = (zy[0] * q) + (zy[1] * p);
  const R rsx = (zx[3] * q) + (zx[2] * p), rsy = (zy[3] * q) + (zy[2] * p);
  braking.x[0] = zx[0]; braking.y[0] = zy[0];
  braking.x[1] = lsx;   braking.y[1] = lsy;
  braking.x[2] = rsx;   braking.y[2] = rsy;
  braking.x[3] = zx[3]; braking.y[3] = zy[3];
  reaction.x[0] = lsx;   reaction.y[0] = lsy;
  reaction.x[1] = zx[1]; reaction.y[1] = zy[1];
  reaction.x[2] = zx[2]; reaction.y[2] = zy[2];
  reaction.x[3] = rsx;   reaction.y[3] = rsy;
  return true;
}

// max over the edges of A of (min over the vertices of B of the signed outside distance).
// > 0  <=>  some edge line of A has all of B strictly outside (a separating axis).
template <typename R>
__device__ __forceinline__ R separation(const Quad<R>& A, const Quad<R>& B) {
  R best = -INFINITY;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    const R ax = A.x[i], ay = A.y[i], ex = A.x[j] - ax, ey = A.y[j] - ay;
    const R len2 = ex * ex + ey * ey;
    R m = INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) m = rmin(m, ex * (B.y[k] - ay) - ey * (B.x[k] - ax));
    if (len2 > R(0)) best = rmax(best, m / rsqrt_(len2));
  }
  return best;
}

// Full separating-axis test (8 edge normals).  Rare path: called only when the AABBs overlap.
template <typename R>
__device__ __noinline__ bool sat_intersects(const Quad<R>& A, const Quad<R>& B, R tau, bool& tangent) {
  const R m = rmax(separation(A, B), separation(B, A));
  if (rabs(m) < tau) tangent = true;
  return m <= R(0);
}

// Shape.intersects for two convex quads.  AABB rejection first: exact and conservative.
template <typename R>
__device__ __forceinline__ bool intersects(const Quad<R>& A, const Aabb<R>& a, const Quad<R>& B, const Aabb<R>& b, R tau,
                                           bool& tangent) {
  if (aabb_gap(a, b) > tau) return false;
  return sat_intersects(A, B, tau, tangent);
}

// outer.contains(inner): every vertex of inner inside or on every edge of outer.
template <typename R>
__device__ __forceinline__ bool contains(const Quad<R>& outer, const Quad<R>& inner, R tau, bool& tangent) {
  R worst = -INFINITY;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    const R ax = outer.x[i], ay = outer.y[i], ex = outer.x[j] - ax, ey = outer.y[j] - ay;
    const R len2 = ex * ex + ey * ey;
    R m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) m = rmax(m, ex * (inner.y[k] - ay) - ey * (inner.x[k] - ax));
    if (len2 > R(0)) worst = rmax(worst, m / rsqrt_(len2));
  }
  if (rabs(worst) < tau) tangent = true;
  return worst <= R(0);
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
// Callers reject AABB-disjoint pairs first, so this is the rare path.
template <typename R>
__device__ __noinline__ R percentage_intersects(const Quad<R>& self, const Quad<R>& other, R tau, bool& tangent) {
  const R m = rmax(separation(self, other), separation(other, self));
  if (rabs(m) < tau) tangent = true;
  if (!(m <= R(0))) return R(0);
  if (contains(other, self, tau, tangent)) return R(1);
  return clip_area(self, other) / quad_area(self);
}

}  // namespace cav
