/* cavgym_oracle.c — CPU restatement of CAV-Gym's stepping hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (cavgym_b200/, include/) may
 * link, import or call this file; it is the checker used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
 *
 * Each function cites the reference file:line it follows (paths relative to the
 * CAV-Gym tree).  Arithmetic is fp64 in the reference's exact operation order
 * (compile with -ffp-contract=off; `x ** 2` is pow(x, 2.0) as CPython does), so
 * body state is BIT-EQUAL to the Python reference on the same libm; this is
 * pinned by tests/test_oracle_golden.py against the tests/golden .npz fixtures, which were
 * produced by running the unmodified reference (oracle/gen_golden.py).
 *
 * Geometry predicates: the reference delegates to Shapely ~=1.7 / GEOS 3.8
 * (library/geometry.py:74-93), which is NOT vendored.  GEOS decides orientation
 * robustly, so the predicates here are exact on their fp64 inputs (float filter +
 * expansion arithmetic).  PARITY UNPINNED at that boundary: the reference holds no
 * tests or golden vectors for it; the fixtures pin this file against the
 * reference's source run over an exact convex stand-in (oracle/standins/shapely).
 *
 * RNG: the reference shares one MT19937 RandomState (config.py:275); the batched
 * engine uses a counter-based Philox4x32-10 stream instead (north star).  The same
 * Philox keying is restated here so engine and oracle agree draw-for-draw; the
 * MT19937 draws of a reference run can be replayed through the override buffers.
 */
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "../include/cavgym.h"

#define STEERING_ERROR 0.0000000000001   /* bodies.py:19 */
#define REACTION_TIME 0.675              /* bodies.py:18 */
#define TARGET_ERROR 0.000000000000001   /* examples/agents/dynamic_body.py:8 */

typedef struct CavOracle {
  CavScenario sc;
  CavBody* bodies;
  CavSpawn* spawns;
  int64_t n, shard, t_global;
  uint64_t seed;
  double tau;
  int threads;
  double *state, *action, *agent;
  int32_t *liveness, *t_ep, *episode, *winner;
  uint8_t *done, *err;
  const double* uni_override;    /* [M][CAV_DRAWS][N] */
  const double* spawn_override;  /* [M][5][N] */
  int64_t stats[CAV_N_STATS];
} CavOracle;

/* ------------------------------------------------------------------ exact predicates */

static inline void two_sum(double a, double b, double* x, double* y) {
  double s = a + b, bv = s - a, av = s - bv;
  *x = s;
  *y = (a - av) + (b - bv);
}

static inline void two_prod(double a, double b, double* x, double* y) {
  double p = a * b;
  *x = p;
  *y = fma(a, b, -p);
}

/* Sign of the exact sum of n doubles (Shewchuk grow-expansion). */
static int exact_sum_sign(const double* t, int n) {
  double e[16];
  int m = 0;
  for (int i = 0; i < n; ++i) {
    double q = t[i];
    for (int j = 0; j < m; ++j) {
      double s, r;
      two_sum(q, e[j], &s, &r);
      e[j] = r;
      q = s;
    }
    e[m++] = q;
  }
  for (int j = m - 1; j >= 0; --j) {
    if (e[j] > 0) return 1;
    if (e[j] < 0) return -1;
  }
  return 0;
}

/* Sign of (b-a) x (c-a), exact on fp64 inputs. */
static int orient_sign(double ax, double ay, double bx, double by, double cx, double cy) {
  double l = (bx - ax) * (cy - ay), r = (by - ay) * (cx - ax);
  double det = l - r, bound = 8.0 * 2.220446049250313e-16 * (fabs(l) + fabs(r));
  if (det > bound) return 1;
  if (det < -bound) return -1;
  double t[12];
  two_prod(bx, cy, &t[0], &t[1]);
  two_prod(-bx, ay, &t[2], &t[3]);
  two_prod(-ax, cy, &t[4], &t[5]);
  two_prod(-by, cx, &t[6], &t[7]);
  two_prod(by, ax, &t[8], &t[9]);
  two_prod(ay, cx, &t[10], &t[11]);
  return exact_sum_sign(t, 12);
}

/* Orientation of a quad's ring: +1 counter-clockwise, -1 clockwise (float; quads here are far from degenerate). */
static int quad_winding(const CavQuad* q) {
  double a = 0.0;
  for (int i = 0; i < 4; ++i) {
    int j = (i + 1) & 3;
    a += q->x[i] * q->y[j] - q->x[j] * q->y[i];
  }
  return a >= 0 ? 1 : -1;
}

/* Signed distance of p outside the directed edge a->b of a ring with winding w (>0 = outside). */
static double outside_distance(const CavQuad* q, int w, int i, double px, double py) {
  int j = (i + 1) & 3;
  double ex = q->x[j] - q->x[i], ey = q->y[j] - q->y[i];
  double cr = ex * (py - q->y[i]) - ey * (px - q->x[i]);
  double len = sqrt(ex * ex + ey * ey);
  return len > 0 ? (-w * cr) / len : -INFINITY;
}

/* Does some edge line of A have every vertex of B strictly outside?  Also tracks the float margin. */
static int separates(const CavQuad* A, const CavQuad* B, double* margin) {
  int w = quad_winding(A), found = 0;
  for (int i = 0; i < 4; ++i) {
    int j = (i + 1) & 3;
    if (A->x[i] == A->x[j] && A->y[i] == A->y[j]) continue;
    int all_out = 1;
    double m = INFINITY;
    for (int k = 0; k < 4; ++k) {
      int s = orient_sign(A->x[i], A->y[i], A->x[j], A->y[j], B->x[k], B->y[k]) * w;
      if (s >= 0) all_out = 0; /* inside or on the line */
      double d = outside_distance(A, w, i, B->x[k], B->y[k]);
      if (d < m) m = d;
    }
    if (m > *margin) *margin = m;
    if (all_out) found = 1;
  }
  return found;
}

/* Shape.intersects (geometry.py:74-75): closed-set intersection of two convex quads.
 * *tangent is set when the float separation margin is within tau of zero. */
static int quad_intersects(const CavQuad* A, const CavQuad* B, double tau, int* tangent) {
  /* Bounding boxes more than a pixel apart: disjoint, and nowhere near tangent (for the rectangles of this path the
   * largest edge-normal separation is >= gap / sqrt(2) >> tau).  Purely an early-out of the all-pairs loop
   * (environment.py:156-177 at M = 320 meets 51,040 pairs per step); the decision and the flag are unchanged. */
  {
    double ax0 = A->x[0], ax1 = A->x[0], ay0 = A->y[0], ay1 = A->y[0], bx0 = B->x[0], bx1 = B->x[0], by0 = B->y[0], by1 = B->y[0];
    for (int i = 1; i < 4; ++i) {
      ax0 = fmin(ax0, A->x[i]); ax1 = fmax(ax1, A->x[i]); ay0 = fmin(ay0, A->y[i]); ay1 = fmax(ay1, A->y[i]);
      bx0 = fmin(bx0, B->x[i]); bx1 = fmax(bx1, B->x[i]); by0 = fmin(by0, B->y[i]); by1 = fmax(by1, B->y[i]);
    }
    if (ax0 - bx1 > 1.0 || bx0 - ax1 > 1.0 || ay0 - by1 > 1.0 || by0 - ay1 > 1.0) return 0;
  }
  double margin = -INFINITY;
  int sep = separates(A, B, &margin) | separates(B, A, &margin);
  if (fabs(margin) < tau) *tangent = 1;
  return !sep;
}

/* Shape.contains (geometry.py:77-78): A.contains(B) for convex quads. */
static int quad_contains(const CavQuad* A, const CavQuad* B, double tau, int* tangent) {
  int w = quad_winding(A), inside = 1;
  double margin = -INFINITY;
  for (int i = 0; i < 4; ++i) {
    int j = (i + 1) & 3;
    if (A->x[i] == A->x[j] && A->y[i] == A->y[j]) continue;
    for (int k = 0; k < 4; ++k) {
      if (orient_sign(A->x[i], A->y[i], A->x[j], A->y[j], B->x[k], B->y[k]) * w < 0) inside = 0;
      double d = outside_distance(A, w, i, B->x[k], B->y[k]);
      if (d > margin) margin = d;
    }
  }
  if (fabs(margin) < tau) *tangent = 1;
  return inside;
}

static double ring_area(const double* x, const double* y, int n) {
  double a = 0.0;
  for (int i = 0; i < n; ++i) {
    int j = (i + 1 == n) ? 0 : i + 1;
    a += x[i] * y[j] - x[j] * y[i];
  }
  return fabs(a) * 0.5;
}

/* area(A ∩ B) by Sutherland–Hodgman clipping of A against the half-planes of B
 * (stands for Shapely's intersection(...).area, geometry.py:85-87). */
static double clip_area(const CavQuad* A, const CavQuad* B) {
  double sx[16], sy[16], ox[16], oy[16];
  int n = 4, w = quad_winding(B);
  for (int i = 0; i < 4; ++i) { sx[i] = A->x[i]; sy[i] = A->y[i]; }
  for (int i = 0; i < 4 && n > 0; ++i) {
    int j = (i + 1) & 3;
    double ax = B->x[i], ay = B->y[i], ex = B->x[j] - ax, ey = B->y[j] - ay;
    if (ex == 0 && ey == 0) continue;
    int m = 0;
    for (int k = 0; k < n; ++k) {
      int l = (k + 1 == n) ? 0 : k + 1;
      double sp = w * (ex * (sy[k] - ay) - ey * (sx[k] - ax));
      double sq = w * (ex * (sy[l] - ay) - ey * (sx[l] - ax));
      if (sp >= 0) { ox[m] = sx[k]; oy[m] = sy[k]; ++m; }
      if ((sp > 0 && sq < 0) || (sp < 0 && sq > 0)) {
        double t = sp / (sp - sq);
        ox[m] = sx[k] + t * (sx[l] - sx[k]);
        oy[m] = sy[k] + t * (sy[l] - sy[k]);
        ++m;
      }
    }
    n = m;
    memcpy(sx, ox, sizeof(double) * n);
    memcpy(sy, oy, sizeof(double) * n);
  }
  return n < 3 ? 0.0 : ring_area(sx, sy, n);
}

/* Shape.percentage_intersects (geometry.py:80-87). */
static double percentage_intersects(const CavQuad* self, const CavQuad* other, double tau, int* tangent) {
  if (!quad_intersects(self, other, tau, tangent)) return 0.0;
  if (quad_contains(other, self, tau, tangent)) return 1.0;
  return clip_area(self, other) / ring_area(self->x, self->y, 4);
}

/* ------------------------------------------------------------------ shapes */

/* make_rectangle(length, width, rear_offset, left_offset=0.5) (geometry.py:241-251)
 * followed by ConvexQuadrilateral.transform(theta, (px, py)) (geometry.py:117-126,
 * Point.transform/rotate :24-31,41-45). */
static void make_box(double length, double width, double rear_offset, double theta, double px, double py, CavQuad* q) {
  double rear = 0.0 - (length * rear_offset), front = 0.0 + (length * (1 - rear_offset));
  double left = 0.0 + (width * 0.5), right = 0.0 - (width * (1 - 0.5));
  double lx[4] = {rear, front, front, rear}, ly[4] = {left, left, right, right};
  if (theta == 0) {
    for (int i = 0; i < 4; ++i) { q->x[i] = px + lx[i]; q->y[i] = py + ly[i]; }
  } else {
    double c = cos(theta), s = sin(theta);
    for (int i = 0; i < 4; ++i) {
      double rx = (c * lx[i]) - (s * ly[i]), ry = (s * lx[i]) + (c * ly[i]);
      q->x[i] = px + rx;
      q->y[i] = py + ry;
    }
  }
}

/* DynamicBody.stopping_zones (bodies.py:122-135) + split_longitudinally (geometry.py:176-191).
 * Returns 0 when the zones are None. */
static int stopping_zones(const CavBodyType* k, const double* st, double steer, CavQuad* braking, CavQuad* reaction) {
  double v = st[2], theta = st[3];
  double bd = pow(v, 2.0) / (2 * -k->min_throttle);
  double rd = v * REACTION_TIME;
  double td = bd + rd;
  if (td == 0) return 0;
  if (!(steer == 0)) return 0;
  double ax, ay; /* Point(length*0.5, 0).transform(theta, position) */
  double hx = k->length * 0.5;
  if (theta == 0) { ax = st[0] + hx; ay = st[1] + 0.0; }
  else {
    double c = cos(theta), s = sin(theta);
    ax = st[0] + ((c * hx) - (s * 0.0));
    ay = st[1] + ((s * hx) + (c * 0.0));
  }
  CavQuad z;
  make_box(td, k->width, 0.0, theta, ax, ay, &z);
  double p = bd / td;
  double lsx = (z.x[0] * (1 - p)) + (z.x[1] * p), lsy = (z.y[0] * (1 - p)) + (z.y[1] * p);
  double rsx = (z.x[3] * (1 - p)) + (z.x[2] * p), rsy = (z.y[3] * (1 - p)) + (z.y[2] * p);
  braking->x[0] = z.x[0]; braking->y[0] = z.y[0];
  braking->x[1] = lsx;    braking->y[1] = lsy;
  braking->x[2] = rsx;    braking->y[2] = rsy;
  braking->x[3] = z.x[3]; braking->y[3] = z.y[3];
  reaction->x[0] = lsx;    reaction->y[0] = lsy;
  reaction->x[1] = z.x[1]; reaction->y[1] = z.y[1];
  reaction->x[2] = z.x[2]; reaction->y[2] = z.y[2];
  reaction->x[3] = rsx;    reaction->y[3] = rsy;
  return 1;
}

/* ------------------------------------------------------------------ kinematics */

static inline double py_max(double a, double b) { return b > a ? b : a; } /* max(a, b) */
static inline double py_min(double a, double b) { return b < a ? b : a; } /* min(a, b) */

/* DynamicBody.step (bodies.py:214-275).  st = x, y, v, theta (in/out); returns the snapped steering angle. */
static double dynamic_body_step(const CavBodyType* k, double* st, double throttle, double steer, double dt) {
  if (fabs(steer) < STEERING_ERROR) steer = 0.0;
  double x = st[0], y = st[1], v = st[2], th = st[3];
  double v1 = py_max(k->min_velocity, py_min(k->max_velocity, v + (throttle * dt)));
  double d = v * dt, c = cos(th), s = sin(th);
  if (fabs(steer) == 0) {
    st[0] = x + d * c;
    st[1] = y + d * s;
    st[2] = v1;
  } else {
    double wbo = k->wheelbase / 2.0;
    double rx = x - wbo * c, ry = y - wbo * s;
    double kk = k->wheelbase / tan(steer);
    double cx = rx - kk * s, cy = ry + kk * c;
    double dx = x - cx, dy = y - cy;
    double theta = (steer < 0 ? -1 : 1) * (d / sqrt(pow(dx, 2.0) + pow(dy, 2.0)));
    double ct = cos(theta), sn = sin(theta);
    st[0] = cx + dx * ct - dy * sn;
    st[1] = cy + dx * sn + dy * ct;
    st[2] = v1;
    double ot = th + theta;
    st[3] = atan2(sin(ot), cos(ot));
  }
  return steer;
}

/* ------------------------------------------------------------------ Philox4x32-10 */

static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void cav_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

/* 53-bit uniform in [0,1) from two words, as numpy's legacy random_sample builds it. */
static inline double u53(uint32_t a, uint32_t b) {
  return ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0;
}

/* Stream layout: key = seed; counter = (env_lo, env_hi | kind<<8 | body<<16, episode, timestep). */
enum { KIND_AGENT0 = 0, KIND_AGENT1 = 1, KIND_SPAWN0 = 2, KIND_SPAWN1 = 3, KIND_SPAWN2 = 4 };

static void draw_block(const CavOracle* o, int64_t env, int body, int kind, uint32_t episode, uint32_t t, double u[2]) {
  uint64_t g = (uint64_t)(o->shard + env);
  uint32_t w[4];
  philox4x32_10((uint32_t)g, (uint32_t)((g >> 32) & 0xFFu) | ((uint32_t)kind << 8) | ((uint32_t)body << 16), episode, t,
                (uint32_t)o->seed, (uint32_t)(o->seed >> 32), w);
  u[0] = u53(w[0], w[1]);
  u[1] = u53(w[2], w[3]);
}

/* ------------------------------------------------------------------ spawn */

static double triangle_area(double rx, double ry, double flx, double fly, double frx, double fry) { /* geometry.py:352-354 */
  double numerator = rx * (fly - fry) + flx * (fry - ry) + frx * (ry - fly);
  return fabs(numerator / 2);
}

/* SpawnPedestrian.spawn (bodies.py:302-312): numpy legacy choice(p=...) = searchsorted(cumsum(p)/last, u, 'right');
 * ConvexQuadrilateral.random_point (geometry.py:223-229), Triangle.random_point (:369-377). */
static void spawn_body(const CavSpawn* sp, const double u[5], double* st) {
  double areas[CAV_MAX_SPAWN_BOXES], larea[CAV_MAX_SPAWN_BOXES], total = 0.0;
  for (int i = 0; i < sp->n_boxes; ++i) {
    const CavQuad* q = &sp->boxes[i];
    /* triangles() geometry.py:207-218: left = (FL, FR, RL), right = (RR, RL, FR) */
    larea[i] = triangle_area(q->x[1], q->y[1], q->x[2], q->y[2], q->x[0], q->y[0]);
    double r = triangle_area(q->x[3], q->y[3], q->x[0], q->y[0], q->x[2], q->y[2]);
    areas[i] = 0.0 + larea[i] + r;
    total += areas[i];
  }
  double cdf[CAV_MAX_SPAWN_BOXES], acc = 0.0;
  for (int i = 0; i < sp->n_boxes; ++i) { acc += areas[i] / total; cdf[i] = acc; }
  int box = 0;
  for (int i = 0; i < sp->n_boxes; ++i) if (cdf[i] / cdf[sp->n_boxes - 1] <= u[0]) ++box;
  if (box >= sp->n_boxes) box = sp->n_boxes - 1;
  const CavQuad* q = &sp->boxes[box];
  double la = larea[box];
  double sa = la + triangle_area(q->x[3], q->y[3], q->x[0], q->y[0], q->x[2], q->y[2]);
  double f = la / sa;
  double c0 = f, c1 = f + (1 - f);
  int tri = ((c0 / c1) <= u[1]) + ((c1 / c1) <= u[1]);
  if (tri > 1) tri = 1;
  double rx, ry, flx, fly, frx, fry;
  if (tri == 0) { rx = q->x[1]; ry = q->y[1]; flx = q->x[2]; fly = q->y[2]; frx = q->x[0]; fry = q->y[0]; }
  else          { rx = q->x[3]; ry = q->y[3]; flx = q->x[0]; fly = q->y[0]; frx = q->x[2]; fry = q->y[2]; }
  double a = u[2], b = u[3];
  if (a + b > 1) { a = 1 - a; b = 1 - b; }
  st[0] = (rx + (flx - rx) * a) + (frx - rx) * b;
  st[1] = (ry + (fly - ry) * a) + (fry - ry) * b;
  st[2] = sp->velocity;
  int oi = (int)floor(u[4] * sp->n_orientations);
  if (oi >= sp->n_orientations) oi = sp->n_orientations - 1;
  st[3] = sp->orientations[oi];
}

void cav_oracle_spawn(const CavSpawn* sp, const double u[5], double st[4]) { spawn_body(sp, u, st); }

/* ------------------------------------------------------------------ agents */

/* make_steering_action (examples/agents/dynamic_body.py:33-49); target NaN = None. */
static double make_steering_action(const CavBodyType* k, const double* st, double dt, double target) {
  double v = st[2], steer;
  if (v == 0 || isnan(target)) {
    steer = 0.0;
  } else {
    double diff = target - st[3];
    double tta = atan2(sin(diff), cos(diff));
    double csa = tta < 0 ? k->min_steering_angle : k->max_steering_angle;
    double wb = k->wheelbase;
    double mta = (csa < 0 ? -2 : 2) * dt * v / sqrt(pow(wb, 2.0) * (1 + 4 / pow(tan(csa), 2.0)));
    double ta = (tta / mta > 1) ? mta : tta;
    steer = (ta < 0 ? -1 : 1) * atan(2 * wb * sqrt(pow(ta, 2.0) / (4 * pow(v, 2.0) * pow(dt, 2.0) - pow(wb, 2.0) * pow(ta, 2.0))));
  }
  return py_min(k->max_steering_angle, py_max(k->min_steering_angle, steer));
}

static inline double point_distance(double sx, double sy, double ox, double oy) { /* Point.distance geometry.py:18-19 */
  return sqrt(pow(oy - sy, 2.0) + pow(ox - sx, 2.0));
}

/* CrossingAgent.choose_crossing_action (pedestrian.py:50-69).  ag = initial_distance, wx, wy, target, prior. */
static double choose_crossing_action(const CavOracle* o, const CavBodyType* k, const double* st, double* ag, int condition) {
  if (isnan(ag[1]) && isnan(ag[3]) && condition) {
    const double* cl = o->sc.centre_line; /* Line.closest_point_from geometry.py:412-416 */
    double dx = cl[2] - cl[0], dy = cl[3] - cl[1];
    double denominator = (dx * dx) + (dy * dy);
    double a = (dy * (st[1] - cl[1]) + dx * (st[0] - cl[0])) / denominator;
    double cx = cl[0] + a * dx, cy = cl[1] + a * dy;
    double rel = atan2(cy - st[1], cx - st[0]);
    if (isnan(ag[0])) ag[0] = point_distance(st[0], st[1], cx, cy);
    ag[1] = cx + ag[0] * cos(rel);
    ag[2] = cy + ag[0] * sin(rel);
    ag[3] = atan2(ag[2] - st[1], ag[1] - st[0]);
    ag[4] = st[3];
  }
  return make_steering_action(k, st, o->sc.time_resolution, ag[3]);
}

/* CrossingAgent.process_feedback (pedestrian.py:36-48) on the post-step state. */
static void crossing_feedback(const double* st, double* ag) {
  if (!isnan(ag[1])) {
    double distance = point_distance(st[0], st[1], ag[1], ag[2]);
    if (distance < 1) { ag[1] = NAN; ag[2] = NAN; ag[3] = ag[4]; ag[4] = NAN; }
  }
  if (!isnan(ag[3])) {
    double diff = ag[3] - st[3];
    if (fabs(atan2(sin(diff), cos(diff))) < TARGET_ERROR) ag[3] = NAN;
  }
}

/* ------------------------------------------------------------------ per-env transition */

#define ST(o, b, c, e) ((o)->state[((int64_t)(b) * 4 + (c)) * (o)->n + (e)])
#define AC(o, b, c, e) ((o)->action[((int64_t)(b) * 2 + (c)) * (o)->n + (e)])
#define AG(o, b, c, e) ((o)->agent[((int64_t)(b) * CAV_AGENT_WORDS + (c)) * (o)->n + (e)])

static void score_episode(CavOracle* o, int64_t e, int64_t* stats) { /* reporting.py:227-243 */
  int m = o->sc.n_bodies;
  int64_t t = o->t_ep[e];
  stats[CAV_STAT_EPISODES] += 1;
  stats[CAV_STAT_SUM_T] += t;
  stats[CAV_STAT_SUM_T2] += t * t;
  if (o->winner[e] > 0) {
    int64_t score = 0;
    for (int b = 1; b < m; ++b) score -= o->liveness[(int64_t)b * o->n + e];
    stats[CAV_STAT_INTERESTING] += 1;
    stats[CAV_STAT_SUM_SCORE] += score;
    stats[CAV_STAT_SUM_SCORE2] += score * score;
    stats[CAV_STAT_SUM_T_INTERESTING] += t;
    stats[CAV_STAT_SUM_T2_INTERESTING] += t * t;
  }
}

static void reset_env(CavOracle* o, int64_t e, const double* init_state) {
  int m = o->sc.n_bodies;
  o->episode[e] += 1;
  for (int b = 0; b < m; ++b) {
    const CavBody* body = &o->bodies[b];
    double st[4] = {body->init_state[0], body->init_state[1], body->init_state[2], body->init_state[3]};
    if (init_state) {
      for (int c = 0; c < 4; ++c) st[c] = init_state[((int64_t)b * 4 + c) * o->n + e];
    } else if ((body->flags & CAV_FLAG_SPAWN) && body->spawn_id >= 0) {
      double u[5];
      if (o->spawn_override) {
        for (int c = 0; c < 5; ++c) u[c] = o->spawn_override[((int64_t)b * 5 + c) * o->n + e];
      } else {
        double w[2];
        draw_block(o, e, b, KIND_SPAWN0, (uint32_t)o->episode[e], 0, w); u[0] = w[0]; u[1] = w[1];
        draw_block(o, e, b, KIND_SPAWN1, (uint32_t)o->episode[e], 0, w); u[2] = w[0]; u[3] = w[1];
        draw_block(o, e, b, KIND_SPAWN2, (uint32_t)o->episode[e], 0, w); u[4] = w[0];
      }
      spawn_body(&o->spawns[body->spawn_id], u, st);
    }
    for (int c = 0; c < 4; ++c) ST(o, b, c, e) = st[c];
    AC(o, b, 0, e) = 0.0; AC(o, b, 1, e) = 0.0;           /* noop_action (bodies.py:88-89,114) */
    for (int c = 0; c < CAV_AGENT_WORDS; ++c) AG(o, b, c, e) = NAN; /* pedestrian.py:27-31 */
    o->liveness[(int64_t)b * o->n + e] = 0;               /* environment.py:228 */
  }
  o->t_ep[e] = 0;
  o->done[e] = 0;
  o->winner[e] = -1;
}

typedef struct StepOut {
  double* state; double* reward; uint8_t* done; int32_t* winner; uint8_t* tangent;
} StepOut;

/* CAVEnv.step (environment.py:119-223) for env e, preceded by the agents' choose_action
 * (simulation.py:71) and followed by process_feedback (simulation.py:86-87). */
static void env_transition(CavOracle* o, int64_t e, const double* ext_actions, int64_t t_global, const StepOut* out, int64_t* stats) {
  const CavScenario* sc = &o->sc;
  const int m = sc->n_bodies;
  const int64_t n = o->n;
  const double dt = sc->time_resolution, tau = o->tau;
  double reward[CAV_MAX_BODIES];
  int tangent = 0;

  if (o->done[e]) { /* frozen until reset */
    if (out->reward) for (int b = 0; b < m; ++b) out->reward[(int64_t)b * n + e] = 0.0;
    if (out->state) for (int b = 0; b < m; ++b) for (int c = 0; c < 4; ++c) out->state[((int64_t)b * 4 + c) * n + e] = ST(o, b, c, e);
    if (out->done) out->done[e] = o->done[e] == 1;
    if (out->winner) out->winner[e] = o->winner[e];
    if (out->tangent) out->tangent[e] = 0;
    return;
  }

  /* --- joint action (simulation.py:71) from the pre-step state */
  double act[CAV_MAX_BODIES][2];
  int valid = 1;
  for (int b = 0; b < m; ++b) {
    const CavBody* body = &o->bodies[b];
    const CavBodyType* k = &sc->types[body->type_id];
    double st[4] = {ST(o, b, 0, e), ST(o, b, 1, e), ST(o, b, 2, e), ST(o, b, 3, e)};
    double a0 = AC(o, b, 0, e), a1 = AC(o, b, 1, e); /* RandomAgent.action is held (template.py:47-56) */
    double u[CAV_DRAWS] = {0, 0, 0};
    int agent = body->agent;
    if (agent == CAV_AGENT_RANDOM || agent == CAV_AGENT_RANDOM_CONSTRAINED) {
      if (o->uni_override) {
        for (int c = 0; c < CAV_DRAWS; ++c) u[c] = o->uni_override[((int64_t)b * CAV_DRAWS + c) * n + e];
      } else {
        double w[2];
        draw_block(o, e, b, KIND_AGENT0, (uint32_t)o->episode[e], (uint32_t)o->t_ep[e], w);
        u[0] = w[0]; u[1] = w[1];
      }
    }
    switch (agent) {
      case CAV_AGENT_EXTERNAL:
        a0 = ext_actions[((int64_t)b * 2 + 0) * n + e];
        a1 = ext_actions[((int64_t)b * 2 + 1) * n + e];
        break;
      case CAV_AGENT_NOOP:
        a0 = 0.0; a1 = 0.0;
        break;
      case CAV_AGENT_RANDOM: /* RandomAgent.choose_action template.py:52-56; Box.sample = low + (high-low)*u */
        if (u[0] < body->agent_epsilon) {
          if (body->kind == CAV_BODY_PELICAN) {
            a0 = floor(u[1] * 4); if (a0 > 3) a0 = 3;
            a1 = 0.0;
          } else {
            if (!o->uni_override) { double w[2]; draw_block(o, e, b, KIND_AGENT1, (uint32_t)o->episode[e], (uint32_t)o->t_ep[e], w); u[2] = w[0]; }
            a0 = k->min_throttle + (k->max_throttle - k->min_throttle) * u[1];
            a1 = k->min_steering_angle + (k->max_steering_angle - k->min_steering_angle) * u[2];
          }
        }
        break;
      case CAV_AGENT_RANDOM_CONSTRAINED: { /* pedestrian.py:72-75 */
        double ag[CAV_AGENT_WORDS];
        for (int c = 0; c < CAV_AGENT_WORDS; ++c) ag[c] = AG(o, b, c, e);
        a0 = 0.0;
        a1 = choose_crossing_action(o, k, st, ag, u[0] < body->agent_epsilon);
        for (int c = 0; c < CAV_AGENT_WORDS; ++c) AG(o, b, c, e) = ag[c];
        break;
      }
      case CAV_AGENT_PROXIMITY: { /* pedestrian.py:78-91: distance to the ego position */
        double ag[CAV_AGENT_WORDS];
        for (int c = 0; c < CAV_AGENT_WORDS; ++c) ag[c] = AG(o, b, c, e);
        int trigger = point_distance(st[0], st[1], ST(o, 0, 0, e), ST(o, 0, 1, e)) < body->agent_threshold;
        a0 = 0.0;
        a1 = choose_crossing_action(o, k, st, ag, trigger);
        for (int c = 0; c < CAV_AGENT_WORDS; ++c) AG(o, b, c, e) = ag[c];
        break;
      }
      default: break;
    }
    act[b][0] = a0; act[b][1] = a1;
    /* action_space.contains (environment.py:120): inclusive Box bounds / Discrete(4) */
    if (body->kind == CAV_BODY_PELICAN) {
      if (!(a0 == floor(a0) && a0 >= 0 && a0 <= 3)) valid = 0;
    } else if (!(a0 >= k->min_throttle && a0 <= k->max_throttle && a1 >= k->min_steering_angle && a1 <= k->max_steering_angle)) {
      valid = 0;
    }
  }
  if (!valid) { /* AssertionError before any mutation */
    o->err[e] = 1;
    if (out->reward) for (int b = 0; b < m; ++b) out->reward[(int64_t)b * n + e] = 0.0;
    if (out->state) for (int b = 0; b < m; ++b) for (int c = 0; c < 4; ++c) out->state[((int64_t)b * 4 + c) * n + e] = ST(o, b, c, e);
    if (out->done) out->done[e] = 0;
    if (out->winner) out->winner[e] = -1;
    if (out->tangent) out->tangent[e] = 0;
    return;
  }

  /* --- body.step for every body (environment.py:122-123) and bounding boxes (info(), :107) */
  CavQuad box[CAV_MAX_BODIES];
  double ego_steer = 0.0;
  for (int b = 0; b < m; ++b) {
    const CavBody* body = &o->bodies[b];
    AC(o, b, 0, e) = act[b][0]; AC(o, b, 1, e) = act[b][1];
    if (body->kind == CAV_BODY_PELICAN) { /* PelicanCrossing.step bodies.py:450-461 */
      int a = (int)act[b][0];
      if (a == 1) ST(o, b, 0, e) = 0.0; else if (a == 2) ST(o, b, 0, e) = 1.0; else if (a == 3) ST(o, b, 0, e) = 2.0;
      box[b] = body->static_box;
      continue;
    }
    const CavBodyType* k = &sc->types[body->type_id];
    double st[4] = {ST(o, b, 0, e), ST(o, b, 1, e), ST(o, b, 2, e), ST(o, b, 3, e)};
    double snapped = dynamic_body_step(k, st, act[b][0], act[b][1], dt);
    if (b == 0) ego_steer = snapped;
    for (int c = 0; c < 4; ++c) ST(o, b, c, e) = st[c];
    make_box(k->length, k->width, 0.5, st[3], st[0], st[1], &box[b]); /* bodies.py:116-117 */
  }

  /* --- rewards (environment.py:131-146) */
  const double c = sc->cost_step, W = sc->viewer_width;
  const double ego_rel = py_max(0.0, py_min(1.0, (W - ST(o, 0, 0, e)) / W));
  const double voff = fabs(ST(o, 0, 2, e) - sc->ego_maintenance_velocity) / sc->ego_max_velocity_offset;
  reward[0] = 0.0;
  reward[0] -= voff * c;
  reward[0] += (1.0 - ego_rel) * c;
  for (int b = 1; b < m; ++b) {
    double p = 0.0;
    int is_static = o->bodies[b].kind == CAV_BODY_PELICAN, unused = 0; /* static box vs static road: no libm, never tangent-flagged */
    for (int r = 0; r < sc->n_roads; ++r) {
      double q = percentage_intersects(&box[b], &sc->roads[r], tau, is_static ? &unused : &tangent);
      if (r == 0 || q > p) p = q;
    }
    reward[b] = 0.0;
    reward[b] -= p * c;
    reward[b] += ego_rel * c;
    if (!is_static && fabs(p - 0.5) < tau) tangent = 1;
    if (p > 0.5) o->liveness[(int64_t)b * n + e] += 1;
  }

  /* --- termination cascade (environment.py:148-206) */
  int terminate = 0, win_ego = 0, win_tester = -1;
  {
    double mn = INFINITY;
    for (int i = 0; i < 4; ++i) if (box[0].x[i] < mn) mn = box[0].x[i];
    if (fabs(mn - W) < tau) tangent = 1;
    if (mn > W) { terminate = 1; win_ego = 1; }
  }
  if (!terminate && sc->terminate_collisions == CAV_COLLISIONS_ALL) { /* :156-177 */
    int hit = 0;
    for (int i = 0; i < m; ++i) {
      if (o->bodies[i].kind == CAV_BODY_PELICAN) continue;
      for (int j = i + 1; j < m; ++j) {
        if (o->bodies[j].kind == CAV_BODY_PELICAN) continue;
        if (quad_intersects(&box[i], &box[j], tau, &tangent)) hit = 1;
      }
      for (int s = 0; s < sc->n_statics; ++s)
        if (quad_intersects(&box[i], &sc->statics[s], tau, &tangent)) hit = 1;
    }
    terminate = hit;
  }
  if (!terminate && sc->terminate_ego_offroad) { /* :179-181 */
    int on_road = 0;
    for (int r = 0; r < sc->n_roads; ++r) if (quad_intersects(&box[0], &sc->roads[r], tau, &tangent)) on_road = 1;
    terminate = !on_road;
  }
  CavQuad braking, reaction;
  int have_zones = 0;
  if (!terminate && (sc->terminate_collisions == CAV_COLLISIONS_EGO || sc->terminate_ego_zones)) {
    double st0[4] = {ST(o, 0, 0, e), ST(o, 0, 1, e), ST(o, 0, 2, e), ST(o, 0, 3, e)};
    have_zones = o->bodies[0].kind == CAV_BODY_DYNAMIC &&
                 stopping_zones(&sc->types[o->bodies[0].type_id], st0, ego_steer, &braking, &reaction);
  }
  if (!terminate && sc->terminate_collisions == CAV_COLLISIONS_EGO) { /* :183-193 */
    int hit = 0;
    for (int b = 1; b < m; ++b) {
      if (!(o->bodies[b].flags & CAV_FLAG_PEDESTRIAN)) continue;
      if (quad_intersects(&box[b], &box[0], tau, &tangent)) hit = 1;
      else if (have_zones && quad_intersects(&box[b], &braking, tau, &tangent)) hit = 1;
    }
    terminate = hit;
  }
  if (!terminate && sc->terminate_ego_zones) { /* :195-206 */
    if (have_zones) {
      for (int b = 1; b < m; ++b) {
        if (!(o->bodies[b].flags & CAV_FLAG_PEDESTRIAN)) continue;
        if (quad_intersects(&box[b], &reaction, tau, &tangent)) { win_tester = b; break; }
      }
    }
    terminate = win_tester >= 0;
  }

  /* --- terminal rewards and winner (environment.py:208-220) */
  int32_t winner = -1;
  if (terminate || t_global == sc->max_timesteps - 1) {
    reward[0] += win_ego ? sc->reward_win : (win_tester >= 0 ? -sc->reward_win : sc->reward_draw);
    for (int b = 1; b < m; ++b)
      reward[b] += win_ego ? -sc->reward_win : (win_tester < 0 ? sc->reward_draw : (win_tester == b ? sc->reward_win : sc->reward_draw));
    if (win_ego) winner = 0; else if (win_tester >= 0) winner = win_tester;
  }

  /* --- agents' process_feedback on the new state (simulation.py:86-87) */
  for (int b = 0; b < m; ++b) {
    int agent = o->bodies[b].agent;
    if (agent == CAV_AGENT_RANDOM_CONSTRAINED || agent == CAV_AGENT_PROXIMITY) {
      double st[4] = {ST(o, b, 0, e), ST(o, b, 1, e), ST(o, b, 2, e), ST(o, b, 3, e)};
      double ag[CAV_AGENT_WORDS];
      for (int k2 = 0; k2 < CAV_AGENT_WORDS; ++k2) ag[k2] = AG(o, b, k2, e);
      crossing_feedback(st, ag);
      for (int k2 = 0; k2 < CAV_AGENT_WORDS; ++k2) AG(o, b, k2, e) = ag[k2];
    }
  }

  /* --- episode accounting (simulation.py:69-97, reporting.py:227-243) */
  o->t_ep[e] += 1;
  o->winner[e] = winner;
  stats[CAV_STAT_ENV_STEPS] += 1;
  stats[CAV_STAT_BODY_STEPS] += m;
  stats[CAV_STAT_TANGENT] += tangent;
  if (terminate) o->done[e] = 1;
  else if (o->t_ep[e] >= sc->max_timesteps) o->done[e] = 2; /* cut off by Simulation.run, not `done` */
  if (o->done[e]) score_episode(o, e, stats);

  if (out->state) for (int b = 0; b < m; ++b) for (int k2 = 0; k2 < 4; ++k2) out->state[((int64_t)b * 4 + k2) * n + e] = ST(o, b, k2, e);
  if (out->reward) for (int b = 0; b < m; ++b) out->reward[(int64_t)b * n + e] = reward[b];
  if (out->done) out->done[e] = (uint8_t)terminate;
  if (out->winner) out->winner[e] = winner;
  if (out->tangent) out->tangent[e] = (uint8_t)tangent;
}

/* ------------------------------------------------------------------ public API (ctypes) */

int cav_oracle_create(const CavScenario* sc, int64_t n_envs, uint64_t seed, CavOracle** out) {
  if (!sc || !out || n_envs <= 0 || sc->n_bodies < 1 || sc->n_bodies > CAV_MAX_BODIES) return CAV_EINVAL;
  CavOracle* o = (CavOracle*)calloc(1, sizeof(CavOracle));
  if (!o) return CAV_ENOMEM;
  o->sc = *sc;
  int m = sc->n_bodies;
  o->bodies = (CavBody*)malloc(sizeof(CavBody) * m);
  memcpy(o->bodies, sc->bodies, sizeof(CavBody) * m);
  o->spawns = (CavSpawn*)malloc(sizeof(CavSpawn) * (sc->n_spawns > 0 ? sc->n_spawns : 1));
  if (sc->n_spawns > 0) memcpy(o->spawns, sc->spawns, sizeof(CavSpawn) * sc->n_spawns);
  o->sc.bodies = o->bodies;
  o->sc.spawns = o->spawns;
  o->n = n_envs;
  o->seed = seed;
  o->tau = 1e-7;
  o->threads = 1;
  o->state = (double*)calloc((size_t)m * 4 * n_envs, sizeof(double));
  o->action = (double*)calloc((size_t)m * 2 * n_envs, sizeof(double));
  o->agent = (double*)calloc((size_t)m * CAV_AGENT_WORDS * n_envs, sizeof(double));
  o->liveness = (int32_t*)calloc((size_t)m * n_envs, sizeof(int32_t));
  o->t_ep = (int32_t*)calloc(n_envs, sizeof(int32_t));
  o->episode = (int32_t*)calloc(n_envs, sizeof(int32_t));
  o->winner = (int32_t*)calloc(n_envs, sizeof(int32_t));
  o->done = (uint8_t*)calloc(n_envs, 1);
  o->err = (uint8_t*)calloc(n_envs, 1);
  for (int64_t e = 0; e < n_envs; ++e) { o->episode[e] = -1; reset_env(o, e, NULL); } /* constructor spawn: episode 0 */
  *out = o;
  return CAV_OK;
}

void cav_oracle_destroy(CavOracle* o) {
  if (!o) return;
  free(o->bodies); free(o->spawns); free(o->state); free(o->action); free(o->agent);
  free(o->liveness); free(o->t_ep); free(o->episode); free(o->winner); free(o->done); free(o->err);
  free(o);
}

void cav_oracle_set_shard(CavOracle* o, int64_t offset) {
  o->shard = offset;
  for (int64_t e = 0; e < o->n; ++e) { o->episode[e] = -1; reset_env(o, e, NULL); }
}
void cav_oracle_set_threads(CavOracle* o, int threads) { o->threads = threads > 0 ? threads : 1; }
void cav_oracle_set_tangent_tolerance(CavOracle* o, double tau) { o->tau = tau; }
void cav_oracle_set_global_timestep(CavOracle* o, int64_t t) { o->t_global = t; }
void cav_oracle_set_uniform_override(CavOracle* o, const double* u) { o->uni_override = u; }
void cav_oracle_set_spawn_override(CavOracle* o, const double* u) { o->spawn_override = u; }

void cav_oracle_reset(CavOracle* o, const uint8_t* mask, const double* init_state) {
  for (int64_t e = 0; e < o->n; ++e)
    if (!mask || mask[e]) reset_env(o, e, init_state);
}

/* Static partition of the env range over o->threads pthreads (envs are independent). */
typedef struct Job {
  CavOracle* o;
  int64_t lo, hi, t0;
  int n_steps, auto_reset;
  const double* actions;
  const StepOut* out;
  int64_t stats[CAV_N_STATS];
} Job;

static void* job_main(void* arg) {
  Job* j = (Job*)arg;
  CavOracle* o = j->o;
  for (int64_t e = j->lo; e < j->hi; ++e) {
    for (int s = 0; s < j->n_steps; ++s) {
      env_transition(o, e, j->actions, j->t0 + s, j->out, j->stats);
      if (j->auto_reset && o->done[e]) reset_env(o, e, NULL);
    }
  }
  return NULL;
}

static void run_jobs(CavOracle* o, const double* actions, const StepOut* out, int n_steps, int auto_reset) {
  int nt = o->threads;
  if ((int64_t)nt > o->n) nt = (int)o->n;
  Job* jobs = (Job*)calloc((size_t)nt, sizeof(Job));
  pthread_t* tid = (pthread_t*)calloc((size_t)nt, sizeof(pthread_t));
  for (int i = 0; i < nt; ++i) {
    jobs[i].o = o;
    jobs[i].lo = o->n * i / nt;
    jobs[i].hi = o->n * (i + 1) / nt;
    jobs[i].t0 = o->t_global;
    jobs[i].n_steps = n_steps;
    jobs[i].auto_reset = auto_reset;
    jobs[i].actions = actions;
    jobs[i].out = out;
    if (i > 0) pthread_create(&tid[i], NULL, job_main, &jobs[i]);
  }
  job_main(&jobs[0]);
  for (int i = 1; i < nt; ++i) pthread_join(tid[i], NULL);
  for (int i = 0; i < nt; ++i)
    for (int k = 0; k < CAV_N_STATS; ++k) o->stats[k] += jobs[i].stats[k];
  o->t_global += n_steps;
  free(jobs);
  free(tid);
}

void cav_oracle_step(CavOracle* o, const double* actions, double* state_out, double* reward_out,
                     uint8_t* done_out, int32_t* winner_out, uint8_t* tangent_out) {
  StepOut out = {state_out, reward_out, done_out, winner_out, tangent_out};
  run_jobs(o, actions, &out, 1, 0);
}

/* Simulation.run's loops (simulation.py:41-97) with the on-device-agent semantics of cavgym_rollout. */
void cav_oracle_rollout(CavOracle* o, int n_steps, int auto_reset) {
  StepOut out = {0, 0, 0, 0, 0};
  run_jobs(o, NULL, &out, n_steps, auto_reset);
}

/* Replayed joint actions [T][M][2][N] with trajectory outputs, as cavgym_replay. */
void cav_oracle_replay(CavOracle* o, int n_steps, const double* actions, double* state_traj, double* reward_traj,
                       uint8_t* done_traj, int32_t* winner_traj, uint8_t* tangent_traj) {
  int64_t m = o->sc.n_bodies, n = o->n;
  for (int s = 0; s < n_steps; ++s)
    cav_oracle_step(o, actions + (int64_t)s * m * 2 * n,
                    state_traj ? state_traj + (int64_t)s * m * 4 * n : NULL,
                    reward_traj ? reward_traj + (int64_t)s * m * n : NULL,
                    done_traj ? done_traj + (int64_t)s * n : NULL,
                    winner_traj ? winner_traj + (int64_t)s * n : NULL,
                    tangent_traj ? tangent_traj + (int64_t)s * n : NULL);
}

void cav_oracle_stats(CavOracle* o, int64_t* out10) {
  memcpy(out10, o->stats, sizeof(o->stats));
  int64_t errors = 0;
  for (int64_t e = 0; e < o->n; ++e) errors += o->err[e];
  out10[CAV_STAT_ERRORS] = errors;
}

double* cav_oracle_state_ptr(CavOracle* o) { return o->state; }
double* cav_oracle_action_ptr(CavOracle* o) { return o->action; }
double* cav_oracle_agent_state_ptr(CavOracle* o) { return o->agent; }
int32_t* cav_oracle_liveness_ptr(CavOracle* o) { return o->liveness; }
int32_t* cav_oracle_timestep_ptr(CavOracle* o) { return o->t_ep; }
int32_t* cav_oracle_winner_ptr(CavOracle* o) { return o->winner; }
uint8_t* cav_oracle_done_ptr(CavOracle* o) { return o->done; }
uint8_t* cav_oracle_error_ptr(CavOracle* o) { return o->err; }

/* CAVEnv.info (environment.py:106-117) for every env: body_polygons [M][8][N] (x of rear_left, front_left, front_right,
 * rear_right, then y) and road_angles [M][N], NaN where the reference returns None (the body's box intersects the major
 * road).  road angle = DynamicBody.line_anchor_relative_angle (bodies.py:206-212): orientation of the line from the body to
 * the closest point of the major road's centre line (geometry.py:412-421), minus the heading, normalised (:380-385). */
void cav_oracle_info(CavOracle* o, double* polygons_out, double* road_angle_out) {
  const CavScenario* sc = &o->sc;
  const int m = sc->n_bodies;
  const int64_t n = o->n;
  const double* cl = sc->centre_line;
  for (int64_t e = 0; e < n; ++e) {
    for (int b = 0; b < m; ++b) {
      const CavBody* body = &o->bodies[b];
      CavQuad box;
      if (body->kind == CAV_BODY_PELICAN) box = body->static_box;
      else {
        const CavBodyType* k = &sc->types[body->type_id];
        make_box(k->length, k->width, 0.5, ST(o, b, 3, e), ST(o, b, 0, e), ST(o, b, 1, e), &box);
      }
      if (polygons_out)
        for (int c = 0; c < 4; ++c) {
          polygons_out[((int64_t)b * 8 + c) * n + e] = box.x[c];
          polygons_out[((int64_t)b * 8 + 4 + c) * n + e] = box.y[c];
        }
      if (road_angle_out) {
        int unused = 0;
        double angle = NAN;
        if (!quad_intersects(&box, &sc->roads[0], o->tau, &unused)) {
          const double px = ST(o, b, 0, e), py = ST(o, b, 1, e);
          const double dx = cl[2] - cl[0], dy = cl[3] - cl[1];
          const double denominator = (dx * dx) + (dy * dy);
          const double a = (dy * (py - cl[1]) + dx * (px - cl[0])) / denominator;
          const double cx = cl[0] + a * dx, cy = cl[1] + a * dy;
          double radians = atan2(cy - py, cx - px) - ST(o, b, 3, e);
          while (radians <= -M_PI) radians += 2 * M_PI;
          while (radians > M_PI) radians -= 2 * M_PI;
          angle = radians == 0 ? radians + 0.0 : radians;
        }
        road_angle_out[(int64_t)b * n + e] = angle;
      }
    }
  }
}

/* ---- single-shot helpers for unit parity tests (geometry known-answer vectors) */

void cav_oracle_make_box(double length, double width, double theta, double px, double py, CavQuad* out) {
  make_box(length, width, 0.5, theta, px, py, out);
}
int cav_oracle_intersects(const CavQuad* a, const CavQuad* b) { int t = 0; return quad_intersects(a, b, 1e-7, &t); }
int cav_oracle_contains(const CavQuad* a, const CavQuad* b) { int t = 0; return quad_contains(a, b, 1e-7, &t); }
double cav_oracle_percentage_intersects(const CavQuad* a, const CavQuad* b) { int t = 0; return percentage_intersects(a, b, 1e-7, &t); }
int cav_oracle_stopping_zones(const CavBodyType* k, const double* st, double steer, CavQuad* braking, CavQuad* reaction) {
  return stopping_zones(k, st, steer, braking, reaction);
}
void cav_oracle_dynamic_body_step(const CavBodyType* k, double* st, double throttle, double steer, double dt) {
  dynamic_body_step(k, st, throttle, steer, dt);
}
double cav_oracle_make_steering_action(const CavBodyType* k, const double* st, double dt, double target) {
  return make_steering_action(k, st, dt, target);
}
