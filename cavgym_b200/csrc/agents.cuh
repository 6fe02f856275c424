// agents.cuh — kinematics, on-device agents and the counter-based RNG.
//
//  body_step              DynamicBody.step                    library/bodies.py:214-275
//  make_steering_action   inverse bicycle arc                 examples/agents/dynamic_body.py:33-49
//  choose_crossing_action CrossingAgent state machine         examples/agents/pedestrian.py:50-69
//  crossing_feedback      CrossingAgent.process_feedback      examples/agents/pedestrian.py:36-48
//  spawn_body             SpawnPedestrian.spawn               library/bodies.py:302-312
//  philox4x32_10          replaces the shared MT19937 RandomState (config.py:275)
#pragma once
#include "geometry.cuh"

namespace cav {

// The fp64 libm routines are ~100-instruction sequences: one out-of-line copy each keeps the kernels' hot path small.
static __device__ __noinline__ void sincos_(double a, double* s, double* c) { sincos(a, s, c); }
static __device__ __noinline__ void sincos_(float a, float* s, float* c) { sincosf(a, s, c); }
static __device__ __noinline__ double sin_(double a) { return sin(a); }
static __device__ __noinline__ double cos_(double a) { return cos(a); }
static __device__ __noinline__ double tan_(double a) { return tan(a); }
static __device__ __noinline__ float tan_(float a) { return tanf(a); }
static __device__ __noinline__ double atan2_(double y, double x) { return atan2(y, x); }
static __device__ __noinline__ float atan2_(float y, float x) { return atan2f(y, x); }
static __device__ __noinline__ double atan_(double a) { return atan(a); }
static __device__ __noinline__ float atan_(float a) { return atanf(a); }
__device__ __forceinline__ bool isnan_(double a) { return isnan(a); }
__device__ __forceinline__ bool isnan_(float a) { return isnan(a); }
template <typename R> __device__ __forceinline__ R nan_() { return R(NAN); }

// ---------------------------------------------------------------- trigonometry for the kinematics
// sin x and cos x - 1 for |x| <= pi/4 (fdlibm kernel polynomials, < 1 ulp), explicit FMAs.  cos x - 1 is returned
// instead of cos x because the turn below multiplies it by the turn radius (up to 1e14 px for a near-zero steering
// angle): forming cos x first would lose every digit of the product.
// Coefficients live in the constant bank (one LDCU.128 per pair) instead of being materialised as immediates.
static __constant__ double kSinPoly[6] = {1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
                                          -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01};
static __constant__ double kCosPoly[6] = {-1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
                                          2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};

__device__ __forceinline__ void sin_cosm1_poly(double x, double& sn, double& cm1) {
  const double z = x * x;
  double ps = fma(z, kSinPoly[0], kSinPoly[1]);
  ps = fma(z, ps, kSinPoly[2]);
  ps = fma(z, ps, kSinPoly[3]);
  ps = fma(z, ps, kSinPoly[4]);
  ps = fma(z, ps, kSinPoly[5]);
  sn = fma(x * z, ps, x);
  double pc = fma(z, kCosPoly[0], kCosPoly[1]);
  pc = fma(z, pc, kCosPoly[2]);
  pc = fma(z, pc, kCosPoly[3]);
  pc = fma(z, pc, kCosPoly[4]);
  pc = fma(z, pc, kCosPoly[5]);
  cm1 = fma(z * z, pc, -0.5 * z);
}

static __device__ __noinline__ void sin_cosm1_wide(double x, double* sn, double* cm1) {  // |x| > pi/4: never in practice
  double h, unused;
  sincos(x, sn, &unused);
  h = sin(0.5 * x);
  *cm1 = -2.0 * h * h;
}

__device__ __forceinline__ void sin_cosm1(double x, double& sn, double& cm1) {
  if (fabs(x) <= 0.78539816339744830962) sin_cosm1_poly(x, sn, cm1);
  else sin_cosm1_wide(x, &sn, &cm1);
}
// wheelbase / tan(steer) and 1 / turn radius of the body centre for an arbitrary steering angle
// (bodies.py:248 tan, :256-262 radius).  The radius is |centre - ICR| = sqrt((wb/2)^2 + (wb/tan)^2) whatever the
// heading, so no square root of position differences is needed.
__device__ __forceinline__ void steer_geometry(double wheelbase, double half_wb, double steer, double& kk, double& inv_r) {
  double sn, cm1, cs;
  const double a = fabs(steer);
  if (a <= 0.78539816339744830962) {
    sin_cosm1_poly(steer, sn, cm1);
    cs = 1.0 + cm1;
  } else if (a <= 2.3561944901923448) {  // one quadrant of Cody-Waite reduction: steer = +-pi/2 + r
    const double k = steer < 0.0 ? -1.0 : 1.0;
    const double r = fma(-k, 6.123233995736766036e-17, fma(-k, 1.57079632679489655800e+00, steer));
    double sr, cr1;
    sin_cosm1_poly(r, sr, cr1);
    sn = k * (1.0 + cr1);   // sin(k pi/2 + r) = k cos r
    cs = -k * sr;           // cos(k pi/2 + r) = -k sin r
  } else {
    sn = sin_(steer);
    cs = cos_(steer);
  }
  kk = wheelbase * fast_div(cs, sn);
  inv_r = rsqrt(fma(kk, kk, half_wb * half_wb));
}
// atan2(sin(a), cos(a)) (bodies.py:274): a wrapped into [-pi, pi].  One conditional +-2 pi instead of three libm calls;
// the result is the value the reference's composition approximates to a few ulp.
template <typename R>
__device__ __forceinline__ R wrap_angle(R a) {
  const R pi = R(3.14159265358979323846), two_pi = R(6.28318530717958647692);
  if (rabs(a) > R(3) * pi) {  // bodies created with headings beyond +-2 pi: general formula
    R sa, ca;
    sincos_(a, &sa, &ca);
    return atan2_(sa, ca);
  }
  if (a > pi) return a - two_pi;
  if (a < -pi) return a + two_pi;
  return a;
}

// ---------------------------------------------------------------- kinematics
// cos/sin of a heading: exact for 0, looked up (host libm values) for the headings bodies are created with,
// device sincos otherwise.
template <typename R>
__device__ __forceinline__ void heading_cs(const DevScenario<R>& sc, R th, R& c, R& s) {
  if (th == R(0)) { c = R(1); s = th; return; }  // cos(+-0) = 1, sin(+-0) = +-0
  bool hit = false;
#pragma unroll
  for (int i = 0; i < CAV_HEADING_CACHE; ++i) {
    if (i < sc.hc_n && th == sc.hc_theta[i]) { c = sc.hc_cos[i]; s = sc.hc_sin[i]; hit = true; }
  }
  if (!hit) sincos_(th, &s, &c);
}

// DynamicBody.step (bodies.py:214-275).  st = x, y, v, theta in/out.  c, s: cos/sin of the CURRENT heading on entry
// (cached in EnvBuffers::cs), of the NEW heading on return; `snapped` is the steering angle after the 1e-13 snap
// (bodies.py:217-218).  Returns true if the heading changed.
//
// The turning branch (bodies.py:241-275) rotates the body centre about the instantaneous centre of rotation
// ICR = rear axle + (wb / tan steer) * left normal by phi = +-distance / |centre - ICR|.  The reference forms the ICR in
// world coordinates and rotates (centre - ICR); here the same rotation is applied to the body-frame offset
//   d = centre - ICR = (wb/2) (c, s) + kk (s, -c),   kk = wb / tan steer,   |d|^2 = (wb/2)^2 + kk^2,
//   centre' = centre + d (cos phi - 1) + d_perp sin phi,
// which is the same map without the cancellation of two ~kk-sized world coordinates (that cancellation costs the
// reference ~1e-11 px per step at small steering angles in fp64 and would cost whole pixels in fp32), and without
// sqrt, division, or atan2: the new heading's cos/sin follow by the angle-addition formulas.
// The turning branch (bodies.py:241-275): d = v dt with the OLD velocity; st[2] has already been updated.
template <typename R>
__device__ __forceinline__ void body_turn(const DevType<R>& k, R st[4], R steer, R d, R& c, R& s) {
  // The turn increment is evaluated in double in BOTH modes: a steering angle is held for many steps, so a few-ulp
  // float error in wb / tan(steer) would be a systematic heading drift (1e-5 rad over an episode), not a random one.
  double kk, inv_r;
  if (steer == k.smax) { kk = k.kk_smax; inv_r = k.inv_r_smax; }
  else if (steer == k.smin) { kk = k.kk_smin; inv_r = k.inv_r_smin; }
  else { steer_geometry((double)k.wheelbase, (double)k.half_wb, (double)steer, kk, inv_r); }
  const double q = (double)d * inv_r;
  const double phi = steer < R(0) ? -q : q;
  double sn_, cm1_;
  sin_cosm1(phi, sn_, cm1_);
  const R sn = (R)sn_, cm1 = (R)cm1_;
  const R dx = k.half_wb * c + (R)kk * s, dy = k.half_wb * s - (R)kk * c;
  st[0] = st[0] + (dx * cm1 - dy * sn);
  st[1] = st[1] + (dx * sn + dy * cm1);
  const R c2 = c + (c * cm1 - s * sn), s2 = s + (s * cm1 + c * sn);
  c = c2; s = s2;
  st[3] = wrap_angle(st[3] + (R)phi);
}
// One out-of-line copy for kernels whose hot loop must stay small (kernels_dense.cuh): same arithmetic.
template <typename R>
__device__ __noinline__ void body_turn_outlined(const DevType<R>& k, R st[4], R steer, R d, R* c, R* s) {
  body_turn(k, st, steer, d, *c, *s);
}

template <typename R, bool OUTLINE_TURN = false>
__device__ __forceinline__ bool body_step(const DevType<R>& k, R st[4], R throttle, R steer, R dt, R& c, R& s, R& snapped) {
  if (rabs(steer) < R(0.0000000000001)) steer = R(0);
  snapped = steer;
  const R v = st[2];
  const R d = v * dt;
  st[2] = rmax(k.vmin, rmin(k.vmax, v + (throttle * dt)));
  if (steer == R(0)) {
    st[0] = st[0] + d * c;
    st[1] = st[1] + d * s;
    return false;
  }
  if (OUTLINE_TURN) body_turn_outlined(k, st, steer, d, &c, &s);
  else body_turn(k, st, steer, d, c, s);
  return true;
}

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0,1) from two words (numpy legacy random_sample construction).
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0;
}

enum { KIND_AGENT0 = 0, KIND_AGENT1 = 1, KIND_SPAWN0 = 2, KIND_SPAWN1 = 3, KIND_SPAWN2 = 4 };

// Stream: key = seed; counter = (env_lo, env_hi | kind<<8 | body<<16, episode, timestep).
__device__ __forceinline__ void draw_block(uint64_t seed, uint64_t global_env, int body, int kind, uint32_t episode, uint32_t t,
                                           double u[2]) {
  uint32_t w[4];
  philox4x32_10((uint32_t)global_env, (uint32_t)((global_env >> 32) & 0xFFu) | ((uint32_t)kind << 8) | ((uint32_t)body << 16),
                episode, t, (uint32_t)seed, (uint32_t)(seed >> 32), w);
  u[0] = u53(w[0], w[1]);
  u[1] = u53(w[2], w[3]);
}

// ---------------------------------------------------------------- spawn
template <typename R>
__device__ __forceinline__ R triangle_area(R rx, R ry, R flx, R fly, R frx, R fry) {  // geometry.py:352-354
  return rabs((rx * (fly - fry) + flx * (fry - ry) + frx * (ry - fly)) / R(2));
}

// numpy legacy choice(p=...) = searchsorted(cumsum(p)/last, u, 'right'); the area arithmetic is done in
// double in both modes so fp32 and fp64 engines pick the same box/triangle for the same draws.
template <typename R>
__device__ __noinline__ void spawn_body(const DevSpawn<R>& sp, const double u[5], R st[4]) {
  double areas[CAV_MAX_SPAWN_BOXES], larea[CAV_MAX_SPAWN_BOXES], total = 0.0;
  for (int i = 0; i < sp.n_boxes; ++i) {
    const Quad<R>& q = sp.boxes[i];
    larea[i] = triangle_area<double>(q.x[1], q.y[1], q.x[2], q.y[2], q.x[0], q.y[0]);
    const double r = triangle_area<double>(q.x[3], q.y[3], q.x[0], q.y[0], q.x[2], q.y[2]);
    areas[i] = 0.0 + larea[i] + r;
    total += areas[i];
  }
  double cdf[CAV_MAX_SPAWN_BOXES], acc = 0.0;
  for (int i = 0; i < sp.n_boxes; ++i) { acc += areas[i] / total; cdf[i] = acc; }
  int box = 0;
  for (int i = 0; i < sp.n_boxes; ++i) box += (cdf[i] / cdf[sp.n_boxes - 1] <= u[0]);
  if (box >= sp.n_boxes) box = sp.n_boxes - 1;
  const Quad<R>& q = sp.boxes[box];
  const double la = larea[box];
  const double sa = la + triangle_area<double>(q.x[3], q.y[3], q.x[0], q.y[0], q.x[2], q.y[2]);
  const double f = la / sa, c0 = f, c1 = f + (1 - f);
  int tri = ((c0 / c1) <= u[1]) + ((c1 / c1) <= u[1]);
  if (tri > 1) tri = 1;
  R rx, ry, flx, fly, frx, fry;
  if (tri == 0) { rx = q.x[1]; ry = q.y[1]; flx = q.x[2]; fly = q.y[2]; frx = q.x[0]; fry = q.y[0]; }
  else          { rx = q.x[3]; ry = q.y[3]; flx = q.x[0]; fly = q.y[0]; frx = q.x[2]; fry = q.y[2]; }
  double a = u[2], b = u[3];
  if (a + b > 1) { a = 1 - a; b = 1 - b; }
  st[0] = R((double(rx) + (double(flx) - double(rx)) * a) + (double(frx) - double(rx)) * b);
  st[1] = R((double(ry) + (double(fly) - double(ry)) * a) + (double(fry) - double(ry)) * b);
  st[2] = sp.velocity;
  int oi = (int)floor(u[4] * sp.n_orient);
  if (oi >= sp.n_orient) oi = sp.n_orient - 1;
  st[3] = sp.orient[oi];
}

// ---------------------------------------------------------------- crossing agents
template <typename R>
__device__ __noinline__ R steering_towards(const DevType<R>& k, R v, R theta, R dt, R target) {
  // atan2(sin(t - theta), cos(t - theta)) (dynamic_body.py:36) is t - theta wrapped into [-pi, pi]: one conditional +-2 pi
  // instead of three libm calls, as in body_step; sqrt(wb^2 (1 + 4 / tan^2(limit))) only ever sees the two steering limits
  // and comes from the host (DevType::max_turn_*).
  const R tta = wrap_angle(target - theta);
  const R csa = tta < R(0) ? k.smin : k.smax;
  const R wb = k.wheelbase;
  // At full lock (the turn wanted exceeds what one step allows: tta / mta > 1 with mta = +-2 dt v / max_turn, the largest
  // turn of one step) the inverse below is atan(|tan(limit)|) = |limit| exactly in real arithmetic — the reference lands within
  // an ulp of it and then clamps — so the limit itself is returned: no atan, sqrt or division on 13 of the 14 steps of a
  // crossing turn, and body_step then takes its cached full-lock constants.
  // The test itself is two IEEE divisions (~90 instructions at one or two active lanes per warp); the quotient of the two
  // rounded divisions is within 4 ulp of |tta| max_turn / (2 dt v), so unless that ratio is within 16 ulp of 1 the products
  // decide it — and when it is, the divisions do.
  const R max_turn = csa < R(0) ? k.max_turn_smin : k.max_turn_smax;
  const R reach = R(2) * dt * v;
  bool full_lock;
  const R want = rabs(tta) * max_turn, band = reach * (sizeof(R) == 8 ? R(4e-15) : R(2e-6));
  if (CAV_LIKELY(v > R(0) && max_turn > R(0) && rabs(want - reach) > band)) {
    full_lock = want > reach;
  } else {
    const R mta = (csa < R(0) ? R(-2) : R(2)) * dt * v / max_turn;
    full_lock = tta / mta > R(1);
  }
  if (full_lock) return csa;
  const R ta = tta;
  const R steer = atan_(R(2) * wb * rsqrt_((ta * ta) / (R(4) * (v * v) * (dt * dt) - (wb * wb) * (ta * ta))));
  return ta < R(0) ? -steer : steer;
}

template <typename R>
__device__ __forceinline__ R make_steering_action(const DevType<R>& k, const R st[4], R dt, R target) {
  R steer = R(0);
  if (!(st[2] == R(0) || isnan_(target))) steer = steering_towards(k, st[2], st[3], dt, target);
  return rmin(k.smax, rmax(k.smin, steer));
}

template <typename R>
__device__ __forceinline__ R point_distance(R sx, R sy, R ox, R oy) {  // Point.distance geometry.py:18-19
  const R dy = oy - sy, dx = ox - sx;
  return rsqrt_((dy * dy) + (dx * dx));
}

// ag = initial_distance, waypoint x, waypoint y, target_orientation, prior_orientation (NaN = None).
template <typename R>
__device__ __noinline__ void start_crossing(const R cl[4], const R st[4], R ag[CAV_AGENT_WORDS]) {
  const R dx = cl[2] - cl[0], dy = cl[3] - cl[1];
  const R denominator = (dx * dx) + (dy * dy);
  const R a = (dy * (st[1] - cl[1]) + dx * (st[0] - cl[0])) / denominator;  // Line.closest_point_from geometry.py:412-416
  const R cx = cl[0] + a * dx, cy = cl[1] + a * dy;
  const R rel = atan2_(cy - st[1], cx - st[0]);
  if (isnan_(ag[0])) ag[0] = point_distance(st[0], st[1], cx, cy);
  R sr, cr;
  sincos_(rel, &sr, &cr);
  ag[1] = cx + ag[0] * cr;
  ag[2] = cy + ag[0] * sr;
  ag[3] = atan2_(ag[2] - st[1], ag[1] - st[0]);
  ag[4] = st[3];
}

template <typename R>
__device__ __forceinline__ R choose_crossing_action(const DevScenario<R>& sc, const DevType<R>& k, const R st[4],
                                                    R ag[CAV_AGENT_WORDS], bool condition, bool& dirty) {
  if (isnan_(ag[1]) && isnan_(ag[3]) && condition) {
    start_crossing(sc.cl, st, ag);
    dirty = true;
  }
  return make_steering_action(k, st, sc.dt, ag[3]);
}

template <typename R>
__device__ __forceinline__ void crossing_feedback(const DevScenario<R>& sc, const R st[4], R ag[CAV_AGENT_WORDS], bool& dirty) {
  if (!isnan_(ag[1])) {
    if (point_distance(st[0], st[1], ag[1], ag[2]) < R(1)) {
      ag[1] = nan_<R>(); ag[2] = nan_<R>(); ag[3] = ag[4]; ag[4] = nan_<R>();
      dirty = true;
    }
  }
  if (!isnan_(ag[3])) {
    // |atan2(sin d, cos d)| < TARGET_ERROR (pedestrian.py:44-47) with d = target - theta wrapped into [-pi, pi]
    if (rabs(wrap_angle(ag[3] - st[3])) < sc.target_err) { ag[3] = nan_<R>(); dirty = true; }
  }
}

}  // namespace cav
