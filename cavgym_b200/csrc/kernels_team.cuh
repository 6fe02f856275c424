// kernels_team.cuh — cavgym_rollout for HETEROGENEOUS scenarios of three to eight bodies (crossroads, bus stop, pelican
// crossing; BASELINE config C3): a TEAM of M warps steps 32 environments, warp b owns body b of each of them.
//
// Why: thread-per-env (kernels_small.cuh) has to walk the bodies of a heterogeneous scenario in a rolled loop — unrolled, a
// step is 60 KB of code for a 32 KB instruction cache — so its per-body arrays are indexed at run time and live in local
// memory: 1.2 KB per thread, more than L1 holds at 16 warps per SM; ncu shows the kernel waiting on that memory, on
// instruction fetch and on the barriers that keep its warps together (profiles/r2_rollout_kernel_busstop_ncu.txt).  Here
//   * a thread owns ONE body of ONE env: x, y, v, theta, cos, sin, the held action, the crossing-agent words and the
//     liveness count stay in registers for the whole launch — no local memory, no run-time indexing;
//   * the lanes of a warp hold the SAME body of 32 envs: the body's kind, class flags and agent kind are warp-uniform, so
//     the per-body code runs without divergence between kinds and a rare path (an arbitrary steering angle, a kerb, a
//     redraw) is paid once per 32 envs — the amortisation thread-per-env has, which one-body-per-lane-of-one-env loses
//     (tried first: 4 or 8 lanes per env, 1.8x SLOWER than thread-per-env);
//   * bodies meet through shared memory, [word][body][env lane] (conflict-free): after DynamicBody.step every warp posts its
//     body's position, AABB half extents and heading; the all-pairs test of environment.py:156-177 is met by ROTATION (in
//     round r warp b takes the pair (b, b + r mod M)), static collidables / off-road / stopping zones / road share are
//     evaluated by the warp of the body concerned, and each warp posts one word of decision bits per env;
//   * every warp then folds the M words of its env in the reference's cascade order (finish line, all-pairs, off-road, ego
//     box / braking zone, reaction zone — environment.py:148-206), so all warps agree on termination, winner and the
//     near-tangent flag without a further exchange.  Later stages are evaluated speculatively and GATED when folding:
//     a stage the reference would not have reached contributes nothing (its near-tangent bits included).
// Two block-wide barriers per step (four with proximity agents, which look at the ego before it moves).
// MEASURED: no faster than the thread-per-env kernel (bus stop 52 ms against 44.5 ms per 100 steps of 1 M envs, pelican
// crossing 31.0 against 31.6; the local-memory stall is gone, the waits at the two barriers and the instruction-fetch stall
// take its place — DESIGN 4.6), so cavgym_rollout uses it only on request (cavgym_set_rollout_path).  It stays as an independent
// third implementation of the transition that must agree bit for bit with the other two.
// The arithmetic and its order are those of transition_body.inc (the same device functions; a pair is evaluated with the
// lower body index first), so results — state, agent state, liveness, counters, near-tangent flags — are bitwise those of
// the thread-per-env and warp-per-env kernels (tests/test_gpu_team.py, tests/test_gpu_dense.py).
// Election agents (arbitration across bodies, election.py) stay with the thread-per-env kernel.
#pragma once
#include "transition.cuh"

namespace cav {

// Shape.percentage_intersects(body box, road r): road_share of transition.cuh on an explicit pose — same cases, same
// order, same arithmetic (GENERIC form).
template <typename R>
__device__ __forceinline__ R team_road_share(const DevScenario<R>& sc, const DevType<R>& k, R px, R py, R th, R c, R s, int r, R ex,
                                             R ey, R tau, bool& near) {
  const Aabb<R> rd = sc.road_bb[r];
  const R m0 = (px - ex) - rd.x0, m1 = rd.x1 - (px + ex), m2 = (py - ey) - rd.y0, m3 = rd.y1 - (py + ey);
  const R mx = rmin(m0, m1), my = rmin(m2, m3);
  if (mx < -((ex + ex) + tau) || my < -((ey + ey) + tau)) return R(0);
  const Pose<R> pose = {px, py, th, c, s, k.length, k.width};
  if (sc.road_axis[r]) {
    const bool x_edge = mx < my;
    const R lo = x_edge ? mx : my, hi = x_edge ? my : mx;
    if (lo >= tau) return R(1);
    const R opposite = x_edge ? rmax(m0, m1) : rmax(m2, m3);
    if (lo > -tau && hi >= tau && opposite >= tau) return kerb_touch_share(c, s, k.hl, k.hw, lo, x_edge, ex, ey, near);
    if (lo <= -tau && hi >= tau && opposite >= tau) {
      const R ac = rabs(c), as = rabs(s);
      const R p = kerb_share(lo + (x_edge ? ex : ey), (x_edge ? ac : as) * k.hl, (x_edge ? as : ac) * k.hw);
      if (rabs(p - R(0.5)) < tau) near = true;
      return p;
    }
    if (mx < tau && my < tau && rmax(m0, m1) >= tau && rmax(m2, m3) >= tau) {
      const bool low_x = m0 < m1, low_y = m2 < m3;
      const R p = corner_share_closed(pose, mx + ex, my + ey, low_x ? R(-1) : R(1), low_x ? -rd.x0 : rd.x1, low_y ? R(-1) : R(1),
                                      low_y ? -rd.y0 : rd.y1, tau);
      if (rabs(p - R(0.5)) < tau) near = true;
      return p;
    }
  }
  const Share<R> share = road_share_general(pose, &sc.quads[r], tau);
  near |= (share.tangent != 0) || (rabs(share.value - R(0.5)) < tau);
  return share.value;
}

// One body of one environment, in registers.
template <typename R>
struct TeamLane {
  R s[4];                   // x, y, v, theta (PelicanCrossing: light state in [0])
  R c, sn;                  // cos, sin of the heading
  R held[2];                // the action taken last (RandomAgent holds it)
  R ag[CAV_AGENT_WORDS];    // crossing-agent state
  int32_t live;
  bool live_dirty, cs_dirty, ag_dirty;
};

// Env-wide values, identical in the G lanes of an env.
struct TeamEnv {
  int32_t t_ep, episode, winner;
  uint8_t done;
};

template <typename R>
__device__ __forceinline__ bool team_uses_agent_state(const DevBody<R>& body) {
  return body.agent == CAV_AGENT_RANDOM_CONSTRAINED || body.agent == CAV_AGENT_PROXIMITY;
}

// CAVEnv.reset for this lane's body (reset_env of transition.cuh, one body).
template <typename R>
__device__ __forceinline__ void team_reset_lane(const DevScenario<R>& sc, const EnvBuffers<R>& buf, const DevBody<R>& body, int b,
                                                int64_t e, uint32_t episode, TeamLane<R>& ln) {
  R st[4];
#pragma unroll
  for (int w = 0; w < 4; ++w) st[w] = body.init[w];
  if ((body.flags & CAV_FLAG_SPAWN) && body.spawn_id >= 0) {
    double u[5];
    if (buf.spawn_override) {
#pragma unroll
      for (int w = 0; w < 5; ++w) u[w] = buf.spawn_override[((int64_t)b * 5 + w) * buf.n + e];
    } else {
      double w2[2];
      const uint64_t g = (uint64_t)(buf.shard + e);
      draw_block(buf.seed, g, b, KIND_SPAWN0, episode, 0u, w2); u[0] = w2[0]; u[1] = w2[1];
      draw_block(buf.seed, g, b, KIND_SPAWN1, episode, 0u, w2); u[2] = w2[0]; u[3] = w2[1];
      draw_block(buf.seed, g, b, KIND_SPAWN2, episode, 0u, w2); u[4] = w2[0];
    }
    spawn_body(buf.spawns[body.spawn_id], u, st);
  }
#pragma unroll
  for (int w = 0; w < 4; ++w) ln.s[w] = st[w];
  ln.held[0] = R(0); ln.held[1] = R(0);
  ln.c = R(1); ln.sn = R(0);
  if (body.kind == CAV_BODY_DYNAMIC) heading_cs(sc, st[3], ln.c, ln.sn);
#pragma unroll
  for (int w = 0; w < CAV_AGENT_WORDS; ++w) ln.ag[w] = nan_<R>();
  ln.live = 0;
  ln.live_dirty = true; ln.cs_dirty = true; ln.ag_dirty = true;
}


// decision bits a warp posts for its body and env (folded by every warp in cascade order)
enum { TB_HIT_ALL = 1, TB_TAN_ALL = 2, TB_OFF = 4, TB_TAN_OFF = 8, TB_ZONE_HIT = 16, TB_ZONE_WIN = 32, TB_NEAR_REACTION = 64,
       TB_TAN_ZONE = 128, TB_TAN_SHARE = 256, TB_INVALID = 512 };
// exchange words
enum { TX_X = 0, TX_Y, TX_EX, TX_EY, TX_C, TX_S, TX_V, TX_STEER, TX_WORDS };

// Resident threads per SM the register budget is set for (launch bounds): a thread holds one body, not a whole env, so more
// warps fit than in the thread-per-env kernels, and the barrier waits of one team are covered by the other teams of the SM.
#ifndef CAV_TEAM_THREADS_PER_SM
#define CAV_TEAM_THREADS_PER_SM 512
#endif
template <int M> __host__ __device__ constexpr int team_min_blocks() {
  return CAV_TEAM_THREADS_PER_SM / (32 * M) > 0 ? CAV_TEAM_THREADS_PER_SM / (32 * M) : 1;
}

template <typename R, int M>
__global__ void __launch_bounds__(32 * M, (team_min_blocks<M>())) team_rollout_kernel(const __grid_constant__ DevScenario<R> sc,
                                                                                     const __grid_constant__ EnvBuffers<R> buf,
                                                                                     int64_t t_global, int n_steps, int auto_reset) {
  __shared__ R xch[TX_WORDS][M][32];
  __shared__ uint32_t flg[M][32];
  __shared__ int32_t liv[M][32];
  const int lane = threadIdx.x & 31;
  const int b = threadIdx.x >> 5;                    // this warp's body
  const int64_t e_raw = buf.lo + (int64_t)blockIdx.x * 32 + lane;
  const bool in_range = e_raw < buf.hi;              // lanes without an env idle through the barriers
  const int64_t e = in_range ? e_raw : buf.hi - 1;
  const int64_t n = buf.n;
  const DevBody<R>& body = sc.bodies[b];
  const DevType<R>& k = body.k;
  const DevType<R>& k0 = sc.bodies[0].k;
  const bool pelican = body.kind == CAV_BODY_PELICAN;
  const bool dynamic = !pelican;
  const bool pedestrian = (body.flags & CAV_FLAG_PEDESTRIAN) != 0;
  const int agent = body.agent;
  const bool crossing_agent = team_uses_agent_state(body);
  bool any_proximity = false;
  uint32_t dynamic_bodies = 0u;   // bit p: body p takes part in the all-pairs test (not a PelicanCrossing)
#pragma unroll
  for (int p = 0; p < M; ++p) {
    any_proximity |= sc.bodies[p].agent == CAV_AGENT_PROXIMITY;
    if (sc.bodies[p].kind != CAV_BODY_PELICAN) dynamic_bodies |= 1u << p;
  }
  const R tau = sc.tau, dt = sc.dt, W = sc.W;

  // ---- load (load_env of kernels_small.cuh, this warp's body)
  TeamLane<R> ln;
  TeamEnv env;
  env.done = in_range ? buf.done[e] : (uint8_t)1;
  env.t_ep = buf.t_ep[e];
  env.winner = env.done ? buf.winner[e] : -1;
  env.episode = buf.episode[e];
  ln.live = 0; ln.live_dirty = false; ln.cs_dirty = false; ln.ag_dirty = false;
  ln.held[0] = R(0); ln.held[1] = R(0);
  ln.c = R(1); ln.sn = R(0);
#pragma unroll
  for (int w = 0; w < CAV_AGENT_WORDS; ++w) ln.ag[w] = nan_<R>();
#pragma unroll
  for (int w = 0; w < 4; ++w) ln.s[w] = buf.state[((int64_t)b * 4 + w) * n + e];
  if (b > 0) ln.live = buf.liveness[(int64_t)b * n + e];
  if (dynamic) { ln.c = buf.cs[((int64_t)b * 2 + 0) * n + e]; ln.sn = buf.cs[((int64_t)b * 2 + 1) * n + e]; }
  if (agent == CAV_AGENT_RANDOM) { ln.held[0] = buf.action[((int64_t)b * 2 + 0) * n + e]; ln.held[1] = buf.action[((int64_t)b * 2 + 1) * n + e]; }
  if (crossing_agent) {
#pragma unroll
    for (int w = 0; w < CAV_AGENT_WORDS; ++w) ln.ag[w] = buf.agent[((int64_t)b * CAV_AGENT_WORDS + w) * n + e];
  }
  bool was_reset = false;

  for (int t = 0; t < n_steps; ++t) {
    if (CAV_UNLIKELY(env.done != 0) && auto_reset && in_range) {
      env.episode += 1;
      team_reset_lane(sc, buf, body, b, e, (uint32_t)env.episode, ln);
      env.t_ep = 0; env.done = 0; env.winner = -1;
      was_reset = true;
    }
    const bool stepping = env.done == 0;   // a finished env without auto-reset (and a lane without an env) idles through the barriers
    R ego_pre_x = R(0), ego_pre_y = R(0);
    if (any_proximity) {   // (launch-uniform) ProximityAgent looks at the ego BEFORE it moves (pedestrian.py:78-91)
      if (b == 0) { xch[TX_X][0][lane] = ln.s[0]; xch[TX_Y][0][lane] = ln.s[1]; }
      __syncthreads();
      ego_pre_x = xch[TX_X][0][lane]; ego_pre_y = xch[TX_Y][0][lane];
      __syncthreads();   // before warp 0 posts its moved body below
    }
    uint32_t bits = 0u;
    R ex = R(0), ey = R(0), snapped = R(0);
    if (stepping) {
      // ---- this body's action from the pre-step state (transition_body.inc)
      R a0 = ln.held[0], a1 = ln.held[1];
      double u[CAV_DRAWS] = {0.0, 0.0, 0.0};
      if (agent == CAV_AGENT_RANDOM || agent == CAV_AGENT_RANDOM_CONSTRAINED) {
        if (CAV_UNLIKELY(buf.uni_override != nullptr)) {
#pragma unroll
          for (int w = 0; w < CAV_DRAWS; ++w) u[w] = buf.uni_override[((int64_t)b * CAV_DRAWS + w) * n + e];
        } else {
          draw_block(buf.seed, (uint64_t)(buf.shard + e), b, KIND_AGENT0, (uint32_t)env.episode, (uint32_t)env.t_ep, u);
        }
      }
      if (agent == CAV_AGENT_NOOP) {
        a0 = R(0); a1 = R(0);
      } else if (agent == CAV_AGENT_RANDOM) {   // RandomAgent.choose_action (template.py:52-56)
        if (u[0] < body.epsilon) {
          if (pelican) {
            a0 = R(floor(u[1] * 4.0)); if (a0 > R(3)) a0 = R(3);
            a1 = R(0);
          } else {
            if (!buf.uni_override) {
              double w2[2];
              draw_block(buf.seed, (uint64_t)(buf.shard + e), b, KIND_AGENT1, (uint32_t)env.episode, (uint32_t)env.t_ep, w2);
              u[2] = w2[0];
            }
            a0 = R(double(k.amin) + (double(k.amax) - double(k.amin)) * u[1]);
            a1 = R(double(k.smin) + (double(k.smax) - double(k.smin)) * u[2]);
          }
        }
      } else if (agent == CAV_AGENT_RANDOM_CONSTRAINED) {   // pedestrian.py:72-75
        a0 = R(0);
        bool dirty = false;
        a1 = choose_crossing_action(sc, k, ln.s, ln.ag, u[0] < body.epsilon, dirty);
        if (dirty) ln.ag_dirty = true;
      } else if (agent == CAV_AGENT_PROXIMITY) {             // pedestrian.py:78-91
        const bool trigger = point_distance(ln.s[0], ln.s[1], ego_pre_x, ego_pre_y) < body.threshold;
        a0 = R(0);
        bool dirty = false;
        a1 = choose_crossing_action(sc, k, ln.s, ln.ag, trigger, dirty);
        if (dirty) ln.ag_dirty = true;
      }
      // action_space.contains (environment.py:120): the on-device agents' actions are valid by construction
      if (pelican ? !(a0 == R(0) || a0 == R(1) || a0 == R(2) || a0 == R(3))
                  : !(a0 >= k.amin && a0 <= k.amax && a1 >= k.smin && a1 <= k.smax)) bits |= TB_INVALID;
      // ---- body.step, then the half extents of the body's AABB
      ln.held[0] = a0; ln.held[1] = a1;
      if (pelican) {   // PelicanCrossing.step (bodies.py:450-461)
        if (a0 == R(1)) ln.s[0] = R(0);
        else if (a0 == R(2)) ln.s[0] = R(1);
        else if (a0 == R(3)) ln.s[0] = R(2);
      } else {
        if (body_step(k, ln.s, a0, a1, dt, ln.c, ln.sn, snapped)) ln.cs_dirty = true;
        box_extents(ln.c, ln.sn, k.hl, k.hw, ex, ey);
      }
      xch[TX_X][b][lane] = ln.s[0]; xch[TX_Y][b][lane] = ln.s[1]; xch[TX_EX][b][lane] = ex; xch[TX_EY][b][lane] = ey;
      xch[TX_C][b][lane] = ln.c; xch[TX_S][b][lane] = ln.sn;
      if (b == 0) { xch[TX_V][0][lane] = ln.s[2]; xch[TX_STEER][0][lane] = snapped; }
    }
    __syncthreads();   // ---------------- every body of the env has moved

    bool tangent_finish = false, win_ego = false;
    if (stepping) {
      const R x0 = xch[TX_X][0][lane], ex0 = xch[TX_EX][0][lane];
      {   // finish line (environment.py:148-154): every warp decides it alike
        const R margin = (x0 - ex0) - W;
        if (CAV_UNLIKELY(margin > -tau)) {
          if (margin < tau) tangent_finish = true;
          if (margin > R(0)) win_ego = true;
        }
      }
      if (!win_ego) {   // (the later stages are not reached on the finishing step; everything below is gated again when folding)
        if (sc.collisions == CAV_COLLISIONS_ALL && dynamic) {
          bool hit = false, tg = false;
#pragma unroll
          for (int r = 1; r <= M / 2; ++r) {
            // rounds 1 .. meet every unordered pair at distance r once; for an even body count round M / 2 would meet twice
            if (2 * r == M && b >= M / 2) continue;
            const int p = b + r >= M ? b + r - M : b + r;
            if (!(dynamic_bodies >> p & 1u)) continue;
            const R qx = xch[TX_X][p][lane], qy = xch[TX_Y][p][lane], qex = xch[TX_EX][p][lane], qey = xch[TX_EY][p][lane];
            const bool apart = rabs(ln.s[0] - qx) - (ex + qex) > tau || rabs(ln.s[1] - qy) - (ey + qey) > tau;
            if (!apart) {
              const DevType<R>& kp = sc.bodies[p].k;
              const Box<R> me = {ln.s[0], ln.s[1], ln.c, ln.sn, k.hl, k.hw};
              const Box<R> other = {qx, qy, xch[TX_C][p][lane], xch[TX_S][p][lane], kp.hl, kp.hw};
              // environment.py:156-177 walks i < j: lower body index first (the margin is not symmetric in the last bits)
              hit |= margin_hit(b < p ? box_margin(me, other) : box_margin(other, me), tau, tg);
            }
          }
#pragma unroll 1
          for (int s = 0; s < sc.n_statics; ++s) {   // static collidables (traffic lights, obstacle: environment.py:94-101)
            const Aabb<R> sbb = sc.static_bb[s];
            const R px = ln.s[0], py = ln.s[1];
            const bool apart = (px - ex) - sbb.x1 > tau || sbb.x0 - (px + ex) > tau || (py - ey) - sbb.y1 > tau || sbb.y0 - (py + ey) > tau;
            if (!apart) {
              if (sc.static_rect[s]) hit |= margin_hit(box_margin(Box<R>{px, py, ln.c, ln.sn, k.hl, k.hw}, sc.static_box[s]), tau, tg);
              else hit |= geo_hit(sat_pose_quad(Pose<R>{px, py, ln.s[3], ln.c, ln.sn, k.length, k.width}, &sc.quads[CAV_MAX_ROADS + s], tau), tg);
            }
          }
          if (hit) bits |= TB_HIT_ALL;
          if (tg) bits |= TB_TAN_ALL;
        }
        if (sc.offroad && b == 0) {   // the ego against every road (environment.py:179-181)
          bool on_road = false, tg = false;
#pragma unroll 1
          for (int r = 0; r < sc.n_roads; ++r) {
            const Aabb<R> rd = sc.road_bb[r];
            const R px = ln.s[0], py = ln.s[1];
            const bool apart = (px - ex) - rd.x1 > tau || rd.x0 - (px + ex) > tau || (py - ey) - rd.y1 > tau || rd.y0 - (py + ey) > tau;
            if (!apart) {
              if (sc.road_rect[r]) on_road |= margin_hit(box_margin(Box<R>{px, py, ln.c, ln.sn, k.hl, k.hw}, sc.road_box[r]), tau, tg);
              else on_road |= geo_hit(sat_pose_quad(Pose<R>{px, py, ln.s[3], ln.c, ln.sn, k.length, k.width}, &sc.quads[r], tau), tg);
            }
          }
          if (!on_road) bits |= TB_OFF;
          if (tg) bits |= TB_TAN_OFF;
        }
        if ((sc.collisions == CAV_COLLISIONS_EGO || sc.zones) && pedestrian && b > 0 && !pelican) {   // environment.py:183-206
          EgoFrame<R> f;
          const R v0 = xch[TX_V][0][lane];
          f.x = x0; f.y = xch[TX_Y][0][lane]; f.c = xch[TX_C][0][lane]; f.s = xch[TX_S][0][lane]; f.hl = k0.hl; f.hw = k0.hw;
          f.bd = (v0 * v0) * k0.inv_2brake;           // bodies.py:123
          f.td = f.bd + v0 * R(0.675);                // bodies.py:124-125
          f.have = !(f.td == R(0)) && (xch[TX_STEER][0][lane] == R(0));
          const EgoMargins<R> m = ego_margins(f, ln.s[0], ln.s[1], ln.c, ln.sn, k.hl, k.hw, tau);
          if (!m.all_clear) {
            bool tg = false;
            if (sc.collisions == CAV_COLLISIONS_EGO) {
              bool h = margin_hit(m.ego, tau, tg);
              if (!h && f.have) h = margin_hit(m.braking, tau, tg);
              if (h) bits |= TB_ZONE_HIT;
            }
            if (sc.zones && f.have) {
              if (rabs(m.reaction) < tau) bits |= TB_NEAR_REACTION;
              if (!(m.reaction > R(0))) bits |= TB_ZONE_WIN;
            }
            if (tg) bits |= TB_TAN_ZONE;
          }
        }
      }
      // ---- liveness (environment.py:139-146; a rollout has no consumer for the rewards themselves)
      if (b > 0) {
        R p = R(0);
        bool near = false;
        if (pelican) {
          p = body.static_share;
        } else {
#pragma unroll 1
          for (int r = 0; r < sc.n_roads; ++r) {
            const R q = team_road_share(sc, k, ln.s[0], ln.s[1], ln.s[3], ln.c, ln.sn, r, ex, ey, tau, near);
            if (r == 0 || q > p) p = q;
          }
        }
        if (p > R(0.5)) { ln.live += 1; ln.live_dirty = true; }
        if (near) bits |= TB_TAN_SHARE;
      }
      // ---- crossing agents' process_feedback on the new state (simulation.py:86-87: a function of the body's own state)
      if (crossing_agent) {
        bool dirty = false;
        crossing_feedback(sc, ln.s, ln.ag, dirty);
        if (dirty) ln.ag_dirty = true;
      }
      flg[b][lane] = bits;
      liv[b][lane] = ln.live;
    }
    __syncthreads();   // ---------------- every body's decision bits are posted

    if (stepping) {
      // ---- fold the env's M words in the reference's cascade order (environment.py:148-206); every warp alike
      uint32_t any = 0u;
      uint32_t w[M];
#pragma unroll
      for (int p = 0; p < M; ++p) { w[p] = flg[p][lane]; any |= w[p]; }
      bool tangent = tangent_finish || (any & TB_TAN_SHARE) != 0;
      bool terminate = win_ego;
      int win_tester = -1;
      if (!terminate && sc.collisions == CAV_COLLISIONS_ALL) {
        tangent |= (any & TB_TAN_ALL) != 0;
        terminate = (any & TB_HIT_ALL) != 0;
      }
      if (!terminate && sc.offroad) {
        tangent |= (w[0] & TB_TAN_OFF) != 0;
        terminate = (w[0] & TB_OFF) != 0;
      }
      if (!terminate && (sc.collisions == CAV_COLLISIONS_EGO || sc.zones)) {
        // The reference walks the pedestrians in index order: ego box / braking zone for each, the reaction zone only while
        // nothing has hit and nobody has won yet; a reaction-zone decision counts towards the near-tangent flag only if the
        // sequential walk would have made it (as in kernels_dense.cuh).
        tangent |= (any & TB_TAN_ZONE) != 0;
        int first_hit = M, first_win = M;
        if (CAV_UNLIKELY((any & (TB_ZONE_HIT | TB_ZONE_WIN | TB_NEAR_REACTION)) != 0)) {   // (almost every step: nobody near a zone)
#pragma unroll
        for (int p = M - 1; p >= 1; --p) if (w[p] & TB_ZONE_HIT) first_hit = p;
#pragma unroll
        for (int p = M - 1; p >= 1; --p) if ((w[p] & TB_ZONE_WIN) && p < first_hit) first_win = p;
#pragma unroll
        for (int p = 1; p < M; ++p) if ((w[p] & TB_NEAR_REACTION) && p < first_hit && p <= first_win) tangent = true;
        }
        if (first_hit < M) { terminate = true; win_tester = -1; }
        else { win_tester = first_win < M ? first_win : -1; terminate = win_tester >= 0; }
      }
      env.t_ep += 1;
      env.winner = terminate ? (win_ego ? 0 : (win_tester >= 0 ? win_tester : -1)) : -1;
      if (terminate) env.done = 1;
      else if (env.t_ep >= sc.max_timesteps) env.done = 2;   // cut off by Simulation.run (simulation.py:69-70)
      if (b == 0) {
        if (CAV_UNLIKELY(any & TB_INVALID)) buf.err[e] = 1;
        if (CAV_UNLIKELY(tangent)) count_tangent(buf.stats);
        if (CAV_UNLIKELY(env.done != 0)) {   // reporting.analyse_episode (reporting.py:227-243)
          long long sum = 0;
#pragma unroll
          for (int p = 1; p < M; ++p) sum += liv[p][lane];
          const EpisodeLog log{buf.ep_log, buf.ep_log_count, buf.ep_log_capacity, buf.shard + e, env.episode};
          score_episode<R, 1>(buf.stats, env.t_ep, env.winner, sum, log);
        }
      }
    }
  }
  if (!in_range) return;

  // ---- store (store_env of kernels_small.cuh, this warp's body)
  if (auto_reset && env.done) {
    env.episode += 1;
    team_reset_lane(sc, buf, body, b, e, (uint32_t)env.episode, ln);
    env.t_ep = 0; env.done = 0; env.winner = -1;
    was_reset = true;
  }
  const bool all = was_reset;
#pragma unroll
  for (int w = 0; w < 4; ++w) buf.state[((int64_t)b * 4 + w) * n + e] = ln.s[w];
  if (all || buf.log_actions || agent == CAV_AGENT_RANDOM) {
    buf.action[((int64_t)b * 2 + 0) * n + e] = ln.held[0];
    buf.action[((int64_t)b * 2 + 1) * n + e] = ln.held[1];
  }
  if (all || ln.live_dirty) buf.liveness[(int64_t)b * n + e] = ln.live;
  if (all || ln.cs_dirty) {
    buf.cs[((int64_t)b * 2 + 0) * n + e] = ln.c;
    buf.cs[((int64_t)b * 2 + 1) * n + e] = ln.sn;
  }
  if (all || (crossing_agent && ln.ag_dirty)) {
#pragma unroll
    for (int w = 0; w < CAV_AGENT_WORDS; ++w) buf.agent[((int64_t)b * CAV_AGENT_WORDS + w) * n + e] = ln.ag[w];
  }
  if (b == 0) {
    buf.t_ep[e] = env.t_ep;
    if (all || env.done) { buf.done[e] = env.done; buf.winner[e] = env.winner; }
    if (all) buf.episode[e] = env.episode;
  }
}

// ---------------------------------------------------------------- host side
template <typename R>
struct TeamLaunchers {
  // false = the scenario is not one this kernel takes (the thread-per-env kernel runs)
  bool (*rollout)(const DevScenario<R>&, const EnvBuffers<R>&, int64_t t_global, int n_steps, int auto_reset, cudaStream_t);
};

template <typename R, int M>
void launch_team_typed(const DevScenario<R>& sc, const EnvBuffers<R>& buf, int64_t t_global, int n_steps, int auto_reset, cudaStream_t stream) {
  const int64_t envs = buf.hi - buf.lo;
  team_rollout_kernel<R, M><<<(unsigned)((envs + 31) / 32), 32 * M, 0, stream>>>(sc, buf, t_global, n_steps, auto_reset);
}

template <typename R>
bool launch_team_rollout(const DevScenario<R>& sc, const EnvBuffers<R>& buf, int64_t t_global, int n_steps, int auto_reset,
                         cudaStream_t stream) {
  if (sc.homogeneous || sc.has_election) return false;
  switch (sc.n_bodies) {
    case 3: launch_team_typed<R, 3>(sc, buf, t_global, n_steps, auto_reset, stream); return true;
    case 4: launch_team_typed<R, 4>(sc, buf, t_global, n_steps, auto_reset, stream); return true;
    case 5: launch_team_typed<R, 5>(sc, buf, t_global, n_steps, auto_reset, stream); return true;
    case 6: launch_team_typed<R, 6>(sc, buf, t_global, n_steps, auto_reset, stream); return true;
    case 7: launch_team_typed<R, 7>(sc, buf, t_global, n_steps, auto_reset, stream); return true;
    case 8: launch_team_typed<R, 8>(sc, buf, t_global, n_steps, auto_reset, stream); return true;
    default: return false;
  }
}

template <typename R> const TeamLaunchers<R>* team_launchers();   // defined in team.cu

}  // namespace cav
