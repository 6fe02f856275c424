// transition.cuh — one joint transition of ONE environment, entirely in registers.
//
// Restates, in the reference's order (library/environment.py:119-223):
//   joint action        agents' choose_action on the pre-step state        simulation.py:71
//   validation          action_space.contains                              environment.py:120
//   body.step           kinematic bicycle / traffic-light state            environment.py:122-123
//   bounding boxes      info()['body_polygons']                            environment.py:107,125-127
//   rewards, liveness                                                      environment.py:131-146
//   termination cascade finish > all-collisions > off-road > ego-collision > reaction zone   :148-206
//   terminal rewards, winner                                               environment.py:208-220
//   process_feedback    crossing agents, post-step state                   simulation.py:86-87
// Used by the step, replay and rollout kernels (kernels_small.cuh), so the three share one
// definition of a transition.  Thread-per-environment: every array below is indexed by
// compile-time constants after unrolling and lives in registers.
//
// Boxes never become corner lists here: each body is (x, y, cos, sin, half length, half width) and every test is
// written on that form (geometry.cuh).  What stays out of line: general quads, road corners, libm.
#pragma once
#include "agents.cuh"

namespace cav {

template <typename R, int M>
struct EnvRegs {
  R s[M][4];         // x, y, v, theta (PelicanCrossing: light state in [0])
  R held[M][2];      // last action of on-device agents (RandomAgent holds it)
  R ag[M][CAV_AGENT_WORDS];
  R cs[M][2];        // cos, sin of the heading (EnvBuffers::cs)
  int32_t live[M];   // episode_liveness (environment.py:144-146); [0] unused
  int32_t t_ep, episode, winner;
  uint32_t live_dirty;  // bit b: live[b] changed since it was loaded
  uint32_t cs_dirty;  // bit b: cs[b] changed since it was loaded
  uint32_t ag_dirty;  // bit b: ag[b] changed since it was loaded
  uint8_t done;
};

// Episode statistics (reporting.py:227-269) are updated with atomics at the moment an episode ends or a near-tangent
// step is flagged — about one env-step in a thousand — so no per-thread accumulators stay live through the step.
__device__ __forceinline__ void count_tangent(unsigned long long* stats) { atomicAdd(&stats[CAV_STAT_TANGENT], 1ull); }

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename R, int M>
__device__ __forceinline__ bool uses_agent_state(const DevScenario<R>& sc, int b) {
  return sc.bodies[b].agent == CAV_AGENT_RANDOM_CONSTRAINED || sc.bodies[b].agent == CAV_AGENT_PROXIMITY;
}

// Body b as a box / as a pose; b must be a compile-time constant at the call site (unrolled loops).
template <typename R, int M>
__device__ __forceinline__ Box<R> body_box(const DevScenario<R>& sc, const EnvRegs<R, M>& env, const int b) {
  return {env.s[b][0], env.s[b][1], env.cs[b][0], env.cs[b][1], sc.bodies[b].k.hl, sc.bodies[b].k.hw};
}
template <typename R, int M>
__device__ __forceinline__ Pose<R> body_pose(const DevScenario<R>& sc, const EnvRegs<R, M>& env, const int b) {
  const DevBody<R>& body = sc.bodies[b];
  return {env.s[b][0], env.s[b][1], env.s[b][3], env.cs[b][0], env.cs[b][1], body.k.length, body.k.width};
}

// closed predicate from a separating-axis margin; flags near-tangent decisions
template <typename R>
__device__ __forceinline__ bool margin_hit(R margin, R tau, bool& tangent) {
  if (rabs(margin) < tau) tangent = true;
  return !(margin > R(0));
}
__device__ __forceinline__ bool geo_hit(int r, bool& tangent) {
  if (r & GEO_TANGENT) tangent = true;
  return (r & GEO_HIT) != 0;
}

// Shape.percentage_intersects(body box, road r) (environment.py:141, geometry.py:80-87).
//   AABB clear of the road            -> 0        (disjoint)
//   axis-aligned road, clearly inside -> 1        (contained)
//   axis-aligned road, one kerb       -> closed form (kerb_share)
//   anything else                     -> general predicates, out of line
// `near` is raised when the share is within tau of the liveness threshold 0.5 (environment.py:144) or a general
// predicate was near-tangent.  GENERIC = false: every road is known to be an axis-aligned rectangle.
template <typename R, int M, bool GENERIC>
__device__ __forceinline__ R road_share(const DevScenario<R>& sc, const EnvRegs<R, M>& env, const int b, int r, R ex, R ey,
                                        R tau, bool& near) {
  const Aabb<R> rd = sc.road_bb[r];
  const R px = env.s[b][0], py = env.s[b][1];
  const R m0 = (px - ex) - rd.x0, m1 = rd.x1 - (px + ex), m2 = (py - ey) - rd.y0, m3 = rd.y1 - (py + ey);
  const R mx = rmin(m0, m1), my = rmin(m2, m3);
  if (mx < -((ex + ex) + tau) || my < -((ey + ey) + tau)) return R(0);
  if (!GENERIC || sc.road_axis[r]) {
    const bool x_edge = mx < my;
    const R lo = x_edge ? mx : my, hi = x_edge ? my : mx;
    if (lo >= tau) return R(1);
    // one kerb: the smallest margin is clearly negative, the opposite edge and the other pair of edges clearly inside
    const R opposite = x_edge ? rmax(m0, m1) : rmax(m2, m3);
    if (lo <= -tau && hi >= tau && opposite >= tau) {
      const R ac = rabs(env.cs[b][0]), as = rabs(env.cs[b][1]);
      const R hl = sc.bodies[b].k.hl, hw = sc.bodies[b].k.hw;
      CAV_DBG(7);
      const R p = kerb_share(lo + (x_edge ? ex : ey), (x_edge ? ac : as) * hl, (x_edge ? as : ac) * hw);
      if (rabs(p - R(0.5)) < tau) near = true;
      return p;
    }
    // a corner of the road: one x-edge and one y-edge crossed (or touched), the opposite edges clearly inside
    if (mx < tau && my < tau && rmax(m0, m1) >= tau && rmax(m2, m3) >= tau) {
      const bool low_x = m0 < m1, low_y = m2 < m3;
      const R p = corner_share(body_pose<R, M>(sc, env, b), low_x ? R(-1) : R(1), low_x ? -rd.x0 : rd.x1, low_y ? R(-1) : R(1),
                               low_y ? -rd.y0 : rd.y1);
      if (rabs(p - R(0.5)) < tau) near = true;
      return p;
    }
  }
  const Share<R> share = road_share_general(body_pose<R, M>(sc, env, b), &sc.quads[r], tau);
  near |= (share.tangent != 0) || (rabs(share.value - R(0.5)) < tau);
  return share.value;
}

// Result of one transition besides the updated EnvRegs.
template <typename R, int M>
struct StepResult {
  R reward[M];
  bool terminate, tangent, invalid;
  int32_t winner;
};

// AGENTS = false compiles the replay-only variant: every body takes its action from `ext`.
// GENERIC = false compiles the homogeneous variant the engine selects when the scenario has no PelicanCrossing body,
// every non-ego body is a Pedestrian and every road is an axis-aligned rectangle (the Pedestrians-v0 family, any
// number of pedestrians): the per-body kind / flag tests disappear at compile time.
struct NoSink {
  template <typename E> __device__ __forceinline__ void operator()(const E&) const {}
};

// `moved(env)` is called once, right after body.step: a kernel that stages state in shared memory writes the new
// state there at that point, so x, y, v, theta need not stay in registers through the geometry.
template <typename R, int M, bool AGENTS, bool GENERIC, typename Sink = NoSink>
__device__ __forceinline__ void transition(const DevScenario<R>& sc, const EnvBuffers<R>& buf, int64_t e, int64_t t_global,
                                           EnvRegs<R, M>& env, const R (&ext)[M][2], StepResult<R, M>& out,
                                           Sink moved = Sink()) {
  const R tau = sc.tau, dt = sc.dt;
  bool tangent = false;
  R act[M][2];
  auto is_pelican = [&](int b) { return GENERIC && sc.bodies[b].kind == CAV_BODY_PELICAN; };
  auto is_pedestrian = [&](int b) { return !GENERIC || (sc.bodies[b].flags & CAV_FLAG_PEDESTRIAN) != 0; };

  // ---- joint action from the pre-step state
  bool valid = true;
#pragma unroll
  for (int b = 0; b < M; ++b) {
    const DevBody<R>& body = sc.bodies[b];
    const DevType<R>& k = body.k;
    R a0 = ext[b][0], a1 = ext[b][1];
    if (AGENTS && body.agent != CAV_AGENT_EXTERNAL) {
      a0 = env.held[b][0]; a1 = env.held[b][1];
      double u[CAV_DRAWS] = {0.0, 0.0, 0.0};
      const bool draws = body.agent == CAV_AGENT_RANDOM || body.agent == CAV_AGENT_RANDOM_CONSTRAINED;
      if (draws) {
        if (buf.uni_override) {
#pragma unroll
          for (int c = 0; c < CAV_DRAWS; ++c) u[c] = buf.uni_override[((int64_t)b * CAV_DRAWS + c) * buf.n + e];
        } else {
          draw_block(buf.seed, (uint64_t)(buf.shard + e), b, KIND_AGENT0, (uint32_t)env.episode, (uint32_t)env.t_ep, u);
        }
      }
      if (body.agent == CAV_AGENT_NOOP) {
        a0 = R(0); a1 = R(0);
      } else if (body.agent == CAV_AGENT_RANDOM) {  // RandomAgent.choose_action (template.py:52-56)
        if (u[0] < body.epsilon) {
          if (is_pelican(b)) {
            a0 = R(floor(u[1] * 4.0)); if (a0 > R(3)) a0 = R(3);
            a1 = R(0);
          } else {
            if (!buf.uni_override) {
              double w[2];
              draw_block(buf.seed, (uint64_t)(buf.shard + e), b, KIND_AGENT1, (uint32_t)env.episode, (uint32_t)env.t_ep, w);
              u[2] = w[0];
            }
            a0 = R(double(k.amin) + (double(k.amax) - double(k.amin)) * u[1]);  // Box.sample = low + (high-low)*u
            a1 = R(double(k.smin) + (double(k.smax) - double(k.smin)) * u[2]);
          }
        }
      } else if (body.agent == CAV_AGENT_RANDOM_CONSTRAINED) {  // pedestrian.py:72-75
        a0 = R(0);
        bool dirty = false;
        a1 = choose_crossing_action(sc, k, env.s[b], env.ag[b], u[0] < body.epsilon, dirty);
        if (dirty) env.ag_dirty |= 1u << b;
      } else if (body.agent == CAV_AGENT_PROXIMITY) {  // pedestrian.py:78-91
        const bool trigger = point_distance(env.s[b][0], env.s[b][1], env.s[0][0], env.s[0][1]) < body.threshold;
        a0 = R(0);
        bool dirty = false;
        a1 = choose_crossing_action(sc, k, env.s[b], env.ag[b], trigger, dirty);
        if (dirty) env.ag_dirty |= 1u << b;
      }
    }
    act[b][0] = a0; act[b][1] = a1;
    if (is_pelican(b)) valid = valid && (a0 == R(0) || a0 == R(1) || a0 == R(2) || a0 == R(3));
    else valid = valid && (a0 >= k.amin && a0 <= k.amax && a1 >= k.smin && a1 <= k.smax);
  }
  out.invalid = !valid;
  out.tangent = false;
  out.terminate = false;
  out.winner = -1;
  if (!valid) {  // AssertionError before any mutation (environment.py:120)
#pragma unroll
    for (int b = 0; b < M; ++b) out.reward[b] = R(0);
    return;
  }

  // ---- body.step, then the half extents of every body's AABB (all later tests start from them)
  R ex[M], ey[M];
  R ego_steer = R(0);
#pragma unroll
  for (int b = 0; b < M; ++b) {
    const DevBody<R>& body = sc.bodies[b];
    if (AGENTS) { env.held[b][0] = act[b][0]; env.held[b][1] = act[b][1]; }
    if (is_pelican(b)) {  // PelicanCrossing.step (bodies.py:450-461)
      if (act[b][0] == R(1)) env.s[b][0] = R(0);
      else if (act[b][0] == R(2)) env.s[b][0] = R(1);
      else if (act[b][0] == R(3)) env.s[b][0] = R(2);
      ex[b] = R(0); ey[b] = R(0);
    } else {
      const DevType<R>& k = body.k;
      R snapped;
      if (body_step(k, env.s[b], act[b][0], act[b][1], dt, env.cs[b][0], env.cs[b][1], snapped)) env.cs_dirty |= 1u << b;
      if (b == 0) ego_steer = snapped;
      box_extents(env.cs[b][0], env.cs[b][1], k.hl, k.hw, ex[b], ey[b]);
    }
  }
  moved(env);

  // ---- termination cascade (evaluated before the rewards: neither depends on the other, and the rare general
  //      road-share call below then happens with almost nothing live)
  const R W = sc.W;
  bool terminate = false, win_ego = false;
  int win_tester = -1;
  {
    const R margin = (env.s[0][0] - ex[0]) - W;  // all four ego corners x > viewer_width  <=>  min corner x > W
    if (margin > -tau) {                         // only the last steps of an episode get here
      CAV_DBG(9);
      if (margin < tau) tangent = true;
      if (margin > R(0)) { terminate = true; win_ego = true; }
    }
  }
  if (!terminate && sc.collisions == CAV_COLLISIONS_ALL) {
    bool hit = false;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      if (is_pelican(i)) continue;
#pragma unroll
      for (int j = i + 1; j < M; ++j) {
        if (is_pelican(j)) continue;
        const bool apart = rabs(env.s[i][0] - env.s[j][0]) - (ex[i] + ex[j]) > tau ||
                           rabs(env.s[i][1] - env.s[j][1]) - (ey[i] + ey[j]) > tau;
        if (!apart) hit |= margin_hit(box_margin(body_box<R, M>(sc, env, i), body_box<R, M>(sc, env, j)), tau, tangent);
      }
#pragma unroll 1
      for (int s = 0; s < sc.n_statics; ++s) {
        const Aabb<R> sb = sc.static_bb[s];
        const R px = env.s[i][0], py = env.s[i][1];
        const bool apart = (px - ex[i]) - sb.x1 > tau || sb.x0 - (px + ex[i]) > tau || (py - ey[i]) - sb.y1 > tau ||
                           sb.y0 - (py + ey[i]) > tau;
        if (!apart) {
          if (sc.static_rect[s]) hit |= margin_hit(box_margin(body_box<R, M>(sc, env, i), sc.static_box[s]), tau, tangent);
          else hit |= geo_hit(sat_pose_quad(body_pose<R, M>(sc, env, i), &sc.quads[CAV_MAX_ROADS + s], tau), tangent);
        }
      }
    }
    terminate = hit;
  }
  if (!terminate && sc.offroad) {
    bool on_road = false;
#pragma unroll 1
    for (int r = 0; r < sc.n_roads; ++r) {
      const Aabb<R> rd = sc.road_bb[r];
      const R px = env.s[0][0], py = env.s[0][1];
      const bool apart = (px - ex[0]) - rd.x1 > tau || rd.x0 - (px + ex[0]) > tau || (py - ey[0]) - rd.y1 > tau ||
                         rd.y0 - (py + ey[0]) > tau;
      if (!apart) {
        if (!GENERIC || sc.road_rect[r]) on_road |= margin_hit(box_margin(body_box<R, M>(sc, env, 0), sc.road_box[r]), tau, tangent);
        else on_road |= geo_hit(sat_pose_quad(body_pose<R, M>(sc, env, 0), &sc.quads[r], tau), tangent);
      }
    }
    terminate = !on_road;
  }
  if (!terminate && (sc.collisions == CAV_COLLISIONS_EGO || sc.zones)) {
    const DevType<R>& k0 = sc.bodies[0].k;
    EgoFrame<R> f;
    f.x = env.s[0][0]; f.y = env.s[0][1]; f.c = env.cs[0][0]; f.s = env.cs[0][1]; f.hl = k0.hl; f.hw = k0.hw;
    const R v0 = env.s[0][2];
    f.bd = (v0 * v0) * k0.inv_2brake;           // bodies.py:123
    f.td = f.bd + v0 * R(0.675);                // bodies.py:124-125 (REACTION_TIME)
    f.have = !(f.td == R(0)) && (ego_steer == R(0));
    const bool ego_mode = sc.collisions == CAV_COLLISIONS_EGO;
    bool hit = false;
#pragma unroll
    for (int b = 1; b < M; ++b) {
      if (!is_pedestrian(b)) continue;   // environment.py:192,199
      // Common case, decided with one predicate chain: the pedestrian is clear (by tau or more) of the strip swept by
      // the ego box and both zones.  Only otherwise are the closed predicates and near-tangent flags evaluated.
      const EgoMargins<R> m = ego_margins(f, env.s[b][0], env.s[b][1], env.cs[b][0], env.cs[b][1], sc.bodies[b].k.hl,
                                          sc.bodies[b].k.hw, tau);
      if (m.all_clear) continue;
      CAV_DBG(8);
      if (ego_mode) {   // environment.py:183-193: ego box, then the braking zone
        bool h = margin_hit(m.ego, tau, tangent);
        if (!h && f.have) h = margin_hit(m.braking, tau, tangent);
        hit |= h;
      }
      // environment.py:195-206: first pedestrian in the reaction zone wins (only consulted if nothing terminated above)
      if (sc.zones && f.have && win_tester < 0 && !hit) {
        if (margin_hit(m.reaction, tau, tangent)) win_tester = b;
      }
    }
    if (hit) { terminate = true; win_tester = -1; }
    else terminate = win_tester >= 0;
  }

  // ---- rewards and liveness
  const R c = sc.cost_step;
  // Both divisors are scenario constants: multiply by the host-computed reciprocal (<= 1 ulp from the division).
  const R ego_rel = rmax(R(0), rmin(R(1), (W - env.s[0][0]) * sc.inv_W));
  const R voff = rabs(env.s[0][2] - sc.v_maint) * sc.inv_v_off;
  R r0 = R(0);
  r0 -= voff * c;
  r0 += (R(1) - ego_rel) * c;
  out.reward[0] = r0;
#pragma unroll
  for (int b = 1; b < M; ++b) {
    R p = R(0);
    if (is_pelican(b)) {
      p = sc.bodies[b].static_share;  // static box vs static roads: a constant of the scenario
    } else {
#pragma unroll 1
      for (int r = 0; r < sc.n_roads; ++r) {  // max over roads (environment.py:141)
        const R q = road_share<R, M, GENERIC>(sc, env, b, r, ex[b], ey[b], tau, tangent);
        if (r == 0 || q > p) p = q;
      }
    }
    R rb = R(0);
    rb -= p * c;
    rb += ego_rel * c;
    out.reward[b] = rb;
    if (p > R(0.5)) { env.live[b] += 1; env.live_dirty |= 1u << b; }
  }

  // ---- terminal rewards and winner
  if (terminate || t_global == sc.max_timesteps - 1) {
    out.reward[0] += win_ego ? sc.reward_win : (win_tester >= 0 ? -sc.reward_win : sc.reward_draw);
#pragma unroll
    for (int b = 1; b < M; ++b)
      out.reward[b] += win_ego ? -sc.reward_win : (win_tester < 0 ? sc.reward_draw : (win_tester == b ? sc.reward_win : sc.reward_draw));
    if (win_ego) out.winner = 0;
    else if (win_tester >= 0) out.winner = win_tester;
  }

  // ---- crossing agents' process_feedback on the new state
  if (AGENTS) {
#pragma unroll
    for (int b = 0; b < M; ++b) {
      if (uses_agent_state<R, M>(sc, b)) {
        bool dirty = false;
        crossing_feedback(sc, env.s[b], env.ag[b], dirty);
        if (dirty) env.ag_dirty |= 1u << b;
      }
    }
  }

  env.t_ep += 1;
  env.winner = out.winner;
  if (terminate) env.done = 1;
  else if (env.t_ep >= sc.max_timesteps) env.done = 2;  // cut off by Simulation.run (simulation.py:69-70), not `done`
  out.terminate = terminate;
  out.tangent = tangent;
}

// reporting.analyse_episode (reporting.py:227-243) for an env whose episode just ended.
template <typename R, int M>
__device__ __noinline__ void score_episode(unsigned long long* stats, int32_t t_ep, int32_t winner, long long liveness_sum) {
  const unsigned long long t = (unsigned long long)t_ep;
  atomicAdd(&stats[CAV_STAT_EPISODES], 1ull);
  atomicAdd(&stats[CAV_STAT_SUM_T], t);
  atomicAdd(&stats[CAV_STAT_SUM_T2], t * t);
  if (winner > 0) {
    const long long score = -liveness_sum;
    atomicAdd(&stats[CAV_STAT_INTERESTING], 1ull);
    atomicAdd(&stats[CAV_STAT_SUM_SCORE], (unsigned long long)score);  // two's complement: wraps to the signed sum
    atomicAdd(&stats[CAV_STAT_SUM_SCORE2], (unsigned long long)(score * score));
  }
}

template <typename R, int M>
__device__ __forceinline__ void score_episode(const EnvBuffers<R>& buf, const EnvRegs<R, M>& env) {
  long long sum = 0;
#pragma unroll
  for (int b = 1; b < M; ++b) sum += env.live[b];
  score_episode<R, M>(buf.stats, env.t_ep, env.winner, sum);
}

// CAVEnv.reset for one env (environment.py:225-229): bodies back to init_state (SpawnPedestrians re-drawn),
// liveness zeroed, agents reset, on-device noop action.  `init` (nullable) replays given initial states.
template <typename R, int M>
__device__ __forceinline__ void reset_env(const DevScenario<R>& sc, const EnvBuffers<R>& buf, const R* init, int64_t e,
                                          EnvRegs<R, M>& env) {
  env.episode += 1;
#pragma unroll
  for (int b = 0; b < M; ++b) {
    const DevBody<R>& body = sc.bodies[b];
    R st[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) st[c] = body.init[c];
    if (init) {
#pragma unroll
      for (int c = 0; c < 4; ++c) st[c] = init[((int64_t)b * 4 + c) * buf.n + e];
    } else if ((body.flags & CAV_FLAG_SPAWN) && body.spawn_id >= 0) {
      double u[5];
      if (buf.spawn_override) {
#pragma unroll
        for (int c = 0; c < 5; ++c) u[c] = buf.spawn_override[((int64_t)b * 5 + c) * buf.n + e];
      } else {
        double w[2];
        const uint64_t g = (uint64_t)(buf.shard + e);
        draw_block(buf.seed, g, b, KIND_SPAWN0, (uint32_t)env.episode, 0u, w); u[0] = w[0]; u[1] = w[1];
        draw_block(buf.seed, g, b, KIND_SPAWN1, (uint32_t)env.episode, 0u, w); u[2] = w[0]; u[3] = w[1];
        draw_block(buf.seed, g, b, KIND_SPAWN2, (uint32_t)env.episode, 0u, w); u[4] = w[0];
      }
      spawn_body(buf.spawns[body.spawn_id], u, st);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) env.s[b][c] = st[c];
    env.held[b][0] = R(0); env.held[b][1] = R(0);
    env.cs[b][0] = R(1); env.cs[b][1] = R(0);
    if (body.kind == CAV_BODY_DYNAMIC) heading_cs(sc, st[3], env.cs[b][0], env.cs[b][1]);
#pragma unroll
    for (int w = 0; w < CAV_AGENT_WORDS; ++w) env.ag[b][w] = nan_<R>();
    env.live[b] = 0;
  }
  env.t_ep = 0;
  env.done = 0;
  env.winner = -1;
  env.ag_dirty = 0xFFFFFFFFu;  // everything must be written back
  env.live_dirty = 0xFFFFFFFFu;
  env.cs_dirty = 0xFFFFFFFFu;
}

}  // namespace cav
