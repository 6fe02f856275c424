// dev_types.cuh — device-side tables and per-environment register state.
//
// The scenario (include/cavgym.h CavScenario) is converted once on the host into
// DevScenario<R> in the engine's arithmetic type R and passed to every kernel BY VALUE
// as a __grid_constant__ parameter: all threads read it through the constant bank
// (uniform loads), no global traffic.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/cavgym.h"

namespace cav {

template <typename R>
struct Quad {
  R x[4], y[4];  // rear_left, front_left, front_right, rear_right — clockwise (y up)
};

#define CAV_HEADING_CACHE 4

template <typename R>
struct DevType {
  R length, width, wheelbase, vmin, vmax, amin, amax, smin, smax;
  R inv_2brake;  // 1 / (2 * -min_throttle): braking distance = v*v * inv_2brake (bodies.py:123)
  R kk_smin, kk_smax;  // wheelbase / tan(steering limit), host libm: full-lock turns skip tan and one division
};

template <typename R>
struct Aabb {
  R x0, x1, y0, y1;
};

template <typename R>
struct DevBody {
  int32_t kind, type_id, flags, agent, spawn_id, pad;
  double epsilon;  // compared against a 53-bit draw in double in both modes
  R threshold;
  R init[4];       // init_state (bodies.py:27); PelicanCrossing: light state in [0]
  DevType<R> k;    // this body's DynamicBodyConstants row, resolved on the host
  Quad<R> sbox;    // PelicanCrossing static box
  Aabb<R> sbox_bb;
};

template <typename R>
struct DevSpawn {
  int32_t n_boxes, n_orient;
  Quad<R> boxes[CAV_MAX_SPAWN_BOXES];
  R orient[CAV_MAX_SPAWN_ORIENT];
  R velocity;
  R init[4];  // unused padding for alignment / future use
};

template <typename R>
struct DevScenario {
  int32_t n_bodies, n_types, n_roads, n_statics, n_spawns;
  int32_t collisions, zones, offroad;
  int64_t max_timesteps;
  R reward_win, reward_draw, cost_step, W, dt, v_maint, v_off;
  R inv_W, inv_v_off;  // reciprocals of the two constant divisors on the always-executed path
  R tau;         // near-tangent tolerance (px)
  R target_err;  // TARGET_ERROR (dynamic_body.py:8); widened for float
  R cl[4];       // centre line start x,y end x,y
  // Headings bodies are created with (spawn orientations, init orientations): cos/sin evaluated ON THE HOST
  // with the C library the reference's math.cos/math.sin use, so the common straight-walking case needs no
  // device sincos and is bit-equal to the reference.
  int32_t hc_n, hc_pad;
  R hc_theta[CAV_HEADING_CACHE], hc_cos[CAV_HEADING_CACHE], hc_sin[CAV_HEADING_CACHE];
  Quad<R> roads[CAV_MAX_ROADS];
  Aabb<R> road_bb[CAV_MAX_ROADS];
  int32_t road_axis[CAV_MAX_ROADS];  // 1 if the road rectangle is exactly axis-aligned (road == its AABB)
  Quad<R> statics[CAV_MAX_STATICS];
  Aabb<R> static_bb[CAV_MAX_STATICS];
  DevBody<R> bodies[CAV_SMALL_M];
};

// Per-env buffers owned by the engine (SoA, env fastest) + optional per-call I/O.
template <typename R>
struct EnvBuffers {
  int64_t n;            // envs in this engine (= stride of every per-env array)
  int64_t lo, hi;       // env range this launch covers (chunked host pipeline); normally [0, n)
  int64_t shard;        // global id of env 0
  uint64_t seed;
  R* state;             // [M][4][N]
  R* action;            // [M][2][N] last joint action (RandomAgent's held action)
  R* agent;             // [M][5][N] crossing-agent state, NaN = None
  R* cs;                // [M][2][N] cos/sin of each body's heading; refreshed only when the heading changes
  int32_t* liveness;    // [M][N]
  int32_t* t_ep;        // [N]
  int32_t* episode;     // [N]
  int32_t* winner;      // [N]
  uint8_t* done;        // [N] 0 live, 1 done, 2 cut off at max_timesteps
  uint8_t* err;         // [N]
  unsigned long long* stats;  // [CAV_N_STATS]
  const DevSpawn<R>* spawns;
  const double* uni_override;    // [M][3][N] or null
  const double* spawn_override;  // [M][5][N] or null
  int32_t log_actions;           // store every on-device agent's action in `action` (default: RandomAgent only)
};

template <typename R>
struct StepIO {
  const R* actions;  // [M][2][N] or null
  R* state_out;      // nullable; == state means in place
  R* reward_out;
  uint8_t* done_out;
  int32_t* winner_out;
  uint8_t* tangent_out;
};

}  // namespace cav
