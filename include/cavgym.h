/* cavgym.h — C-ABI of libcavgym_sm100.so, the B200 batched stepping engine for CAV-Gym.
 *
 * The reference (TSL-UOB/CAV-Gym) has no FFI: its boundary is the Python protocol
 *   CAVEnv(bodies, constants, env_config, np_random)        library/environment.py:59
 *   CAVEnv.reset() -> joint_obs                              library/environment.py:225-229
 *   CAVEnv.step(joint_action) -> (obs, reward, done, info)   library/environment.py:119-223
 * Every entry point below names the reference interface it replaces.  Plain
 * pointers and sizes only; no torch types.  Device buffers are raw device
 * pointers (torch.Tensor.data_ptr()); the caller owns every I/O buffer, the
 * engine owns its tables, body state, agent state, RNG counters and liveness.
 *
 * Layout (all per-env arrays are SoA with the environment index fastest):
 *   state    [M][4][N]  x, y, velocity, orientation          (bodies.py:48-60)
 *   actions  [M][2][N]  throttle, steering_angle             (bodies.py:104-109)
 *                       PelicanCrossing body: TrafficLightAction value in [.][0]
 *   reward   [M][N]     joint reward                          (environment.py:131-146,208-213)
 *   done     [N] u8, winner [N] i32 (-1 = no 'winner' key)    (environment.py:148-220)
 *   liveness [M][N] i32 episode_liveness                      (environment.py:144-146)
 * real = double (dtype 0) or float (dtype 1).
 *
 * Errors: every call returns 0 or a negative CAV_E* code and never throws;
 * cavgym_last_error() gives the message of the calling thread's last failure.
 * One engine per device; an engine is not thread-safe.  No call synchronises the
 * host with the device except cavgym_stats, cavgym_error_count and the *_host calls.
 */
#ifndef CAVGYM_H
#define CAVGYM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define CAV_MAX_ROADS 4         /* road_map.roads: major + minor roads (assets.py:115-120) */
#define CAV_MAX_STATICS 8       /* traffic lights + obstacle (environment.py:94-101) */
#define CAV_MAX_SPAWN_BOXES 2   /* pedestrians.py:45-68 */
#define CAV_MAX_SPAWN_ORIENT 4
#define CAV_MAX_TYPES 8
#define CAV_SMALL_M 8           /* thread-per-env fused kernels up to this many bodies */
#define CAV_MAX_BODIES 512      /* warp-per-env dense kernels (bodies staged in shared memory) up to this many bodies */
#define CAV_AGENT_WORDS 5       /* CrossingAgent state (pedestrian.py:14-31); NaN = None */
#define CAV_DRAWS 3             /* [0,1) draws an agent may consume per step */

enum { CAV_F64 = 0, CAV_F32 = 1 };

enum {                          /* negative return codes */
  CAV_OK = 0,
  CAV_EINVAL = -22,             /* bad argument / unsupported scenario */
  CAV_ENOMEM = -12,
  CAV_ECUDA = -5,               /* a CUDA runtime call failed */
  CAV_ESTATE = -1               /* call not valid in the engine's current state */
};

enum { CAV_BODY_DYNAMIC = 0,    /* DynamicBody and subclasses (bodies.py:78-346) */
       CAV_BODY_PELICAN = 1 };  /* PelicanCrossing (bodies.py:409-461) */

enum { CAV_FLAG_PEDESTRIAN = 1, /* isinstance(body, Pedestrian) (environment.py:192,199) */
       CAV_FLAG_SPAWN = 2 };    /* SpawnPedestrian: re-drawn on reset (bodies.py:290-312) */

enum { CAV_AGENT_EXTERNAL = 0,            /* action comes from the `actions` buffer */
       CAV_AGENT_NOOP = 1,                /* NoopAgent            (template.py:24-37) */
       CAV_AGENT_RANDOM = 2,              /* RandomAgent          (template.py:40-62) */
       CAV_AGENT_RANDOM_CONSTRAINED = 3,  /* RandomConstrainedAgent (pedestrian.py:72-75) */
       CAV_AGENT_PROXIMITY = 4,           /* ProximityAgent       (pedestrian.py:78-91) */
       CAV_AGENT_ELECTION = 5 };          /* ElectionAgent arbitrated by Election (pedestrian.py:94-116, election.py:4-57);
                                             scenarios of at most CAV_SMALL_M bodies */

enum { CAV_COLLISIONS_NONE = 0, CAV_COLLISIONS_EGO = 1, CAV_COLLISIONS_ALL = 2 }; /* config.py:237-240 */

/* Convex quadrilateral, corners in the reference's order rear_left, front_left,
 * front_right, rear_right (geometry.py:96-107). */
typedef struct CavQuad { double x[4]; double y[4]; } CavQuad;

/* DynamicBodyConstants (bodies.py:63-75); track is cosmetic and omitted. */
typedef struct CavBodyType {
  double length, width, wheelbase;
  double min_velocity, max_velocity;
  double min_throttle, max_throttle;
  double min_steering_angle, max_steering_angle;
} CavBodyType;

/* SpawnPedestrianState (bodies.py:283-287). */
typedef struct CavSpawn {
  int32_t n_boxes, n_orientations;
  CavQuad boxes[CAV_MAX_SPAWN_BOXES];
  double orientations[CAV_MAX_SPAWN_ORIENT];
  double velocity;
} CavSpawn;

typedef struct CavBody {
  int32_t kind;              /* CAV_BODY_* */
  int32_t type_id;           /* row of CavScenario.types (dynamic bodies) */
  int32_t flags;             /* CAV_FLAG_* */
  int32_t spawn_id;          /* row of CavScenario.spawns, or -1 */
  int32_t agent;             /* CAV_AGENT_* */
  int32_t reserved;
  double agent_epsilon;      /* RandomAgent / RandomConstrainedAgent epsilon */
  double agent_threshold;    /* ProximityAgent distance_threshold */
  double init_state[4];      /* x, y, v, theta; PelicanCrossing: TrafficLightState in [0] */
  CavQuad static_box;        /* PelicanCrossing.static_bounding_box (bodies.py:415) */
} CavBody;

/* Everything CAVEnv.__init__ receives, flattened (environment.py:59-92). */
typedef struct CavScenario {
  int32_t n_bodies, n_types, n_roads, n_statics, n_spawns;
  int32_t terminate_collisions;      /* CAV_COLLISIONS_* (config.json "terminate_collisions") */
  int32_t terminate_ego_zones;       /* bool */
  int32_t terminate_ego_offroad;     /* bool */
  int64_t max_timesteps;
  double reward_win, reward_draw, cost_step;
  double viewer_width;               /* CAVEnvConstants.viewer_width (environment.py:47-51) */
  double time_resolution;            /* 1/frequency = 1/60 (environment.py:65-66) */
  double ego_maintenance_velocity;   /* environment.py:87 */
  double ego_max_velocity_offset;    /* environment.py:88 */
  double centre_line[4];             /* major road longitudinal_line(): start x,y, end x,y (config.py:363) */
  CavQuad roads[CAV_MAX_ROADS];      /* roads[0] = major road */
  CavQuad statics[CAV_MAX_STATICS];  /* collidable traffic lights, then the obstacle */
  CavBodyType types[CAV_MAX_TYPES];
  const CavSpawn* spawns;            /* host pointer, n_spawns rows (copied by cavgym_create) */
  const CavBody* bodies;             /* host pointer, n_bodies rows (copied by cavgym_create) */
} CavScenario;

typedef struct CavEngine CavEngine;

/* One finished episode as reporting.analyse_episode scores it (reporting.py:227-243) — the fields of an episode.log
 * row (reporting.py:157-158) that exist for an environment of a batch.  interesting <=> winner > 0, score = -liveness_sum
 * when interesting (else NaN), completed <=> timesteps == max_timesteps. */
typedef struct CavEpisodeRow {
  int64_t env;           /* global env id (shard offset included) */
  int32_t episode;       /* the env's own episode count, 1 = its first episode */
  int32_t timesteps;
  int32_t winner;        /* info['winner'] of the last step, -1 when absent */
  int32_t liveness_sum;  /* sum(env.episode_liveness[1:]) */
} CavEpisodeRow;

/* Indices of cavgym_stats' twelve int64 counters (reporting.py:227-269 definitions).  SUM_T / SUM_T2 run over every
 * finished episode; SUM_T_INTERESTING / SUM_T2_INTERESTING only over episodes with winner > 0, which is what the reference's
 * run summary reports as the timesteps of its "interesting" tests (reporting.py:246-255). */
enum { CAV_STAT_EPISODES = 0, CAV_STAT_INTERESTING = 1, CAV_STAT_SUM_T = 2, CAV_STAT_SUM_T2 = 3,
       CAV_STAT_SUM_SCORE = 4, CAV_STAT_SUM_SCORE2 = 5, CAV_STAT_ENV_STEPS = 6, CAV_STAT_BODY_STEPS = 7,
       CAV_STAT_TANGENT = 8, CAV_STAT_ERRORS = 9, CAV_STAT_SUM_T_INTERESTING = 10, CAV_STAT_SUM_T2_INTERESTING = 11,
       CAV_N_STATS = 12 };

/* ---- lifecycle ------------------------------------------------------------------- */

/* CAVEnv.__init__ x n_envs (environment.py:59-92; Config.setup config.py:272-415).
 * Copies the tables to `device`, allocates per-env storage and performs the
 * constructor-time spawn of every SpawnPedestrian (bodies.py:296).  `seed` keys
 * the Philox streams of the on-device agents and spawners. */
int cavgym_create(const CavScenario* tables, int64_t n_envs, int dtype, int device, uint64_t seed, CavEngine** out);
int cavgym_destroy(CavEngine* engine);

/* Env-sharding (experiments.py:121-122 analogue): this engine holds global env ids
 * [offset, offset + n_envs).  Philox is keyed by the GLOBAL id, so results do not
 * depend on how envs are split across GPUs.  Call before the first reset. */
int cavgym_set_shard(CavEngine* engine, int64_t global_env_offset);

/* CAVEnv.reset (environment.py:225-229; Body.reset bodies.py:32-33,111-114;
 * SpawnPedestrian.reset bodies.py:298-300; Agent.reset pedestrian.py:27-31).
 * mask  : device u8[N] or NULL (= all); envs with mask!=0 are reset.
 * init_state: device real[M][4][N] or NULL; when given it REPLACES the spawn draw
 *         (replay of reference-generated initial states). */
int cavgym_reset(CavEngine* engine, const uint8_t* mask, const void* init_state, cudaStream_t stream);

/* ---- stepping -------------------------------------------------------------------- */

/* CAVEnv.step (environment.py:119-223) over all N envs, one fused launch.
 * actions : device real[M][2][N]; may be NULL iff no body has CAV_AGENT_EXTERNAL
 *           (bodies with an on-device agent ignore their rows).
 * state_out, reward_out, done_out, winner_out, tangent_flag_out: device, each
 *           nullable.  state_out may be cavgym_state_ptr() (no extra write).
 * An env whose episode has ended stays frozen (reward 0, done 1) until it is reset.
 * Invalid actions (environment.py:120) set the env's error flag instead of aborting
 * the batch; see cavgym_error_count. */
int cavgym_step(CavEngine* engine, const void* actions, void* state_out, void* reward_out,
                uint8_t* done_out, int32_t* winner_out, uint8_t* tangent_flag_out, cudaStream_t stream);

/* Simulation.run's inner loop (simulation.py:70-97) on device: n_steps fused
 * transitions with the on-device agents; with auto_reset, an env whose episode ends
 * (done, or max_timesteps reached, simulation.py:69-70) is scored
 * (reporting.py:227-243) and reset in the same kernel. */
int cavgym_rollout(CavEngine* engine, int n_steps, int auto_reset, cudaStream_t stream);

/* Replayed joint actions, n_steps transitions in one launch:
 * actions real[T][M][2][N]; trajectory outputs (each nullable) state real[T][M][4][N],
 * reward real[T][M][N], done u8[T][N], winner i32[T][N], tangent u8[T][N].
 * Every pointer may also be page-locked, device-mapped HOST memory (cavgym_host_alloc, tensor.pin_memory()): the kernel
 * then streams actions in and trajectories out over PCIe inside the launch (BatchedCAVEnv.replay_host).
 * Calls that follow one another directly (no other call on this engine in between) on the same stream are ordered env tile
 * by env tile rather than grid by grid, so their launches overlap on the device; to everything else in the stream — and to
 * the host — they complete in order as usual.  Engine state (cavgym_state_ptr etc.) must not be written by other work of the
 * stream between two such calls without an API call on the engine in between (cavgym_reset, cavgym_step ... all qualify). */
int cavgym_replay(CavEngine* engine, int n_steps, const void* actions, void* state_traj, void* reward_traj,
                  uint8_t* done_traj, int32_t* winner_traj, uint8_t* tangent_traj, cudaStream_t stream);

/* cavgym_step with HOST buffers (same shapes); returns when the outputs are on the host.  Pinned buffers
 * (cudaHostAlloc / cudaHostRegister / torch pin_memory) are read and written by the step kernel itself, zero copy;
 * pageable buffers are copied in and out in env-chunks pipelined over internal streams. */
int cavgym_step_host(CavEngine* engine, const void* actions, void* state_out, void* reward_out,
                     uint8_t* done_out, int32_t* winner_out, uint8_t* tangent_flag_out);

/* cavgym_step_host for an fp64 engine with a float32 WIRE format: actions, state_out and reward_out are float32 host
 * arrays (same shapes); the engine widens the actions, steps in double and narrows the results, so half the bytes cross
 * PCIe.  What the caller gives up is the last 29 bits of every action and observation, nothing else: the engine's state and
 * all events stay those of the fp64 engine fed the widened actions.  Buffers must be page-locked and device-mapped
 * (cavgym_host_alloc, tensor.pin_memory()); scenarios of at most CAV_SMALL_M bodies; CAV_ESTATE otherwise. */
int cavgym_step_host_f32(CavEngine* engine, const float* actions, float* state_out, float* reward_out, uint8_t* done_out,
                         int32_t* winner_out, uint8_t* tangent_flag_out);

/* CAVEnv.reset (environment.py:225-229) with HOST buffers, for callers that hold no device memory (the host-buffer twin
 * of cavgym_reset, as cavgym_step_host is of cavgym_step): mask host u8[N] or NULL, init_state host real[M][4][N] or NULL
 * are copied in; the post-reset state is copied to state_out (host real[M][4][N], nullable).  Synchronises. */
int cavgym_reset_host(CavEngine* engine, const uint8_t* mask, const void* init_state, void* state_out);

/* CAVEnv.info (environment.py:106-117) for the current state of every env, on demand (nothing on the step path reads it):
 * polygons_out   device real[M][8][N]: body_polygons, x of rear_left, front_left, front_right, rear_right, then y
 *                (bodies.py:116-117; PelicanCrossing: its static box), nullable;
 * road_angle_out device real[M][N]: road_angles — DynamicBody.line_anchor_relative_angle towards the major road's centre line
 *                (bodies.py:206-212), NaN where the reference gives None (the body's box intersects the major road), nullable. */
int cavgym_info(CavEngine* engine, void* polygons_out, void* road_angle_out, cudaStream_t stream);

/* ---- accounting ------------------------------------------------------------------ */

/* reporting.analyse_episode / analyse_run (reporting.py:227-269) sums over episodes
 * scored so far; out is a HOST int64[CAV_N_STATS].  Synchronises. */
int cavgym_stats(CavEngine* engine, int64_t* out);
int cavgym_error_count(CavEngine* engine, int64_t* out);   /* envs with the invalid-action flag set */
int cavgym_launch_count(CavEngine* engine, int64_t* out);  /* kernels launched by this engine so far */

/* Per-episode rows (Simulation.run's episode.log, simulation.py:103-105): with capacity > 0 every episode scored from
 * now on is also appended to a device ring of `capacity` rows (0 switches it off and frees nothing).  An episode that finds
 * the ring full is still counted in cavgym_stats but its row is dropped; drain often enough. */
int cavgym_set_episode_log(CavEngine* engine, int64_t capacity);
/* Copies the rows appended since the last drain to `out` (HOST, room for max_rows) in append order, empties the ring,
 * *n_rows = rows copied, *dropped = rows lost to a full ring or to max_rows.  Synchronises. */
int cavgym_drain_episodes(CavEngine* engine, CavEpisodeRow* out, int64_t max_rows, int64_t* n_rows, int64_t* dropped);

/* ---- engine-owned device buffers (zero-copy views for the host shim) --------------- */
void* cavgym_state_ptr(CavEngine* engine);          /* real[M][4][N] */
int32_t* cavgym_liveness_ptr(CavEngine* engine);    /* i32[M][N] episode_liveness */
void* cavgym_agent_state_ptr(CavEngine* engine);    /* real[M][5][N] CrossingAgent state, NaN = None */
void* cavgym_action_ptr(CavEngine* engine);         /* real[M][2][N] last joint action taken */
int32_t* cavgym_timestep_ptr(CavEngine* engine);    /* i32[N] timesteps into the current episode */
uint8_t* cavgym_done_ptr(CavEngine* engine);        /* u8[N] latched done */
int32_t* cavgym_winner_ptr(CavEngine* engine);      /* i32[N] latched winner */
uint8_t* cavgym_error_ptr(CavEngine* engine);       /* u8[N] invalid-action flag */

/* ---- knobs ----------------------------------------------------------------------- */

/* Replace the Philox draws by caller-provided [0,1) numbers (replay of the
 * reference's MT19937 draws): device f64[M][CAV_DRAWS][N] read by every later
 * step, NULL restores Philox.  Slot 0 = epsilon test (template.py:61-62), slots
 * 1,2 = Box.sample components / Discrete.sample. */
int cavgym_set_uniform_override(CavEngine* engine, const double* uniforms);

/* Replace the five spawn draws (box, triangle, u, v, orientation) of every
 * SpawnPedestrian by caller-provided [0,1) numbers: device f64[M][5][N], NULL restores Philox. */
int cavgym_set_spawn_override(CavEngine* engine, const double* draws);

/* cavgym_action_ptr normally holds only what a RandomAgent must remember; with logging on,
 * every on-device agent's chosen action is stored there each step (parity tests, compat view). */
int cavgym_set_action_logging(CavEngine* engine, int enabled);

/* cavgym_step with replayed actions normally runs the persistent TMA-staged kernel (whole 128-env tiles, 16-byte
 * aligned buffers) and the plain thread-per-env kernel for whatever is left; use_tma = 0 forces the plain kernel
 * everywhere (A/B measurements, parity of the two paths). */
int cavgym_set_step_path(CavEngine* engine, int use_tma);

/* cavgym_rollout of a heterogeneous scenario of 3..CAV_SMALL_M bodies (crossroads, bus stop, pelican crossing:
 * examples/environments/*.py) has a second implementation: a team of M warps per 32 environments, one warp per body, bodies
 * in registers and meeting through shared memory (kernels_team.cuh).  Results are bitwise those of the thread-per-environment
 * kernel; measured it is no faster, so it is opt-in: team = 1 selects it, 0 (default) goes back.  Scenarios of
 * the Pedestrians-v0 family, one- and two-body scenarios and election agents always run thread per environment. */
int cavgym_set_rollout_path(CavEngine* engine, int team);

/* Scenarios with more than CAV_SMALL_M bodies always run the warp-per-environment kernels (one env per warp, bodies
 * staged in shared memory, fp32 broad phase + exact narrow phase for the all-pairs collision test of
 * environment.py:156-177); force = 1 selects them for a small scenario too (parity of the two paths), 0 goes back. */
int cavgym_set_dense_path(CavEngine* engine, int force);

/* cavgym_step_host with pinned (page-locked, mapped) buffers runs as one launch that reads and writes host memory
 * directly over PCIe; zero_copy = 0 forces the staged path (chunked async copies through device buffers), which is also
 * what pageable buffers get.  Tuning values: 2..8 = zero copy with (zero_copy - 1) resident CTAs per SM. */
int cavgym_set_host_path(CavEngine* engine, int zero_copy);

/* CAVEnv.current_timestep (environment.py:90,222) is never reset by the reference;
 * the engine keeps ONE counter of step calls for the whole batch. */
int cavgym_set_global_timestep(CavEngine* engine, int64_t t);

/* Near-tangent tolerance in pixels (default 1e-7 for f64, 5e-2 for f32). */
int cavgym_set_tangent_tolerance(CavEngine* engine, double tau);

/* ---- stand-alone kernels (per-function parity and the Body.step plugin hook) --------- */

/* DynamicBody.step (bodies.py:214-275) for n independent bodies of one type:
 * state real[4][n] in place, actions real[2][n], device pointers. */
int cavgym_bodies_step(const CavBodyType* type, void* state, const void* actions, int64_t n, double time_resolution,
                       int dtype, cudaStream_t stream);

/* Shape.intersects / contains / percentage_intersects (geometry.py:74-87) for n quad pairs:
 * quads_a, quads_b real[8][n] (x0..x3, y0..y3); out real[4][n] = a.intersects(b), b.contains(a),
 * a.percentage_intersects(b), near-tangent flag. */
int cavgym_geometry_probe(const void* quads_a, const void* quads_b, void* out, int64_t n, int dtype, cudaStream_t stream);

/* DynamicBody.stopping_zones (bodies.py:122-135, geometry.py:176-191) for n independent bodies of one type, from the
 * representation the step kernels test against (braking and reaction rectangles laid end to end in the body's frame):
 * state real[4][n], steering real[n] (the body's current steering angle) -> zones real[16][n] = braking x0..x3, y0..y3,
 * reaction x0..x3, y0..y3 in the reference's corner order (rear left, front left, front right, rear right), and
 * have u8[n] = 0 where the reference returns (None, None): total distance 0 or steering != 0. */
int cavgym_zones_probe(const CavBodyType* type, const void* state, const void* steering, void* zones_out, uint8_t* have_out,
                       int64_t n, int dtype, cudaStream_t stream);

const char* cavgym_last_error(void);
/* Page-locked, device-mapped host memory for the buffers of cavgym_step_host / cavgym_reset_host (what a caller without
 * PyTorch uses instead of tensor.pin_memory()).  write_combined != 0 allocates write-combined memory: meant for buffers the
 * host only WRITES and the GPU reads (the joint actions) — GPU reads of it do not snoop the CPU caches; host reads of it are
 * slow.  Free with cavgym_host_free. */
int cavgym_host_alloc(size_t bytes, int write_combined, void** out);
int cavgym_host_free(void* ptr);

const char* cavgym_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CAVGYM_H */
