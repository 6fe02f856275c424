// kernels_tma.cuh — the two kernels that run on replayed joint actions (every body CAV_AGENT_EXTERNAL), written as
// warp-specialised TMA pipelines:
//
//   step_tma_kernel     cavgym_step    one CAVEnv.step (environment.py:119-223) over N envs; persistent CTAs walk
//                                      128-env tiles; the pipeline runs over TILES
//   replay_tma_kernel   cavgym_replay  T fused steps with the state in registers; one CTA per 224-env tile; the
//                                      pipeline runs over TIME STEPS (actions in, trajectory rows out)
//
// Why: a step moves ~200 B per env through HBM and keeps ~20 values per env live while it computes.  In the plain
// thread-per-env kernels (kernels_small.cuh) those values sit in registers from the moment their loads are issued,
// every access costs a 64-bit address computation, and DRAM latency can only be hidden by more resident warps — which
// the register footprint forbids (ncu: 22 % occupancy, long-scoreboard the top stall).  Here instead
//   * one PRODUCER warp per CTA moves data and does no arithmetic: it brings a tile's SoA rows into shared memory with
//     one `cp.async.bulk` per row (TMA engine, SASS UBLKCP) — the env index is the fastest axis, so each row of a tile
//     is one contiguous segment (1 KiB in fp64) — completing on an mbarrier, and sends result rows back to HBM with
//     bulk stores (shared -> global);
//   * CONSUMER warps (one env per thread) wait on that mbarrier, read their env from shared memory (conflict-free:
//     lane i touches word i), run the same `transition` as every other kernel, write results to shared memory and
//     hand the stage back with one mbarrier arrival per warp — no block-wide barrier anywhere in the loop;
//   * kTmaStages stages are in flight, so loads for later tiles / steps land while the current one is computed and no
//     register is ever held for a load in flight.
// Per-thread global addressing is left only on the rare paths (heading cache, liveness and done latches, errors).
//
// Alignment (checked by the host, otherwise the plain kernels run): every row segment must start and end on 16 bytes.
// With n % 16 == 0 a ragged last tile is handled here too (shorter bulk copies); with only n % 4 == 0 the step kernel
// takes the whole tiles and leaves the tail to the plain kernel.
#pragma once
#include <atomic>

#include "kernels_small.cuh"

namespace cav {

#ifndef CAV_TMA_STEP_WARPS
#define CAV_TMA_STEP_WARPS 4       // consumer warps per CTA in step_tma_kernel: 128-env tiles
#endif
// replay_tma_kernel: a launch is ONE wave of long-running CTAs, so the tile size decides whether every env is resident
// at once.  The fp64 transition with the state in registers needs 128 registers; 7 consumer warps + the producer are
// 256 threads, 2 CTAs per SM use the whole register file, and 65,536 envs are 293 tiles of 224 <= 296 slots.
#ifndef CAV_TMA_REPLAY_WARPS
#define CAV_TMA_REPLAY_WARPS 7
#endif
#ifndef CAV_MIN_BLOCKS_REPLAY
#define CAV_MIN_BLOCKS_REPLAY 2
#endif
#ifndef CAV_MIN_BLOCKS_TMA
#define CAV_MIN_BLOCKS_TMA 3
#endif
#ifndef CAV_TMA_STAGES
#define CAV_TMA_STAGES 3
#endif
constexpr int kTmaStages = CAV_TMA_STAGES;
#ifndef CAV_REPLAY_STAGES
#define CAV_REPLAY_STAGES CAV_TMA_STAGES
#endif
constexpr int kReplayStages = CAV_REPLAY_STAGES;   // pipeline depth of replay_tma_kernel (over time steps)
constexpr int kMaxSmemPerBlock = 227 * 1024;   // opt-in dynamic shared memory per CTA on sm_100
constexpr int kMaxDevices = 64;                // per-device caches of the launchers below
// The TMA-staged kernels pay off while three CTAs (step) / two 256-thread CTAs (replay) fit an SM, i.e. for one or two
// bodies.  From three bodies on the staging buffers crowd the SM and the plain kernels are faster — measured at 1M envs,
// fp64, step: M=3 0.135 vs 0.111 ms, M=4 0.353 vs 0.170, M=5 0.616 vs 0.243, bus stop 0.97 vs 0.47; replay of 50 steps at
// 65,536 envs: bus stop 2.76 vs 1.74 ms (scripts/profile_replay.py) — so the launchers hand those scenarios to them.
#ifndef CAV_TMA_MAX_BODIES
#define CAV_TMA_MAX_BODIES 2
#endif
constexpr int kStepWarps = CAV_TMA_STEP_WARPS, kStepTile = 32 * kStepWarps;
constexpr int kReplayWarps = CAV_TMA_REPLAY_WARPS, kReplayTile = 32 * kReplayWarps;

#ifndef CAV_MBAR_SUSPEND_NS
#define CAV_MBAR_SUSPEND_NS 0
#endif
// ---------------------------------------------------------------- PTX: mbarrier and bulk asynchronous copies
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
#if CAV_MBAR_SUSPEND_NS > 0
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"   // suspend (no issue slots) up to the hint, woken on completion
#else
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
#endif
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
#if CAV_MBAR_SUSPEND_NS > 0
        , "r"((uint32_t)CAV_MBAR_SUSPEND_NS)
#endif
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared, completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make this thread's shared-memory writes (generic proxy) visible to the bulk-copy engine (async proxy)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Programmatic dependent launch (launch_replay_tma): `launch_dependents` lets the next kernel of the stream start its CTAs
// as soon as SM resources free up; `wait` blocks until the previous grid has completed and its writes are visible.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Launch chaining of replay_tma_kernel.  Consecutive cavgym_replay launches over the same tiles depend on each other tile by
// tile only: CTA i of a launch needs the env state CTA i of the previous launch wrote, nothing else.  So instead of waiting
// for the whole previous grid (griddepcontrol.wait: its slowest CTA, its drain and the ramp of this one — a third of a 20-step
// launch), a chained launch waits for ITS tile's sequence number, published with release semantics by the CTA that finished
// the tile, and the launches of a train overlap like the iterations of one persistent kernel.  No deadlock: a grid only
// starts once every CTA of the grid before it is resident or done (that is when its launch_dependents resolves), so the CTA
// waited for is always running or finished.  A wait that outlasts ten seconds traps instead of hanging the device.
__device__ __forceinline__ int32_t ld_acquire_gpu(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int32_t* p, int32_t v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void wait_tile_generation(const int32_t* gen, int32_t want) {
  if (ld_acquire_gpu(gen) == want) return;
  const unsigned long long start = global_timer_ns();
  while (ld_acquire_gpu(gen) != want) {
    __nanosleep(200);
    if (global_timer_ns() - start > 10000000000ull) __trap();
  }
}

// One SoA row as the producer sees it: where its segment for unit 0 (tile 0 / step 0) starts in HBM, how far the next
// unit's segment is, where it lives in a stage, and the element size (segment bytes = envs in the tile * elem).
struct TmaRow {
  unsigned long long gbase, unit_stride;
  uint32_t smem_off, elem;
};

// ================================================================ cavgym_step
template <typename R, int M>
struct StepLayout {
  static constexpr int T = kStepTile;
  static constexpr int kRow = T * (int)sizeof(R);   // one R row of a tile
  static constexpr int kRowI = T * 4;               // one int32 row
  static constexpr int kLiveRows = M > 1 ? M - 1 : 1;
  static constexpr int oState = 0;                            // R [M*4][T]   in / out (in place)
  static constexpr int oCs = oState + M * 4 * kRow;           // R [M*2][T]   in
  static constexpr int oAct = oCs + M * 2 * kRow;             // R [M*2][T]   in
  static constexpr int oReward = oAct + M * 2 * kRow;         // R [M][T]     out
  static constexpr int oTep = oReward + M * kRow;             // i32 [T]      in / out (in place)
  static constexpr int oLive = oTep + kRowI;                  // i32 [M-1][T] in   (bodies 1..)
  static constexpr int oWinner = oLive + kLiveRows * kRowI;   // i32 [T]      in (latch) / out
  static constexpr int oDone = oWinner + kRowI;               // u8 [T]       in   (latched done)
  static constexpr int oDoneOut = oDone + T;                  // u8 [T]       out
  static constexpr int oTangent = oDoneOut + T;               // u8 [T]       out
  static constexpr int kStageBytes = oTangent + T;            // multiple of 128
  static constexpr int kMaxRows = M * 4 * 2 + M * 2 * 2 + M + 8;
  static constexpr int kBarOffset = kTmaStages * kStageBytes;
  static constexpr int kTableOffset = kBarOffset + 128;
  static constexpr int kSmemBytes = kTableOffset + 2 * kMaxRows * (int)sizeof(TmaRow);
};

template <typename R, int M, bool GENERIC>
__global__ void __launch_bounds__(kStepTile + 32, CAV_MIN_BLOCKS_TMA) step_tma_kernel(
    const __grid_constant__ DevScenario<R> sc, const __grid_constant__ EnvBuffers<R> buf, const __grid_constant__ StepIO<R> io,
    int64_t t_global, int64_t n_tiles) {
  using L = StepLayout<R, M>;
  constexpr int T = kStepTile;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);   // [stages] tile landed (producer -> consumers)
  uint64_t* done = full + kTmaStages;                                    // [stages] tile computed (consumers -> producer)
  // [stages] bit k: state row k (body k / 4, component k % 4) changed for some env of the tile.  A row nobody changed — the
  // velocity and heading of a body that neither accelerates nor turns: three of the eight rows in the stock scenario — is
  // not written back to the engine's state.
  uint32_t* changed = reinterpret_cast<uint32_t*>(smem + L::kBarOffset + 64);
  TmaRow* in_rows = reinterpret_cast<TmaRow*>(smem + L::kTableOffset);
  TmaRow* out_rows = in_rows + L::kMaxRows;
  __shared__ int n_in_s, n_out_s, in_elems_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n = buf.n, lo = buf.lo;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], kStepWarps); changed[s] = 0u; }
    mbar_fence_init();
    // row tables (a few dozen entries, once per CTA)
    int ni = 0, no = 0, elems = 0;
    auto in = [&](const void* base, int64_t row, int elem, int off) {
      in_rows[ni++] = {(unsigned long long)base + (unsigned long long)((row * n + lo) * elem), (unsigned long long)(T * elem),
                       (uint32_t)off, (uint32_t)elem};
      elems += elem;
    };
    auto out = [&](void* base, int64_t row, int elem, int off) {
      out_rows[no++] = {(unsigned long long)base + (unsigned long long)((row * n + lo) * elem), (unsigned long long)(T * elem),
                        (uint32_t)off, (uint32_t)elem};
    };
    for (int k = 0; k < M * 4; ++k) in(buf.state, k, sizeof(R), L::oState + k * L::kRow);
    for (int k = 0; k < M * 2; ++k) in(buf.cs, k, sizeof(R), L::oCs + k * L::kRow);
    for (int k = 0; k < M * 2; ++k) in(io.actions, k, sizeof(R), L::oAct + k * L::kRow);
    in(buf.t_ep, 0, 4, L::oTep);
    for (int b = 1; b < M; ++b) in(buf.liveness, b, 4, L::oLive + (b - 1) * L::kRowI);
    in(buf.done, 0, 1, L::oDone);
    in(buf.winner, 0, 4, L::oWinner);   // latched winner of finished envs: reported again while they stay frozen
    for (int k = 0; k < M * 4; ++k) out(buf.state, k, sizeof(R), L::oState + k * L::kRow);
    if (io.state_out && io.state_out != buf.state)
      for (int k = 0; k < M * 4; ++k) out(io.state_out, k, sizeof(R), L::oState + k * L::kRow);
    if (io.reward_out)
      for (int b = 0; b < M; ++b) out(io.reward_out, b, sizeof(R), L::oReward + b * L::kRow);
    out(buf.t_ep, 0, 4, L::oTep);
    if (io.winner_out) out(io.winner_out, 0, 4, L::oWinner);
    if (io.done_out) out(io.done_out, 0, 1, L::oDoneOut);
    if (io.tangent_out) out(io.tangent_out, 0, 1, L::oTangent);
    n_in_s = ni; n_out_s = no; in_elems_s = elems;
  }
  __syncthreads();
  const int n_in = n_in_s, n_out = n_out_s;
  const uint32_t in_elems = (uint32_t)in_elems_s;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  auto envs_in = [&](int64_t tile) { return (uint32_t)((buf.hi - lo - tile * T) < T ? (buf.hi - lo - tile * T) : T); };

  if (warp == kStepWarps) {
    // ================= producer warp: HBM -> shared (bulk loads), shared -> HBM (bulk stores); no arithmetic
    auto issue_loads = [&](int s, int64_t tile) {
      unsigned char* st = smem + s * L::kStageBytes;
      const uint32_t cnt = envs_in(tile);
      if (lane == 0) mbar_expect_tx(&full[s], in_elems * cnt);
      __syncwarp();
      for (int r = lane; r < n_in; r += 32) {
        const TmaRow row = in_rows[r];
        bulk_load(st + row.smem_off, reinterpret_cast<const void*>(row.gbase + (unsigned long long)tile * row.unit_stride),
                  row.elem * cnt, &full[s]);
      }
    };
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) {
      const int64_t tile = first + (int64_t)s * stride;
      if (tile < n_tiles) issue_loads(s, tile);
    }
    int it = 0;
    for (int64_t tile = first; tile < n_tiles; tile += stride, ++it) {
      const int s = it % kTmaStages;
      const uint32_t parity = (uint32_t)(it / kTmaStages) & 1u;
      unsigned char* st = smem + s * L::kStageBytes;
      const uint32_t cnt = envs_in(tile);
      mbar_wait(&done[s], parity);   // every consumer warp has written its results for this tile
      const uint32_t dirty = changed[s];   // (the consumers' atomicOr precedes their arrival on done[s])
      for (int r = lane; r < n_out; r += 32) {
        if (r < M * 4 && !(dirty >> r & 1u)) continue;   // the first M * 4 output rows are the engine's state, in place
        const TmaRow row = out_rows[r];
        bulk_store(reinterpret_cast<void*>(row.gbase + (unsigned long long)tile * row.unit_stride), st + row.smem_off, row.elem * cnt);
      }
      bulk_commit();
      __syncwarp();
      if (lane == 0) changed[s] = 0u;      // before the stage is handed out again (ordered by the arrival in issue_loads)
      const int64_t next = tile + (int64_t)kTmaStages * stride;
      if (next < n_tiles) {
        bulk_wait_read_all();   // the stores above have read the stage: it may be overwritten
        __syncwarp();
        issue_loads(s, next);
      }
    }
    bulk_wait_read_all();   // shared memory must outlive the last bulk stores
    return;
  }

  // ================= consumer warps: one env per thread, no block-wide synchronisation
  int it = 0;
  for (int64_t tile = first; tile < n_tiles; tile += stride, ++it) {
    const int s = it % kTmaStages;
    const uint32_t parity = (uint32_t)(it / kTmaStages) & 1u;
    unsigned char* st = smem + s * L::kStageBytes;
    mbar_wait(&full[s], parity);
    uint32_t mine_changed = 0u;
    if ((uint32_t)tid < envs_in(tile)) {
      // ---- this thread's env: shared memory -> registers
      const int64_t e = lo + tile * T + tid;
      R* sS = reinterpret_cast<R*>(st + L::oState);
      const R* sC = reinterpret_cast<const R*>(st + L::oCs);
      const R* sA = reinterpret_cast<const R*>(st + L::oAct);
      R* sRw = reinterpret_cast<R*>(st + L::oReward);
      int32_t* sT = reinterpret_cast<int32_t*>(st + L::oTep);
      const int32_t* sL = reinterpret_cast<const int32_t*>(st + L::oLive);
      EnvRegs<R, M> env;
      R ext[M][2];
      env.done = st[L::oDone + tid];
      env.t_ep = sT[tid];
      env.winner = -1;
      env.episode = 0;
      env.ag_dirty = 0; env.cs_dirty = 0; env.live_dirty = 0;
      env.live[0] = 0;
#pragma unroll
      for (int b = 0; b < M; ++b) {
#pragma unroll
        for (int c = 0; c < 4; ++c) env.s[b][c] = sS[(b * 4 + c) * T + tid];
        env.cs[b][0] = sC[(b * 2 + 0) * T + tid];
        env.cs[b][1] = sC[(b * 2 + 1) * T + tid];
        ext[b][0] = sA[(b * 2 + 0) * T + tid];
        ext[b][1] = sA[(b * 2 + 1) * T + tid];
        env.held[b][0] = R(0); env.held[b][1] = R(0);
        if (b > 0) env.live[b] = sL[(b - 1) * T + tid];
      }

      // ---- one transition (same bookkeeping as `advance` in kernels_small.cuh)
      StepResult<R, M> res;
      if (env.done) {  // frozen until reset
#pragma unroll
        for (int b = 0; b < M; ++b) res.reward[b] = R(0);
        res.terminate = env.done == 1;
        res.winner = reinterpret_cast<const int32_t*>(st + L::oWinner)[tid];
        res.tangent = false;
      } else {
        // the new state goes back to shared memory as soon as the bodies have moved
        auto moved = [&](const EnvRegs<R, M>& now) {
#pragma unroll
          for (int b = 0; b < M; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const R was = sS[(b * 4 + c) * T + tid];
              if (!(was == now.s[b][c])) mine_changed |= 1u << (b * 4 + c);
              sS[(b * 4 + c) * T + tid] = now.s[b][c];
            }
        };
        transition<R, M, false, GENERIC>(sc, buf, e, t_global, env, ext, res, moved);
        if (res.invalid) buf.err[e] = 1;
        if (res.tangent) count_tangent(buf.stats);
        if (env.done) {
          score_episode<R, M>(buf, env, e);
          buf.done[e] = env.done;
          buf.winner[e] = env.winner;
        }
#pragma unroll
        for (int b = 0; b < M; ++b) {
          if (env.cs_dirty >> b & 1u) {   // heading changed: refresh the cached cos/sin (rare)
            buf.cs[((int64_t)b * 2 + 0) * n + e] = env.cs[b][0];
            buf.cs[((int64_t)b * 2 + 1) * n + e] = env.cs[b][1];
          }
          if (b > 0 && (env.live_dirty >> b & 1u)) buf.liveness[(int64_t)b * n + e] = env.live[b];
        }
        sT[tid] = env.t_ep;
      }
#pragma unroll
      for (int b = 0; b < M; ++b) sRw[b * T + tid] = res.reward[b];
      reinterpret_cast<int32_t*>(st + L::oWinner)[tid] = res.winner;
      st[L::oDoneOut + tid] = res.terminate ? 1 : 0;
      st[L::oTangent + tid] = res.tangent ? 1 : 0;
    }
    // ---- hand the tile to the producer: writes visible to the bulk-copy engine, one arrival per warp
    mine_changed = __reduce_or_sync(0xffffffffu, mine_changed);
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (mine_changed) atomicOr(&changed[s], mine_changed);
      mbar_arrive(&done[s]);
    }
  }
}

// ================================================================ cavgym_replay
//   full[s]   producer -> consumers   actions of step t (stage s = t % S) have landed            (tx count)
//   done[s]   consumers -> producer   step t computed: actions consumed, trajectory rows written (one arrival per warp)
//   empty[s]  producer -> consumers   the bulk stores of the step that last used stage s have read it
template <typename R, int M>
struct ReplayLayout {
  static constexpr int T = kReplayTile;
  static constexpr int kRow = T * (int)sizeof(R);
  static constexpr int oAct = 0;                          // R [M*2][T]   in
  static constexpr int oState = oAct + M * 2 * kRow;      // R [M*4][T]   out
  static constexpr int oReward = oState + M * 4 * kRow;   // R [M][T]     out
  static constexpr int oWinner = oReward + M * kRow;      // i32 [T]      out
  static constexpr int oDoneOut = oWinner + T * 4;        // u8 [T]       out
  static constexpr int oTangent = oDoneOut + T;           // u8 [T]       out
  static constexpr int kStageBytes = ((oTangent + T + 127) / 128) * 128;
  static constexpr int kMaxRows = M * 4 + M * 2 + M + 4;
  static constexpr int kBarOffset = kReplayStages * kStageBytes;
  static constexpr int kTableOffset = kBarOffset + 128;
  static constexpr int kSmemBytes = kTableOffset + 2 * kMaxRows * (int)sizeof(TmaRow);
};

#ifdef CAV_REPLAY_MAXNREG
template <typename R, int M, bool GENERIC>
__global__ void __maxnreg__(CAV_REPLAY_MAXNREG) replay_tma_kernel(
#else
template <typename R, int M, bool GENERIC>
__global__ void __launch_bounds__(kReplayTile + 32, CAV_MIN_BLOCKS_REPLAY) replay_tma_kernel(
#endif

    const __grid_constant__ DevScenario<R> sc, const __grid_constant__ EnvBuffers<R> buf, const __grid_constant__ StepIO<R> io,
    int64_t t_global, int n_steps) {
  using L = ReplayLayout<R, M>;
  constexpr int T = kReplayTile;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* done = full + kReplayStages;
  uint64_t* empty = done + kReplayStages;
  TmaRow* in_rows = reinterpret_cast<TmaRow*>(smem + L::kTableOffset);
  TmaRow* out_rows = in_rows + L::kMaxRows;
  __shared__ int n_out_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n = buf.n;
  const int64_t tile_lo = buf.lo + (int64_t)blockIdx.x * T;   // first env of this CTA's tile
  const uint32_t cnt = (uint32_t)((buf.hi - tile_lo) < T ? (buf.hi - tile_lo) : T);
  constexpr int n_in = M * 2;
  pdl_launch_dependents();

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kReplayStages; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], kReplayWarps); mbar_init(&empty[s], 1); }
    mbar_fence_init();
    int no = 0;
    auto out = [&](void* base, int64_t rows_per_step, int64_t row, int elem, int off) {
      out_rows[no++] = {(unsigned long long)base + (unsigned long long)((row * n + tile_lo) * elem),
                        (unsigned long long)(rows_per_step * n * elem), (uint32_t)off, (uint32_t)elem};
    };
    for (int k = 0; k < n_in; ++k)
      in_rows[k] = {(unsigned long long)io.actions + (unsigned long long)((k * n + tile_lo) * (int64_t)sizeof(R)),
                    (unsigned long long)((int64_t)M * 2 * n * (int64_t)sizeof(R)), (uint32_t)(L::oAct + k * L::kRow), (uint32_t)sizeof(R)};
    if (io.state_out) for (int k = 0; k < M * 4; ++k) out(io.state_out, M * 4, k, sizeof(R), L::oState + k * L::kRow);
    if (io.reward_out) for (int b = 0; b < M; ++b) out(io.reward_out, M, b, sizeof(R), L::oReward + b * L::kRow);
    if (io.winner_out) out(io.winner_out, 1, 0, 4, L::oWinner);
    if (io.done_out) out(io.done_out, 1, 0, 1, L::oDoneOut);
    if (io.tangent_out) out(io.tangent_out, 1, 0, 1, L::oTangent);
    n_out_s = no;
  }
  __syncthreads();
  const int n_out = n_out_s;

  if (warp == kReplayWarps) {
    // ================= producer warp
    auto issue_loads = [&](int s, int t) {
      unsigned char* st = smem + s * L::kStageBytes;
      if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)(n_in * sizeof(R)) * cnt);
      __syncwarp();
      for (int r = lane; r < n_in; r += 32) {
        const TmaRow row = in_rows[r];
        bulk_load(st + row.smem_off, reinterpret_cast<const void*>(row.gbase + (unsigned long long)t * row.unit_stride), row.elem * cnt,
                  &full[s]);
      }
    };
    for (int t = 0; t < kReplayStages && t < n_steps; ++t) issue_loads(t, t);
    for (int t = 0; t < n_steps; ++t) {
      const int s = t % kReplayStages;
      const uint32_t parity = (uint32_t)(t / kReplayStages) & 1u;
      unsigned char* st = smem + s * L::kStageBytes;
      mbar_wait(&done[s], parity);
      for (int r = lane; r < n_out; r += 32) {
        const TmaRow row = out_rows[r];
        bulk_store(reinterpret_cast<void*>(row.gbase + (unsigned long long)t * row.unit_stride), st + row.smem_off, row.elem * cnt);
      }
      bulk_commit();
      if (t + kReplayStages < n_steps) issue_loads(s, t + kReplayStages);   // the action rows of stage s were consumed
      // every lane waits until all but the newest S - 2 of its store groups have read shared memory; then the stage of
      // step t - (S - 2) may be written again
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kReplayStages >= 2 ? kReplayStages - 2 : 0) : "memory");
      __syncwarp();
      const int freed = t - (kReplayStages >= 2 ? kReplayStages - 2 : 0);
      if (freed >= 0 && lane == 0) mbar_arrive(&empty[freed % kReplayStages]);
    }
    if (io.tile_gen) {
      // the tile is published below: its trajectory rows must have LANDED by then (a launch three slabs later may write the
      // same rows of a rotating slab), not just have left shared memory
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(kReplayTile + 32) : "memory");
    } else {
      bulk_wait_read_all();
    }
    return;
  }

  // ================= consumer warps
  // The set-up above and the producer's first action loads do not depend on the previous launch and overlap its tail; the
  // env state does (the previous launch wrote it): read it once the programmatic dependency has resolved.  Everything this
  // kernel writes follows from here (the producer's stores wait for the consumers' first step).
  if (io.chained) wait_tile_generation(io.tile_gen + blockIdx.x, io.seq - 1);   // this tile's state is in place (see above)
  else pdl_wait();
  const bool active = (uint32_t)tid < cnt;
  const int64_t e = tile_lo + (active ? tid : 0);
  EnvRegs<R, M> env;
  load_env<R, M, false>(sc, buf, e, env);
  const bool was_live = env.done == 0;
  for (int t = 0; t < n_steps; ++t) {
    const int s = t % kReplayStages;
    const uint32_t parity = (uint32_t)(t / kReplayStages) & 1u;
    unsigned char* st = smem + s * L::kStageBytes;
    mbar_wait(&full[s], parity);
    StepResult<R, M> res;
    if (active) {
      const R* sA = reinterpret_cast<const R*>(st + L::oAct);
      R ext[M][2];
#pragma unroll
      for (int b = 0; b < M; ++b) { ext[b][0] = sA[(b * 2 + 0) * T + tid]; ext[b][1] = sA[(b * 2 + 1) * T + tid]; }
      if (env.done) {  // frozen until reset
#pragma unroll
        for (int b = 0; b < M; ++b) res.reward[b] = R(0);
        res.terminate = env.done == 1;
        res.winner = env.winner;
        res.tangent = false;
      } else {
        transition<R, M, false, GENERIC>(sc, buf, e, t_global + t, env, ext, res);
        if (res.invalid) buf.err[e] = 1;
        if (res.tangent) count_tangent(buf.stats);
        if (env.done) score_episode<R, M>(buf, env, e);
      }
    }
    // the stage's output rows are free once the stores of the step that used it last have read them
    mbar_wait(&empty[s], parity ^ 1u);
    if (active && n_out > 0) {
      R* sS = reinterpret_cast<R*>(st + L::oState);
      R* sRw = reinterpret_cast<R*>(st + L::oReward);
#pragma unroll
      for (int b = 0; b < M; ++b) {
#pragma unroll
        for (int c = 0; c < 4; ++c) sS[(b * 4 + c) * T + tid] = env.s[b][c];
        sRw[b * T + tid] = res.reward[b];
      }
      reinterpret_cast<int32_t*>(st + L::oWinner)[tid] = res.winner;
      st[L::oDoneOut + tid] = res.terminate ? 1 : 0;
      st[L::oTangent + tid] = res.tangent ? 1 : 0;
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(&done[s]);
  }
  if (active && was_live) store_env<R, M, false>(sc, buf, e, env, false);
  if (io.tile_gen) {   // publish the tile: every consumer's state is stored, then one release store of this launch's number
    asm volatile("bar.sync 1, %0;" ::"n"(kReplayTile + 32) : "memory");   // with the producer warp: its bulk stores have landed
    if (tid == 0) { __threadfence(); st_release_gpu(io.tile_gen + blockIdx.x, io.seq); }
  }
}

// ================================================================ host side
// Which part of [lo, hi) can a TMA kernel with `tile`-env tiles take?  Returns the number of envs (0 = none).
// Every bulk copy must start and end on 16 bytes: rows of `elem`-byte words start at (row * n + lo + k * tile) * elem.
template <typename R>
inline int64_t tma_span(const EnvBuffers<R>& buf, const StepIO<R>& io, int tile, bool u8_rows_strided) {
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!io.actions || buf.n % 4 != 0 || buf.lo % 16 != 0) return 0;
  if (!aligned(io.actions) || !aligned(io.state_out) || !aligned(io.reward_out) || !aligned(io.done_out) || !aligned(io.winner_out) ||
      !aligned(io.tangent_out))
    return 0;
  const int64_t span = buf.hi - buf.lo;
  if (buf.n % 16 == 0 && span % 16 == 0) return span;   // ragged last tile handled in the kernel
  if (u8_rows_strided) return 0;                         // u8 trajectory rows advance by n bytes per step
  return span / tile * tile;
}

template <typename R, int M>
bool launch_step_tma(const DevScenario<R>& sc, const EnvBuffers<R>& buf, const StepIO<R>& io, int64_t t_global, cudaStream_t stream,
                     int64_t* envs_done) {
  using L = StepLayout<R, M>;
  *envs_done = 0;
  if constexpr (M > CAV_TMA_MAX_BODIES || L::kSmemBytes > kMaxSmemPerBlock) {
    return true;   // plain kernel (see CAV_TMA_MAX_BODIES); the TMA kernel is not even instantiated for this body count
  } else {
  const int64_t span = tma_span(buf, io, kStepTile, false);
  if (span == 0) return true;
  const int64_t tiles = (span + kStepTile - 1) / kStepTile;
  // per DEVICE and instantiation: the shared-memory opt-in is a per-device function attribute (a second engine on another
  // GPU of the same process needs its own), so the cache is keyed by the current device; 0 unknown, -1 unusable
  static std::atomic<int> resident[kMaxDevices][2];
  static std::atomic<int> sm_count[kMaxDevices];
  auto kernel = sc.homogeneous ? step_tma_kernel<R, M, false> : step_tma_kernel<R, M, true>;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return true;
  std::atomic<int>& slot = resident[dev][sc.homogeneous ? 0 : 1];
  int res = slot.load(std::memory_order_acquire);
  if (res < 0) return true;   // found unusable before (shared memory): plain kernel
  if (res == 0) {
    int sms = 0;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kSmemBytes) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, kernel, kStepTile + 32, L::kSmemBytes) != cudaSuccess || res < 1 ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      cudaGetLastError();     // not an error of the caller's: this body count's staging does not fit, the plain kernel runs
      slot.store(-1, std::memory_order_release);
      return true;
    }
    sm_count[dev].store(sms, std::memory_order_relaxed);
    slot.store(res, std::memory_order_release);   // idempotent: two threads racing here store the same values
  }
  const int sms = sm_count[dev].load(std::memory_order_relaxed);
  EnvBuffers<R> range = buf;
  range.hi = buf.lo + span;
  const int per_sm = io.ctas_per_sm > 0 && io.ctas_per_sm < res ? io.ctas_per_sm : res;
  const int64_t grid = tiles < (int64_t)sms * per_sm ? tiles : (int64_t)sms * per_sm;
  kernel<<<(unsigned)grid, kStepTile + 32, L::kSmemBytes, stream>>>(sc, range, io, t_global, tiles);
  *envs_done = span;
  return true;
  }
}

template <typename R, int M>
bool launch_replay_tma(const DevScenario<R>& sc, const EnvBuffers<R>& buf, const StepIO<R>& io, int64_t t_global, int n_steps,
                       cudaStream_t stream, int64_t* envs_done) {
  using L = ReplayLayout<R, M>;
  *envs_done = 0;
  if constexpr (M > CAV_TMA_MAX_BODIES || L::kSmemBytes > kMaxSmemPerBlock) {
    return true;   // plain kernel (see CAV_TMA_MAX_BODIES)
  } else {
  const int64_t span = tma_span(buf, io, kReplayTile, io.done_out != nullptr || io.tangent_out != nullptr);
  if (span == 0) return true;
  static std::atomic<int> ready[kMaxDevices][2];   // per device (see launch_step_tma): 0 unknown, 1 usable, -1 does not fit
  auto kernel = sc.homogeneous ? replay_tma_kernel<R, M, false> : replay_tma_kernel<R, M, true>;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return true;
  std::atomic<int>& slot = ready[dev][sc.homogeneous ? 0 : 1];
  const int ok = slot.load(std::memory_order_acquire);
  if (ok < 0) return true;
  if (ok == 0) {
    int blocks = 0;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kSmemBytes) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, kReplayTile + 32, L::kSmemBytes) != cudaSuccess || blocks < 1) {
      cudaGetLastError();
      slot.store(-1, std::memory_order_release);
      return true;          // plain replay kernel for this body count
    }
    slot.store(1, std::memory_order_release);
  }
  EnvBuffers<R> range = buf;
  range.hi = buf.lo + span;
  const int64_t tiles = (span + kReplayTile - 1) / kReplayTile;
  // Programmatic dependent launch: behind another replay launch in the stream this grid's CTAs start (barrier set-up, first
  // action loads) while that grid drains, and read the env state after griddepcontrol.wait; behind any other kernel the
  // attribute changes nothing.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)tiles);
  cfg.blockDim = dim3(kReplayTile + 32);
  cfg.dynamicSmemBytes = L::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kernel, sc, range, io, t_global, n_steps) != cudaSuccess) return false;
  *envs_done = span;
  return true;
  }
}

template <typename R, int M>
constexpr SmallLaunchers<R> make_launchers() {
  return {&launch_step<R, M>, &launch_replay<R, M>, &launch_rollout<R, M>, &launch_reset<R, M>, &launch_step_tma<R, M>,
          &launch_replay_tma<R, M>, &launch_step_wire32<R, M>};
}

}  // namespace cav
