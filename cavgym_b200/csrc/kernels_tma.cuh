// kernels_tma.cuh — the per-step kernel for replayed joint actions (cavgym_step with an `actions` buffer, every body
// CAV_AGENT_EXTERNAL) as a persistent, TMA-staged stream.
//
// One CAVEnv.step (environment.py:119-223) over N envs moves ~200 B per env through HBM and keeps ~20 values per
// env live while it computes.  In the plain thread-per-env kernel (kernels_small.cuh) those values sit in registers
// from the moment their loads are issued, every access costs a 64-bit address computation, and the only way to hide
// DRAM latency is more resident warps — which the register footprint forbids.  Here instead:
//
//   * CTAs are persistent (grid = SMs x resident CTAs) and walk tiles of 128 consecutive envs;
//   * warp 0 brings a tile's SoA rows (state 4M, cos/sin 2M, actions 2M, timestep, liveness, done) into shared
//     memory with one `cp.async.bulk` (TMA, SASS UBLKCP) per row — each row of a tile is one contiguous 1 KiB
//     (fp64) / 512 B (fp32) segment because the env index is the fastest axis — completing on an mbarrier;
//   * tiles are double-buffered: the rows of tile i+1 land while tile i is being computed, so DRAM latency is hidden
//     without any registers being held for loads in flight;
//   * each thread reads its env from shared memory (conflict-free: lane i touches word i), runs the same
//     `transition` as every other kernel, and writes the new state / reward / flags back to shared memory;
//   * warp 0 sends the output rows to HBM with bulk stores (shared -> global), again one per row.
// Per-thread global addressing is left only on the rare paths (heading cache, liveness and done latches, errors).
//
// Requirements checked by the host (otherwise the plain kernel runs): every row segment 16-byte aligned (n % 4 == 0,
// tile-aligned range, 16-byte aligned caller buffers).  A ragged last tile goes through the plain kernel.
#pragma once
#include "kernels_small.cuh"

namespace cav {

#ifndef CAV_TMA_CONSUMER_WARPS
#define CAV_TMA_CONSUMER_WARPS 4
#endif
#ifndef CAV_MIN_BLOCKS_TMA
#define CAV_MIN_BLOCKS_TMA 3
#endif
constexpr int kConsumerWarps = CAV_TMA_CONSUMER_WARPS;
constexpr int kTile = 32 * kConsumerWarps;   // envs per tile = consumer threads per CTA
constexpr int kTmaThreads = kTile + 32;      // + one producer warp that only moves data
#ifndef CAV_TMA_STAGES
#define CAV_TMA_STAGES 3
#endif
constexpr int kTmaStages = CAV_TMA_STAGES;

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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
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

// ---------------------------------------------------------------- shared-memory layout of one stage
template <typename R, int M>
struct TmaLayout {
  static constexpr int kRow = kTile * (int)sizeof(R);   // one R row of a tile
  static constexpr int kRowI = kTile * 4;               // one int32 row
  static constexpr int kLiveRows = M > 1 ? M - 1 : 1;
  static constexpr int oState = 0;                            // R [M*4][T]   in / out (in place)
  static constexpr int oCs = oState + M * 4 * kRow;           // R [M*2][T]   in
  static constexpr int oAct = oCs + M * 2 * kRow;             // R [M*2][T]   in
  static constexpr int oReward = oAct + M * 2 * kRow;         // R [M][T]     out
  static constexpr int oTep = oReward + M * kRow;             // i32 [T]      in / out (in place)
  static constexpr int oLive = oTep + kRowI;                  // i32 [M-1][T] in   (bodies 1..)
  static constexpr int oWinner = oLive + kLiveRows * kRowI;   // i32 [T]      in (latch) / out
  static constexpr int oDone = oWinner + kRowI;               // u8 [T]       in   (latched done)
  static constexpr int oDoneOut = oDone + kTile;              // u8 [T]       out
  static constexpr int oTangent = oDoneOut + kTile;           // u8 [T]       out
  static constexpr int kStageBytes = oTangent + kTile;        // multiple of 128
  static constexpr int kMaxRows = M * 4 * 2 + M * 2 * 2 + M + 8;
  static constexpr int kBarOffset = kTmaStages * kStageBytes;
  static constexpr int kTableOffset = kBarOffset + 64;
  static constexpr int kSmemBytes = kTableOffset + 2 * kMaxRows * 16;
};

struct TmaRow {               // one row segment per tile: global address of tile 0, smem offset, bytes (stride = bytes)
  unsigned long long gbase;
  uint32_t smem_off, bytes;
};

template <typename R, int M, bool GENERIC>
#ifdef CAV_TMA_MAXNREG
__global__ void __launch_bounds__(kTmaThreads) __maxnreg__(CAV_TMA_MAXNREG) step_tma_kernel(
#else
__global__ void __launch_bounds__(kTmaThreads, CAV_MIN_BLOCKS_TMA) step_tma_kernel(
#endif
    const __grid_constant__ DevScenario<R> sc,
                                                                              const __grid_constant__ EnvBuffers<R> buf,
                                                                              const __grid_constant__ StepIO<R> io,
                                                                              int64_t t_global, int64_t n_tiles) {
  using L = TmaLayout<R, M>;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);   // [stages] tile landed (producer -> consumers)
  uint64_t* done = full + kTmaStages;                                    // [stages] tile computed (consumers -> producer)
  TmaRow* in_rows = reinterpret_cast<TmaRow*>(smem + L::kTableOffset);
  TmaRow* out_rows = in_rows + L::kMaxRows;
  __shared__ int n_in_s, n_out_s, in_bytes_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n = buf.n, lo = buf.lo;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], kConsumerWarps); }
    mbar_fence_init();
    // row tables (a few dozen entries, once per CTA)
    int ni = 0, no = 0;
    auto in = [&](const void* base, int64_t row, int elem, int off) {
      in_rows[ni++] = {(unsigned long long)base + (unsigned long long)((row * n + lo) * elem), (uint32_t)off, (uint32_t)(kTile * elem)};
    };
    auto out = [&](void* base, int64_t row, int elem, int off) {
      out_rows[no++] = {(unsigned long long)base + (unsigned long long)((row * n + lo) * elem), (uint32_t)off, (uint32_t)(kTile * elem)};
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
    n_in_s = ni; n_out_s = no;
    uint32_t total = 0;
    for (int r = 0; r < ni; ++r) total += in_rows[r].bytes;
    in_bytes_s = (int)total;
  }
  __syncthreads();
  const int n_in = n_in_s, n_out = n_out_s;
  const uint32_t in_bytes = (uint32_t)in_bytes_s;

  // whole-warp call (warp 0): arm the stage's barrier with the tile's byte count, then one bulk copy per row
  auto issue_loads = [&](int s, int64_t tile) {
    unsigned char* st = smem + s * L::kStageBytes;
    if (lane == 0) mbar_expect_tx(&full[s], in_bytes);
    __syncwarp();
    for (int r = lane; r < n_in; r += 32) {
      const TmaRow row = in_rows[r];
      bulk_load(st + row.smem_off, reinterpret_cast<const void*>(row.gbase + (unsigned long long)tile * row.bytes), row.bytes, &full[s]);
    }
  };

  const int64_t first = blockIdx.x, stride = gridDim.x;

  if (warp == kConsumerWarps) {
    // ================= producer warp: HBM -> shared (bulk loads), shared -> HBM (bulk stores); no arithmetic
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
      mbar_wait(&done[s], parity);   // every consumer warp has written its results for this tile
      for (int r = lane; r < n_out; r += 32) {
        const TmaRow row = out_rows[r];
        bulk_store(reinterpret_cast<void*>(row.gbase + (unsigned long long)tile * row.bytes), st + row.smem_off, row.bytes);
      }
      bulk_commit();
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

    // ---- this thread's env: shared memory -> registers
    const int64_t e = lo + tile * kTile + tid;
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
      for (int c = 0; c < 4; ++c) env.s[b][c] = sS[(b * 4 + c) * kTile + tid];
      env.cs[b][0] = sC[(b * 2 + 0) * kTile + tid];
      env.cs[b][1] = sC[(b * 2 + 1) * kTile + tid];
      ext[b][0] = sA[(b * 2 + 0) * kTile + tid];
      ext[b][1] = sA[(b * 2 + 1) * kTile + tid];
      env.held[b][0] = R(0); env.held[b][1] = R(0);
      if (b > 0) env.live[b] = sL[(b - 1) * kTile + tid];
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
          for (int c = 0; c < 4; ++c) sS[(b * 4 + c) * kTile + tid] = now.s[b][c];
      };
      transition<R, M, false, GENERIC>(sc, buf, e, t_global, env, ext, res, moved);
      if (res.invalid) buf.err[e] = 1;
      if (res.tangent) count_tangent(buf.stats);
      if (env.done) {
        score_episode<R, M>(buf, env);
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
    for (int b = 0; b < M; ++b) sRw[b * kTile + tid] = res.reward[b];
    reinterpret_cast<int32_t*>(st + L::oWinner)[tid] = res.winner;
    st[L::oDoneOut + tid] = res.terminate ? 1 : 0;
    st[L::oTangent + tid] = res.tangent ? 1 : 0;

    // ---- hand the tile to the producer: writes visible to the bulk-copy engine, one arrival per warp
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(&done[s]);
  }
}

// Host side: can [lo, hi) of this engine be stepped by the TMA kernel?  Returns the number of whole tiles (0 = no).
template <typename R>
inline int64_t tma_tiles(const EnvBuffers<R>& buf, const StepIO<R>& io) {
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!io.actions || buf.n % 4 != 0 || buf.lo % kTile != 0) return 0;
  if (!aligned(io.actions) || !aligned(io.state_out) || !aligned(io.reward_out) || !aligned(io.done_out) || !aligned(io.winner_out) ||
      !aligned(io.tangent_out))
    return 0;
  return (buf.hi - buf.lo) / kTile;
}

template <typename R, int M>
bool launch_step_tma(const DevScenario<R>& sc, const EnvBuffers<R>& buf, const StepIO<R>& io, int64_t t_global, cudaStream_t stream,
                     int64_t* envs_done) {
  using L = TmaLayout<R, M>;
  *envs_done = 0;
  const int64_t tiles = tma_tiles(buf, io);
  if (tiles == 0) return true;
  static int resident[2] = {0, 0};  // [generic]: CTAs per SM, queried once per instantiation
  static int sms = 0;
  auto kernel = sc.homogeneous ? step_tma_kernel<R, M, false> : step_tma_kernel<R, M, true>;
  int& res = resident[sc.homogeneous ? 0 : 1];
  if (res == 0) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kSmemBytes) != cudaSuccess) return false;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, kernel, kTmaThreads, L::kSmemBytes) != cudaSuccess || res < 1) { res = 0; return false; }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int64_t grid = tiles < (int64_t)sms * res ? tiles : (int64_t)sms * res;
  kernel<<<(unsigned)grid, kTmaThreads, L::kSmemBytes, stream>>>(sc, buf, io, t_global, tiles);
  *envs_done = tiles * kTile;
  return true;
}

template <typename R, int M>
constexpr SmallLaunchers<R> make_launchers() {
  return {&launch_step<R, M>, &launch_replay<R, M>, &launch_rollout<R, M>, &launch_reset<R, M>, &launch_step_tma<R, M>};
}

}  // namespace cav
