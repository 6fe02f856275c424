"""Environment sharding across the GPUs of one box and the single end-of-run reduction.

Environments are independent (reference library/environment.py:119-223 has no cross-env term; the reference itself
parallelises by process, experiments.py:121-122), so rank r of R owns the contiguous global env range
[r * n, (r + 1) * n) (weak scaling, n envs per GPU), the Philox streams are keyed by the GLOBAL env id
(cavgym_set_shard), and there is no collective in the step.  Only the twelve episode counters (reporting.py:227-269) are
summed at the end with ONE all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import os

from . import _abi

STAT_KEYS = _abi.STAT_NAMES


def rank_world():
    """RANK / WORLD_SIZE / LOCAL_RANK as torchrun sets them (1 process per GPU)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def shard_offset(rank, envs_per_rank):
    """Global id of env 0 of this rank."""
    if rank < 0 or envs_per_rank <= 0:
        raise ValueError("rank must be >= 0 and envs_per_rank positive")
    return rank * envs_per_rank


def split_envs(total_envs, world):
    """Strong-scaling split of a fixed batch: (offset, count) per rank, contiguous, sizes differing by at most one."""
    if world <= 0 or total_envs < world:
        raise ValueError("need at least one env per rank")
    base, extra = divmod(total_envs, world)
    out, offset = [], 0
    for r in range(world):
        count = base + (1 if r < extra else 0)
        out.append((offset, count))
        offset += count
    return out


def reduce_stats(stats, device=None, group=None):
    """Sum the episode counters of every rank: one all-reduce of twelve int64 words.  `stats` is the dict returned by
    BatchedCAVEnv.stats(); returns the same dict summed over ranks (unchanged without a process group)."""
    import torch
    import torch.distributed as dist
    vec = torch.tensor([int(stats[k]) for k in STAT_KEYS], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(STAT_KEYS, [int(v) for v in vec.tolist()]))


def reduce_timing(elapsed_ms, units, device=None, group=None):
    """Whole-job throughput inputs: MAX of the per-rank device time, SUM of the units processed."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(t.item()), float(u.item())
