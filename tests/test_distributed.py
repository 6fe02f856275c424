"""Multi-GPU host logic on CPU: two gloo ranks, each owning a contiguous shard of the global env range.

What must hold: Philox is keyed by the GLOBAL env id, so (a) the union of the two shards' results equals one
process running the whole batch, state for state, and (b) the single end-of-run all-reduce of the ten episode counters
gives the totals of the whole batch.  The engine cannot run without a GPU, so the CPU oracle — which restates the same
keying (oracle/cavgym_oracle.c draw_block) — stands in for it on each rank; the GPU-side twin of (a) is
tests/test_gpu_replay.py::test_sharded_engines_match_one_engine.
"""
import os
import socket

import numpy as np
import pytest

from helpers import compile_from_meta, load_golden

WORLD, PER_RANK, STEPS = 2, 48, 400


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from cavgym_b200 import sharding
    from oracle.oracle import Oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.rank_world() == (rank, world, rank)
        meta, _ = load_golden("pedestrians_rc_eps05_seed1")
        sim = Oracle(compile_from_meta(meta, mode="device"), PER_RANK, seed=5)
        sim.set_shard(sharding.shard_offset(rank, PER_RANK))
        sim.reset()
        sim.rollout(STEPS, auto_reset=True)
        local = sim.stats()
        total = sharding.reduce_stats(local)
        elapsed, units = sharding.reduce_timing(10.0 + rank, local["env_steps"])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), state=sim.state.copy(), local=np.array([local[k] for k in sharding.STAT_KEYS]),
                 total=np.array([total[k] for k in sharding.STAT_KEYS]), timing=np.array([elapsed, units]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_equal_one_process(tmp_path):
    import torch.multiprocessing as mp
    from cavgym_b200 import sharding
    from oracle.oracle import Oracle
    port = _free_port()
    mp.spawn(_worker, args=(WORLD, port, str(tmp_path)), nprocs=WORLD, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(WORLD)]
    meta, _ = load_golden("pedestrians_rc_eps05_seed1")
    whole = Oracle(compile_from_meta(meta, mode="device"), WORLD * PER_RANK, seed=5, threads=4)
    whole.reset()
    whole.rollout(STEPS, auto_reset=True)
    want = whole.stats()
    # (a) shard r holds exactly the envs [r * n, (r + 1) * n) of the single-process batch
    for r in range(WORLD):
        assert np.array_equal(ranks[r]["state"], whole.state[:, :, r * PER_RANK:(r + 1) * PER_RANK])
    # (b) the all-reduce gives the whole batch's counters, on every rank
    for r in range(WORLD):
        assert ranks[r]["total"].tolist() == [want[k] for k in sharding.STAT_KEYS]
    assert (ranks[0]["local"] + ranks[1]["local"]).tolist() == ranks[0]["total"].tolist()
    assert want["episodes"] >= 8 and want["env_steps"] == WORLD * PER_RANK * STEPS
    # timing reduction: MAX of the times, SUM of the units
    assert ranks[0]["timing"].tolist() == [11.0, float(want["env_steps"])] == ranks[1]["timing"].tolist()


def test_split_envs_covers_the_range_once():
    from cavgym_b200 import sharding
    for total, world in ((1048576, 8), (100000, 8), (10, 3), (7, 7)):
        parts = sharding.split_envs(total, world)
        assert parts[0][0] == 0 and sum(c for _, c in parts) == total
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        sharding.split_envs(3, 4)
    assert sharding.shard_offset(3, 65536) == 196608
