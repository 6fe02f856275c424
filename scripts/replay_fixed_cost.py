"""GPU-side duration of one cavgym_replay launch as a function of the fused step count, with every launch enqueued
behind a gate kernel so that host launch latency cannot sit inside the event windows.

    python scripts/replay_fixed_cost.py [--envs 65536] [--start 5] [--reps 30]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from cavgym_b200 import BatchedCAVEnv  # noqa: E402
from cavgym_b200._native import check  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--envs", type=int, default=65536)
    parser.add_argument("--start", type=int, default=5, help="trace position the measured launch starts from")
    parser.add_argument("--reps", type=int, default=30)
    parser.add_argument("--dtype", default="float64")
    parser.add_argument("--steps", default="1,2,5,10,20,40,80,160")
    parser.add_argument("--no-tma", action="store_true")
    parser.add_argument("--no-record", action="store_true")
    parser.add_argument("--train", type=int, default=40, help="also time TRAIN launches of --train-steps steps issued back to back (one event pair)")
    parser.add_argument("--train-steps", type=int, default=20)
    parser.add_argument("--chain", type=int, default=0, help="also time CHAIN consecutive launches (no kernel in between) of the longest step count / CHAIN")
    args = parser.parse_args()
    dev = torch.device("cuda", 0)
    n, m = args.envs, 2
    step_list = [int(s) for s in args.steps.split(",")]
    longest = max(step_list)
    init, actions = bench.make_trace(torch, dev, n, args.start + longest, args.dtype, 0)
    env = BatchedCAVEnv(None, None, None, num_envs=n, dtype=args.dtype, compiled=bench.scenario("external"), device=dev)
    if args.no_tma:
        env.set_step_path(False)
    longest_slab = max(longest, args.start)
    slab = {"state": torch.empty((longest_slab, m, 4, n), dtype=env.dtype, device=dev),
            "reward": torch.empty((longest_slab, m, n), dtype=env.dtype, device=dev),
            "done": torch.empty((longest_slab, n), dtype=torch.uint8, device=dev),
            "winner": torch.empty((longest_slab, n), dtype=torch.int32, device=dev),
            "tangent": torch.empty((longest_slab, n), dtype=torch.uint8, device=dev)}
    lib, handle, stream = env._lib, env._handle, env._stream()

    def ptr(t):
        return None if (t is None or args.no_record) else C.c_void_p(t.data_ptr())

    def replay(first, count):
        check(lib.cavgym_replay(handle, count, C.c_void_p(actions[first].data_ptr()), ptr(slab["state"]), ptr(slab["reward"]),
                                ptr(slab["done"]), ptr(slab["winner"]), ptr(slab["tangent"]), stream))

    # state at trace position `start`, to restart every repetition from the same place
    env.reset(init_state=init)
    if args.start:
        replay(0, args.start)
    at_start = env.state.clone()
    torch.cuda.synchronize()
    print(f"envs {n} dtype {args.dtype} tma {not args.no_tma} record {not args.no_record} start {args.start}")
    for steps in step_list:
        for gated in (True, False):
            events = []
            for _ in range(3):
                env.reset(init_state=at_start)
                replay(args.start, steps)
            torch.cuda.synchronize()
            if gated:
                torch.cuda._sleep(int(0.03 * 1.9e9))   # 30 ms: the host enqueues everything below meanwhile
            for _ in range(args.reps):
                env.reset(init_state=at_start)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                replay(args.start, steps)
                b.record()
                events.append((a, b))
            torch.cuda.synchronize()
            us = sorted(a.elapsed_time(b) * 1e3 for a, b in events)
            tag = "queued " if gated else "starved"
            print(f"steps {steps:4d} {tag}: min {us[0]:8.1f}  med {us[len(us) // 2]:8.1f}  max {us[-1]:8.1f} us   "
                  f"med/step {us[len(us) // 2] / steps:6.2f} us")
    if args.train:
        steps, k = args.train_steps, args.train
        totals = []
        for rep in range(8):
            env.reset(init_state=at_start)
            torch.cuda.synchronize()
            torch.cuda._sleep(int(0.01 * 1.9e9))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for j in range(k):
                replay(args.start + (j * steps) % max(1, longest - steps + 1), steps)
            b.record()
            torch.cuda.synchronize()
            totals.append(a.elapsed_time(b) * 1e3 / k)
        totals.sort()
        print(f"train of {k} x {steps} steps back to back: min {totals[0]:8.1f}  med {totals[len(totals) // 2]:8.1f} us per launch   "
              f"{totals[len(totals) // 2] / steps:6.2f} us per step")
    if args.chain:
        k = args.chain
        steps = longest // k
        per = [[] for _ in range(k)]
        for rep in range(args.reps):
            env.reset(init_state=at_start)
            if rep == 0:
                torch.cuda.synchronize()
                torch.cuda._sleep(int(0.03 * 1.9e9))
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
            evs[0].record()
            for j in range(k):
                replay(args.start + j * steps, steps)
                evs[j + 1].record()
            per[0].append(evs)
        torch.cuda.synchronize()
        for j in range(k):
            us = sorted(e[j].elapsed_time(e[j + 1]) * 1e3 for e in per[0])
            print(f"chain of {k} x {steps} steps, launch {j}: min {us[0]:8.1f} med {us[len(us) // 2]:8.1f} max {us[-1]:8.1f} us")
    env.close()


if __name__ == "__main__":
    main()
