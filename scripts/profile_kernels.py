"""Small driver for ncu: a few launches of the per-step kernel (large batch) and of the fused replay kernel.
    python scripts/profile_kernels.py [--envs N] [--dtype float64] [--mode step|replay|rollout|all]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bench  # noqa: E402
from cavgym_b200 import BatchedCAVEnv  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--envs", type=int, default=1 << 20)
    parser.add_argument("--dtype", default="float64")
    parser.add_argument("--mode", default="all")
    parser.add_argument("--launches", type=int, default=4)
    parser.add_argument("--replay-steps", type=int, default=0, help="steps fused per replay launch (default: bench.CHUNK)")
    parser.add_argument("--advance", type=int, default=0, help="steps to advance (fused replay) before the measured step launches")
    args = parser.parse_args()
    device = torch.device("cuda", 0)
    if args.mode in ("step", "all"):
        init, actions = bench.make_trace(torch, device, args.envs, 8, args.dtype, 0, advance=args.advance)
        env = BatchedCAVEnv(None, None, None, num_envs=args.envs, dtype=args.dtype, compiled=bench.scenario("external"), device=device)
        env.reset(init_state=init)
        times = []
        for t in range(args.launches + 2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); env.step(actions[t % 8]); b.record()
            times.append((a, b))
        torch.cuda.synchronize()
        print("step_kernel ms:", [round(a.elapsed_time(b), 4) for a, b in times])
        env.close()
    if args.mode in ("replay", "all"):
        n = 65536
        steps = args.replay_steps or bench.CHUNK
        init, actions = bench.make_trace(torch, device, n, steps, args.dtype, 0)
        env = BatchedCAVEnv(None, None, None, num_envs=n, dtype=args.dtype, compiled=bench.scenario("external"), device=device)
        env.reset(init_state=init)
        times = []
        for t in range(args.launches):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); env.replay(actions[:steps]); b.record()
            times.append((a, b))
        torch.cuda.synchronize()
        print(f"replay_kernel ({steps} steps) ms:", [round(a.elapsed_time(b), 4) for a, b in times])
        env.close()
    if args.mode in ("rollout", "all"):
        env = BatchedCAVEnv(None, None, None, num_envs=args.envs, dtype=args.dtype, compiled=bench.scenario("device"), device=device)
        env.reset()
        times = []
        for t in range(args.launches):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); env.rollout(50); b.record()
            times.append((a, b))
        torch.cuda.synchronize()
        print("rollout_kernel (50 steps) ms:", [round(a.elapsed_time(b), 4) for a, b in times])
        print(env.stats())
        env.close()


if __name__ == "__main__":
    main()
