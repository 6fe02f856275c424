"""Dense-traffic configuration (BASELINE config C4: 64 cars + 256 spawned pedestrians per env, all-pairs collisions) on one
GPU: env-steps/s, body-steps/s and pair-tests/s of the warp-per-env kernels, on-device agents with auto-reset.

    python scripts/bench_dense.py [--envs 100000] [--dtype float64] [--steps 20] [--chunk 5] [--replay]
"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cavgym_b200 import BatchedCAVEnv  # noqa: E402
from cavgym_b200.examples.environments import dense_traffic  # noqa: E402
from cavgym_b200.scenario import AgentSpec, compile_scenario  # noqa: E402


def scenario(cars, peds, epsilon, external=False, collisions="all", order="class"):
    road_map, constants = dense_traffic.make_world()
    bodies = dense_traffic.make_bodies(cars, peds, np_random=np.random.RandomState(0), road_map=road_map, order=order)
    cfg = SimpleNamespace(terminate_collisions=collisions, terminate_ego_zones=True, terminate_ego_offroad=False, max_timesteps=1000,
                          reward_win=6000.0, reward_draw=2000.0, cost_step=4.0)
    if external:
        specs = [AgentSpec("external") for _ in bodies]
    else:
        from cavgym_b200.library.bodies import Pedestrian
        specs = [AgentSpec("random-constrained", epsilon=epsilon) if isinstance(body, Pedestrian) else AgentSpec("noop") for body in bodies]
    return compile_scenario(bodies, constants, cfg, specs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=100000)
    ap.add_argument("--cars", type=int, default=64)
    ap.add_argument("--peds", type=int, default=256)
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--steps", type=int, default=20, help="timed launches")
    ap.add_argument("--chunk", type=int, default=5, help="env-steps per launch")
    ap.add_argument("--warm", type=int, default=60, help="untimed env-steps first (crossings under way)")
    ap.add_argument("--epsilon", type=float, default=2e-4)
    ap.add_argument("--collisions", default="all")
    ap.add_argument("--order", default="class")
    ap.add_argument("--replay", action="store_true", help="cavgym_step with a (noop) actions buffer instead of on-device agents")
    args = ap.parse_args()
    m = args.cars + args.peds
    env = BatchedCAVEnv(None, None, None, num_envs=args.envs, dtype=args.dtype, compiled=scenario(args.cars, args.peds, args.epsilon, args.replay, args.collisions, args.order),
                        device="cuda:0", seed=1)
    env.reset()
    if args.replay:
        actions = torch.zeros((m, 2, args.envs), dtype=env.dtype, device=env.device)
        run = lambda: [env.step(actions) for _ in range(args.chunk)]
    else:
        run = lambda: env.rollout(args.chunk, auto_reset=True)
        for _ in range(args.warm // args.chunk):
            run()
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    before = env.stats()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        run()
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop)
    after = env.stats()
    env_steps = after["env_steps"] - before["env_steps"]
    pairs = m * (m - 1) // 2
    real = 8 if args.dtype == "float64" else 4
    out = {"workload": f"dense traffic: {args.cars} cars + {args.peds} spawned pedestrians x {args.envs} envs, collisions={args.collisions}, "
                       + ("replayed noop actions, cavgym_step" if args.replay else f"on-device agents (eps={args.epsilon}), auto-reset, {args.chunk} steps/launch"),
           "dtype": args.dtype, "ms_per_env_step_batch": ms / (args.steps * args.chunk), "env_steps_per_sec": env_steps / ms * 1e3,
           "body_steps_per_sec": env_steps * m / ms * 1e3, "pair_tests_per_sec": env_steps * pairs / ms * 1e3,
           "live_env_steps": env_steps, "episodes": after["episodes"] - before["episodes"], "tangent": after["tangent"] - before["tangent"],
           "algorithmic_GBps": env_steps * m * 11 * real / ms * 1e3 / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
