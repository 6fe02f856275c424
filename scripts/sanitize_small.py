"""Small shapes of every hot kernel for compute-sanitizer (memcheck / racecheck / synccheck): the mbarrier pipelines of
step_tma_kernel and replay_tma_kernel (ragged last tile included), the plain step / replay / rollout kernels of a five-body
scenario, and the warp-per-env dense kernel with its warp-synchronous shared-memory sweeps.

    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import compile_from_meta, load_golden, soa  # noqa: E402

from cavgym_b200 import BatchedCAVEnv  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    # ---- two bodies, replayed actions: step_tma_kernel + replay_tma_kernel, 400 envs (3 tiles of 128 + a ragged one of 16)
    meta, episodes = load_golden("pedestrians_rc_eps05_seed1")
    n, steps = 400, 12
    init = soa(np.stack([episodes[e % len(episodes)]["init_state"] for e in range(n)]))
    actions = np.stack([np.stack([episodes[e % len(episodes)]["actions"][t] for e in range(n)], axis=-1) for t in range(steps)])
    for dtype in ("float64", "float32"):
        env = BatchedCAVEnv(None, None, None, num_envs=n, dtype=dtype, compiled=compile_from_meta(meta), device=dev)
        env.reset(init_state=init)
        for t in range(4):
            env.step(actions[t])
        env.replay(actions[4:])
        torch.cuda.synchronize()
        print(dtype, "two-body step/replay", env.stats()["env_steps"])
        env.close()
    # ---- five bodies (bus stop): plain step / replay kernels and the rollout kernel with on-device agents
    meta, episodes = load_golden("busstop_random_all_seed8")
    ep = episodes[0]
    env = BatchedCAVEnv(None, None, None, num_envs=96, dtype="float64", compiled=compile_from_meta(meta), device=dev)
    env.reset(init_state=soa(np.repeat(ep["init_state"][None], 96, axis=0)))
    env.step(np.repeat(ep["actions"][0][..., None], 96, axis=-1))
    env.replay(np.repeat(ep["actions"][1:9][..., None], 96, axis=-1))
    env.close()
    env = BatchedCAVEnv(None, None, None, num_envs=96, dtype="float64", compiled=compile_from_meta(meta, mode="device"), device=dev, seed=2)
    env.reset()
    env.rollout(10, auto_reset=True)
    torch.cuda.synchronize()
    print("five-body step/replay/rollout", env.stats()["env_steps"])
    env.close()
    # ---- 320 bodies: dense_kernel (rollout with on-device agents, auto-reset) and the dense step on replayed actions
    from cavgym_b200.examples.environments import dense_traffic
    from cavgym_b200.library.bodies import Pedestrian
    from cavgym_b200.scenario import AgentSpec, compile_scenario
    road_map, constants = dense_traffic.make_world()
    bodies = dense_traffic.make_bodies(np_random=np.random.RandomState(0), road_map=road_map)
    cfg = SimpleNamespace(terminate_collisions="all", terminate_ego_zones=True, terminate_ego_offroad=False, max_timesteps=1000,
                          reward_win=6000.0, reward_draw=2000.0, cost_step=4.0)
    specs = [AgentSpec("random-constrained", epsilon=0.05) if isinstance(b, Pedestrian) else AgentSpec("noop") for b in bodies]
    env = BatchedCAVEnv(None, None, None, num_envs=10, dtype="float64", compiled=compile_scenario(bodies, constants, cfg, specs), device=dev, seed=3)
    env.reset()
    env.rollout(6, auto_reset=True)
    torch.cuda.synchronize()
    print("dense rollout", env.stats()["env_steps"])
    env.close()
    env = BatchedCAVEnv(None, None, None, num_envs=10, dtype="float64",
                        compiled=compile_scenario(bodies, constants, cfg, [AgentSpec("external") for _ in bodies]), device=dev, seed=3)
    env.reset()
    env.step(torch.zeros((len(bodies), 2, 10), dtype=torch.float64, device=dev))
    torch.cuda.synchronize()
    print("dense step", env.stats()["env_steps"])
    env.close()


if __name__ == "__main__":
    main()
