"""A few cavgym_rollout launches of one stock scenario with on-device agents (for ncu):
    python scripts/profile_rollout.py [--scenario bus-stop] [--envs 131072] [--chunk 100] [--launches 4]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import compile_from_meta, load_golden
from cavgym_b200 import BatchedCAVEnv
GOLDEN = {"pedestrians": "pedestrians_rc_seed0", "crossroads": "crossroads_random_all_seed6", "bus-stop": "busstop_random_all_seed8",
          "pelican-crossing": "pelican_random_all_seed10"}
ap = argparse.ArgumentParser()
ap.add_argument("--scenario", default="bus-stop"); ap.add_argument("--envs", type=int, default=131072)
ap.add_argument("--chunk", type=int, default=100); ap.add_argument("--launches", type=int, default=4); ap.add_argument("--dtype", default="float64"); ap.add_argument("--pedestrians", type=int, default=0)
args = ap.parse_args()
meta, _ = load_golden(GOLDEN[args.scenario])
meta["config"]["tester_config"]["epsilon"] = 0.01
if args.pedestrians:
    meta["config"]["scenario_config"]["num_pedestrians"] = args.pedestrians
env = BatchedCAVEnv(None, None, None, num_envs=args.envs, dtype=args.dtype, compiled=compile_from_meta(meta, mode="device"), device="cuda:0", seed=0)
env.reset()
times = []
for _ in range(args.launches):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env.rollout(args.chunk, auto_reset=True); b.record(); times.append((a, b))
torch.cuda.synchronize()
ms = [round(a.elapsed_time(b), 3) for a, b in times]
print(args.scenario, "bodies", env.num_bodies, "rollout ms per launch:", ms, "env-steps/s:", args.envs * args.chunk / (ms[-1] * 1e-3), env.stats())
