"""cavgym_rollout of a stock heterogeneous scenario on the team-of-warps kernel (kernels_team.cuh) and on the thread-per-env kernel:
    python scripts/ab_team.py [--envs 1048576] [--chunk 100] [--launches 5] [--scenarios bus-stop,crossroads,pelican-crossing]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import compile_from_meta, load_golden
from cavgym_b200 import BatchedCAVEnv
GOLDEN = {"crossroads": "crossroads_random_all_seed6", "bus-stop": "busstop_random_all_seed8", "pelican-crossing": "pelican_random_all_seed10"}
ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1048576); ap.add_argument("--chunk", type=int, default=100)
ap.add_argument("--launches", type=int, default=5); ap.add_argument("--dtype", default="float64")
ap.add_argument("--scenarios", default="bus-stop,crossroads,pelican-crossing"); ap.add_argument("--paths", default="0,1")
args = ap.parse_args()
for name in args.scenarios.split(","):
    for team in [int(p) for p in args.paths.split(",")]:
        meta, _ = load_golden(GOLDEN[name])
        meta["config"]["tester_config"]["epsilon"] = 0.01
        env = BatchedCAVEnv(None, None, None, num_envs=args.envs, dtype=args.dtype, compiled=compile_from_meta(meta, mode="device"), device="cuda:0", seed=0)
        env.set_rollout_path(bool(team))
        env.reset()
        times = []
        for _ in range(args.launches):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); env.rollout(args.chunk, auto_reset=True); b.record(); times.append((a, b))
        torch.cuda.synchronize()
        ms = [round(a.elapsed_time(b), 3) for a, b in times]
        st = env.stats()
        print(name, "team" if team else "thread-per-env", "ms per launch:", ms, "G env-steps/s (all launches): %.3f" % (args.envs * args.chunk * len(ms) / sum(ms) / 1e6),
              "episodes", st["episodes"], "tangent", st["tangent"], flush=True)
        env.close()
