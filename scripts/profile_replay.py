"""cavgym_replay (50 fused steps, trajectories recorded) and cavgym_step timings for one stock scenario on replayed actions:
    python scripts/profile_replay.py [--scenario bus-stop] [--pedestrians K] [--envs 65536]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import compile_from_meta, load_golden
from cavgym_b200 import BatchedCAVEnv
GOLDEN = {"pedestrians": "pedestrians_rc_seed0", "crossroads": "crossroads_random_all_seed6", "bus-stop": "busstop_random_all_seed8",
          "pelican-crossing": "pelican_random_all_seed10"}
ap = argparse.ArgumentParser()
ap.add_argument("--scenario", default="bus-stop"); ap.add_argument("--envs", type=int, default=65536); ap.add_argument("--pedestrians", type=int, default=0)
ap.add_argument("--dtype", default="float64"); ap.add_argument("--steps", type=int, default=50)
args = ap.parse_args()
meta, _ = load_golden(GOLDEN[args.scenario])
meta["config"]["tester_config"]["epsilon"] = 0.01
if args.pedestrians:
    meta["config"]["scenario_config"]["num_pedestrians"] = args.pedestrians
dev = torch.device("cuda", 0)
gen = BatchedCAVEnv(None, None, None, num_envs=args.envs, dtype=args.dtype, compiled=compile_from_meta(meta, mode="device"), device=dev, seed=0)
gen.set_action_logging(True); gen.reset(); gen.rollout(100, auto_reset=True)
init = gen.state.clone()
actions = torch.empty((args.steps, gen.num_bodies, 2, args.envs), dtype=gen.dtype, device=dev)
for t in range(args.steps):
    gen.step(None); actions[t].copy_(gen.actions_taken)
gen.close()
env = BatchedCAVEnv(None, None, None, num_envs=args.envs, dtype=args.dtype, compiled=compile_from_meta(meta), device=dev)
def timed(fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
def replay():
    env.reset(init_state=init); env.replay(actions)
def reset_only():
    env.reset(init_state=init)
ms_replay = timed(replay, 5) - timed(reset_only, 5)
env.reset(init_state=init)
ms_step = timed(lambda: env.step(actions[0]), 20)
env.set_step_path(use_tma=False)
ms_plain = timed(lambda: env.step(actions[0]), 20)
ms_replay_plain = timed(replay, 5) - timed(reset_only, 5)
print(f"plain step {ms_plain:.4f} ms, plain replay {ms_replay_plain:.3f} ms;", end=" ")
print(f"{args.scenario} bodies {env.num_bodies} envs {args.envs}: replay {args.steps} steps {ms_replay:.3f} ms ({args.envs * args.steps / ms_replay / 1e6:.1f} G env-steps/s... x1e-3), step {ms_step:.4f} ms")
