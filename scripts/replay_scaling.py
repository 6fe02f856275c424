"""Time of one cavgym_replay launch as a function of the number of fused steps (fixed overhead vs per-step cost)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from cavgym_b200 import BatchedCAVEnv
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
init, actions = bench.make_trace(torch, dev, n, 200, "float64", 0, advance=200)
for use_tma in (True, False):
    env = BatchedCAVEnv(None, None, None, num_envs=n, dtype="float64", compiled=bench.scenario("external"), device=dev)
    env.set_step_path(use_tma)
    for record in (("state", "reward", "done", "winner", "tangent"), ()):
        for steps in (1, 2, 5, 10, 25, 50, 100, 200):
            ms = []
            for rep in range(5):
                env.reset(init_state=init)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); env.replay(actions[:steps], record=record); b.record()
                torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            print(f"tma={use_tma} record={len(record)} steps {steps:4d}: {min(ms)*1e3:9.1f} us  ({min(ms)*1e3/steps:7.2f} us/step)")
