import sys, os, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "scripts"))
import bench_dense
from cavgym_b200 import BatchedCAVEnv
env = BatchedCAVEnv(None, None, None, num_envs=20000, dtype="float64", compiled=bench_dense.scenario(64, 256, 2e-4), device="cuda:0", seed=1)
env.reset(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): env.reset()
b.record(); torch.cuda.synchronize()
print("reset ms per call (20000 envs):", a.elapsed_time(b) / 5)
a.record()
for _ in range(5): env.rollout(1, auto_reset=True)
b.record(); torch.cuda.synchronize()
print("1-step rollout ms:", a.elapsed_time(b) / 5)
