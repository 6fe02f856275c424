"""Does launch time depend on how long the GPU has been busy?  Replays 50-step chunks back to back and prints the
per-launch time at several points, with the SM clock NVML reports."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, pynvml
import bench
from cavgym_b200 import BatchedCAVEnv
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda", 0)
n = 65536
init, actions = bench.make_trace(torch, dev, n, 100, "float64", 0)
env = BatchedCAVEnv(None, None, None, num_envs=n, dtype="float64", compiled=bench.scenario("external"), device=dev)
slab = None
for rounds in (5, 20, 100, 400, 400):
    env.reset(init_state=init)
    evs = []
    for i in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if i % 2 == 0: env.reset(init_state=init)
        a.record(); env.replay(actions[(i % 2) * 50:(i % 2) * 50 + 50]); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    print(f"rounds {rounds:4d}: first {ms[0]:.3f} median {sorted(ms)[len(ms)//2]:.3f} last {ms[-1]:.3f} min {min(ms):.3f} ms; sm clock {pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)} MHz")
