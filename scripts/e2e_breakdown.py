"""Wall time per cavgym_step_host call (pinned buffers, zero copy) at the C2 shape; run under
`ncu --metrics gpu__time_duration.sum -k regex:step_tma` to see how much of it is the kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from cavgym_b200 import BatchedCAVEnv
n, m, steps = 65536, 2, int(sys.argv[1]) if len(sys.argv) > 1 else 200
device = torch.device("cuda", 0)
init, actions = bench.make_trace(torch, device, n, steps + 3, "float64", 0)
env = BatchedCAVEnv(None, None, None, num_envs=n, dtype="float64", compiled=bench.scenario("external"), device=device)
h_actions = torch.empty((steps + 3, m, 2, n), dtype=env.dtype).pin_memory(); h_actions.copy_(actions)
h_state = torch.empty((m, 4, n), dtype=env.dtype).pin_memory(); h_reward = torch.empty((m, n), dtype=env.dtype).pin_memory()
h_done = torch.empty(n, dtype=torch.uint8).pin_memory(); h_winner = torch.empty(n, dtype=torch.int32).pin_memory(); h_tangent = torch.empty(n, dtype=torch.uint8).pin_memory()
env.reset(init_state=init)
for t in range(3):
    env.step_host(h_actions[t], h_state, h_reward, h_done, h_winner, h_tangent)
t0 = time.perf_counter()
for t in range(steps):
    env.step_host(h_actions[3 + t], h_state, h_reward, h_done, h_winner, h_tangent)
dt = time.perf_counter() - t0
print(f"step_host: {dt / steps * 1e6:.1f} us per call, {n * steps / dt / 1e6:.1f} M env-steps/s")
# only the outputs a trainer needs every step (reward, done): how much of the time is the state read-back?
t0 = time.perf_counter()
for t in range(steps):
    env.step_host(h_actions[3 + t], None, h_reward, h_done, None, None)
dt = time.perf_counter() - t0
print(f"step_host without state/winner/tangent outputs: {dt / steps * 1e6:.1f} us per call")
