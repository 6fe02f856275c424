"""Wall time per cavgym_step_host call at the C2 shape (65,536 envs x 2 bodies, fp64, pinned buffers) for the settings of
the host path: resident CTAs per SM of the zero-copy launch, the plain (non-TMA) kernel on mapped memory, staged copies.

    python scripts/e2e_breakdown.py [steps]
"""
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
joint = [h_actions[t] for t in range(steps + 3)]


def run(label, outputs=True):
    env.reset(init_state=init)
    torch.cuda.synchronize()
    for t in range(3):
        env.step_host(joint[t], h_state, h_reward, h_done, h_winner, h_tangent)
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        for t in range(steps):
            if outputs:
                env.step_host(joint[3 + t], h_state, h_reward, h_done, h_winner, h_tangent)
            else:
                env.step_host(joint[3 + t], None, h_reward, h_done, None, None)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(f"{label:58s} {best / steps * 1e6:7.1f} us per call   {n * steps / best / 1e6:7.1f} M env-steps/s")


for ctas in (1, 2, 3):
    env.set_host_path(1 + ctas)
    run(f"zero copy, TMA-staged kernel, {ctas} CTA(s) per SM")
env.set_host_path(1)
run("default path, reward + done only", outputs=False)
from cavgym_b200.engine import HostBuffer
wc = HostBuffer((steps + 3, m, 2, n), "float64", write_combined=True)
wc.array[...] = h_actions.numpy()
pinned_joint = joint
joint = [wc.array[t] for t in range(steps + 3)]
env.set_host_path(1)
run("zero copy, actions in WRITE-COMBINED pinned memory")
joint = pinned_joint
env.set_step_path(False)
run("zero copy, plain kernel (LDG / STG on mapped memory)")
env.set_step_path(True)
env.set_host_path(0)
run("staged: chunked cudaMemcpyAsync + kernel")
