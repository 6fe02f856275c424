"""cavgym_replay with PINNED HOST buffers: actions read from and trajectories written to host memory by the fused kernel
(per step 2.1 MB in, 5.6 MB out over PCIe), against cavgym_step_host called once per step.
    python scripts/replay_host_probe.py"""
import ctypes as C
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from cavgym_b200 import BatchedCAVEnv
from cavgym_b200._native import check

n, m, T = 65536, 2, 200
dev = torch.device("cuda", 0)
init, actions = bench.make_trace(torch, dev, n, T, "float64", 0)
env = BatchedCAVEnv(None, None, None, num_envs=n, dtype="float64", compiled=bench.scenario("external"), device=dev)
h_actions = torch.empty((T, m, 2, n), dtype=torch.float64).pin_memory(); h_actions.copy_(actions)
out = {"state": torch.empty((T, m, 4, n), dtype=torch.float64).pin_memory(), "reward": torch.empty((T, m, n), dtype=torch.float64).pin_memory(),
       "done": torch.empty((T, n), dtype=torch.uint8).pin_memory(), "winner": torch.empty((T, n), dtype=torch.int32).pin_memory(),
       "tangent": torch.empty((T, n), dtype=torch.uint8).pin_memory()}
p = lambda t: C.c_void_p(t.data_ptr())
lib, handle, stream = env._lib, env._handle, env._stream()
for chunk in (1, 5, 20, 50, 200):
    best = None
    for rep in range(3):
        env.reset(init_state=init)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for at in range(0, T, chunk):
            check(lib.cavgym_replay(handle, chunk, p(h_actions[at]), p(out["state"][at]), p(out["reward"][at]), p(out["done"][at]),
                                    p(out["winner"][at]), p(out["tangent"][at]), stream))
            torch.cuda.synchronize()      # the caller reads the chunk's results before it issues the next one
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(f"cavgym_replay on pinned host buffers, {chunk:3d} steps per call: {best / T * 1e6:7.1f} us per step  {n * T / best / 1e6:7.1f} M env-steps/s")
ref = BatchedCAVEnv(None, None, None, num_envs=n, dtype="float64", compiled=bench.scenario("external"), device=dev)
ref.reset(init_state=init)
got = ref.replay(actions)
torch.cuda.synchronize()
print("host-buffer trajectories equal the device-buffer ones:", torch.equal(got["state"].cpu(), out["state"]) and torch.equal(got["done"].cpu(), out["done"]))
