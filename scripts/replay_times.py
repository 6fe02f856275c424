import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import bench
from cavgym_b200 import BatchedCAVEnv
dev = torch.device("cuda", 0)
n = 65536
init, actions = bench.make_trace(torch, dev, n, 60, "float64", 0, advance=200)
env = BatchedCAVEnv(None, None, None, num_envs=n, dtype="float64", compiled=bench.scenario("external"), device=dev)
dbg = torch.zeros(2 * 3 * n, dtype=torch.float64, device=dev)   # shape the override API expects [M,3,N]
env.set_uniform_override(dbg.view(2, 3, n))
for rep in range(3):
    env.reset(init_state=init)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env.replay(actions[:50]); b.record(); torch.cuda.synchronize()
    print("launch ms", a.elapsed_time(b))
d = dbg.cpu().numpy()
se = d[:2 * 293].reshape(293, 2)
t0 = se[:, 0].min()
print("CTA start (us) min/med/max", (se[:, 0] - t0).min() / 1e3, np.median(se[:, 0] - t0) / 1e3, (se[:, 0] - t0).max() / 1e3)
print("CTA end   (us) min/med/max", (se[:, 1] - t0).min() / 1e3, np.median(se[:, 1] - t0) / 1e3, (se[:, 1] - t0).max() / 1e3)
dur = (se[:, 1] - se[:, 0]) / 1e3
print("CTA duration us: min %.1f p10 %.1f med %.1f p90 %.1f max %.1f" % (dur.min(), np.percentile(dur, 10), np.median(dur), np.percentile(dur, 90), dur.max()))
steps = d[4096:4096 + 4 * 8 * 60].reshape(4, 8, 60)
for blk in range(2):
    for w in range(7):
        print("blk", blk, "warp", w, "step start (us):", np.round(steps[blk, w, :50:7] / 1e3, 1))

print("path counters (cumulative at block 0 exit of last launch): sat_quad, share_general, sincos_wide, wrap_slow, steer_libm, steer_general, turn, kerb, ego_near, finish_near")
print(d[2048:2058])
