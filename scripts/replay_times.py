"""Per-CTA timeline of one cavgym_replay launch (needs a library built with -DCAV_DEBUG_TIMES, see CAVGYM_LIB):
CTA entry, end of load_env, end of every warp's first and last step, CTA exit, on the global timer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import bench
from cavgym_b200 import BatchedCAVEnv
dev = torch.device("cuda", 0)
n = 65536
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
start = int(sys.argv[2]) if len(sys.argv) > 2 else 5
chain = int(sys.argv[3]) if len(sys.argv) > 3 else 1   # consecutive launches; the timeline is that of the LAST one
init, actions = bench.make_trace(torch, dev, n, start + steps * chain, "float64", 0)
env = BatchedCAVEnv(None, None, None, num_envs=n, dtype="float64", compiled=bench.scenario("external"), device=dev)
dbg = torch.zeros(2 * 3 * n, dtype=torch.float64, device=dev)   # shape the override API expects [M,3,N]
env.reset(init_state=init)
if start:
    env.replay(actions[:start])
at_start = env.state.clone()
env.set_uniform_override(dbg.view(2, 3, n))
for rep in range(4):
    env.reset(init_state=at_start)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for j in range(chain):
        if j == chain - 1:
            a.record()
        env.replay(actions[start + j * steps:start + (j + 1) * steps])
    b.record(); torch.cuda.synchronize()
    raw_probe = dbg.cpu().numpy()[:293 * 20].reshape(293, 20)
    sm = raw_probe[:, 19].copy()
    d = raw_probe / 1e3
    t0 = d[:, 0].min()
    d = d - t0
    q = lambda v: "min %.1f p10 %.1f med %.1f p90 %.1f max %.1f" % (v.min(), np.percentile(v, 10), np.median(v), np.percentile(v, 90), v.max())
    print(f"--- rep {rep}: launch {a.elapsed_time(b) * 1e3:.1f} us (event to event)")
    print("CTA entry            ", q(d[:, 0]))
    print("load_env done - entry", q(d[:, 1] - d[:, 0]))
    first = d[:, 2:9]
    last = d[:, 10:17]
    print("first step done - load_env (slowest warp)", q(first.max(1) - d[:, 1]))
    print("first step done - load_env (fastest warp)", q(first.min(1) - d[:, 1]))
    print("per-step after the first, slowest warp   ", q((last.max(1) - first.max(1)) / max(1, steps - 1)))
    print("last step done, slowest warp (abs)       ", q(last.max(1)))
    print("last step done, fastest warp (abs)       ", q(last.min(1)))
    print("CTA exit (abs)                           ", q(d[:, 18]))
    print("exit - last step done                    ", q(d[:, 18] - last[:, 0]))
    dur = d[:, 18] - d[:, 0]
    print("CTA duration                             ", q(dur))
    order = np.argsort(-d[:, 18])[:8]
    print("slowest CTAs:", [(int(i), round(float(d[i, 18]), 1)) for i in order])
    if os.environ.get("CAV_COUNTERS"):
        import ctypes as C
        raw = C.CDLL(os.environ["CAVGYM_LIB"])
        buf = (C.c_ulonglong * (512 * 32))()
        raw.cavgym_debug_counters(buf)
        cnt = np.frombuffer(buf, dtype=np.uint64).reshape(512, 32)[:293].astype(np.float64)
        names = ["sat_quad", "share_general", "sincos_wide", "wrap_slow", "steer_libm", "steer_general", "turn", "kerb", "ego_near",
                 "finish_near", "corner"]
        print("path                warp-execs/CTA: mean    max | corr with CTA duration | slowest 4 CTAs")
        for i, nm in enumerate(names):
            c = cnt[:, i]
            if c.sum() == 0:
                continue
            corr = np.corrcoef(c, dur)[0, 1] if c.std() > 0 else 0.0
            print(f"{nm:18s} {c.mean():10.1f} {c.max():8.0f}   lanes/exec {cnt[:, 16 + i].sum() / c.sum():5.2f}   corr {corr:5.2f}   {[int(c[j]) for j in order[:4]]}")
        by_sm = {}
        for j in range(293):
            by_sm.setdefault(int(sm[j]), []).append(j)
        print("slowest CTAs' SMs:", [(int(j), int(sm[j]), [int(x) for x in by_sm[int(sm[j])]]) for j in order[:6]])
        print("CTAs alone on their SM:", [(v[0], round(float(dur[v[0]]), 1)) for v in by_sm.values() if len(v) == 1])
