import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from helpers import compile_from_meta, load_golden, soa
from cavgym_b200 import BatchedCAVEnv
name = sys.argv[1] if len(sys.argv) > 1 else "pedestrians_rc_seed0"
dtype = sys.argv[2] if len(sys.argv) > 2 else "float64"
meta, episodes = load_golden(name)
n, m = 5, meta["n_bodies"]
t_len = max(ep["actions"].shape[0] for ep in episodes)
init = soa(np.stack([episodes[e % len(episodes)]["init_state"] for e in range(n)]))
actions = np.zeros((t_len, m, 2, n))
for e in range(n):
    a = episodes[e % len(episodes)]["actions"][:t_len]
    actions[:a.shape[0], :, :, e] = a
outs = []
for dense in (False, True):
    env = BatchedCAVEnv(None, None, None, num_envs=n, dtype=dtype, compiled=compile_from_meta(meta))
    env.set_dense_path(dense)
    env.reset(init_state=init)
    out = {k: v.cpu().numpy() for k, v in env.replay(actions).items()}
    outs.append(out)
a, b = outs
for key in a:
    diff = np.argwhere(~np.isclose(a[key], b[key], rtol=0, atol=0, equal_nan=True))
    print(key, "first diff:", diff[:3].tolist())
    if len(diff):
        t = diff[0][0]
        print("  small", a[key][t].reshape(-1)[:16])
        print("  dense", b[key][t].reshape(-1)[:16])
        if t > 0:
            print("  prev small", a[key][t - 1].reshape(-1)[:16])
print("done small", a["done"].argmax(axis=0), "dense", b["done"].argmax(axis=0))
print("actions env1 body1 t0..3", actions[:4, 1, :, 1].tolist())
for t in range(3):
    print(t, "small", a["state"][t, 1, :, 1].tolist(), "dense", b["state"][t, 1, :, 1].tolist())
env = BatchedCAVEnv(None, None, None, num_envs=n, dtype=dtype, compiled=compile_from_meta(meta))
env.set_dense_path(True)
env.reset(init_state=init)
act = torch.tensor(actions, dtype=env.dtype, device=env.device)
for t in range(3):
    s = env.step(act[t])[0]
    print(t, "dense step()", s[1, :, 1].cpu().tolist())
