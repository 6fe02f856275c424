"""Fast GPU sanity check used while tuning kernels: M=2 golden cases through step and replay, fp64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import compile_from_meta, load_golden, soa, state_err, rel_err
from cavgym_b200 import BatchedCAVEnv

worst = 0.0
for name in ("pedestrians_rc_seed0", "pedestrians_rc_eps05_seed1", "pedestrians_proximity_seed3", "pedestrians_random_none_seed5"):
    meta, eps = load_golden(name)
    env = BatchedCAVEnv(None, None, None, num_envs=1, dtype="float64", compiled=compile_from_meta(meta))
    for ep in eps:
        env.reset(init_state=soa(ep["init_state"][None]))
        env.set_global_timestep(int(ep["t_global_start"]))
        out = env.replay(ep["actions"][..., None])
        st = out["state"].cpu().numpy()[..., 0]
        dn, wn, tg = (out[k].cpu().numpy()[:, 0] for k in ("done", "winner", "tangent"))
        mism = (dn != ep["done"]) | (wn != ep["winner"])
        assert not np.any(mism & ~tg.astype(bool)), (name, np.nonzero(mism)[0][:4])
        T = int(np.nonzero(mism)[0][0]) if mism.any() else len(dn)
        e1 = state_err(st[:T], ep["state"][:T]); e2 = rel_err(out["reward"].cpu().numpy()[:T, :, 0], ep["reward"][:T])
        worst = max(worst, e1, e2)
        assert e1 < 1e-9 and e2 < 1e-9, (name, e1, e2)
        live = env.episode_liveness.cpu().numpy()[:, 0]
        assert mism.any() or np.array_equal(live, ep["liveness"][-1]) or tg.any(), (name, live, ep["liveness"][-1])
print("quick parity ok, worst rel err", worst)
