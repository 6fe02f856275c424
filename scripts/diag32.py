import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from helpers import *
from cavgym_b200 import BatchedCAVEnv
for name in ["pedestrians_rc_seed0","pedestrians3_rc_seed2","busstop_random_all_seed8","pelican_random_all_seed10","crossroads_random_ego_seed7"]:
    meta, eps = load_golden(name)
    env = BatchedCAVEnv(None,None,None,num_envs=1,dtype="float32",compiled=compile_from_meta(meta))
    for ei, ep in enumerate(eps):
        env.reset(init_state=soa(ep["init_state"][None]))
        env.set_global_timestep(int(ep["t_global_start"]))
        out = env.replay(ep["actions"][...,None])
        st = out["state"].double().cpu().numpy()[...,0]
        T = st.shape[0]
        err = np.abs(st - ep["state"])
        d = st[...,3]-ep["state"][...,3]; err[...,3] = np.abs(np.arctan2(np.sin(d),np.cos(d)))
        worst = np.unravel_index(np.argmax(err/np.maximum(1,np.abs(ep["state"]))), err.shape)
        dn = out["done"].cpu().numpy()[:,0]; wn = out["winner"].cpu().numpy()[:,0]; tg = out["tangent"].cpu().numpy()[:,0]
        mism = np.nonzero((dn!=ep["done"])|(wn!=ep["winner"]))[0]
        print(name, ei, "T",T,"maxabs per comp", err.reshape(-1,4).max(0), "worst", worst, st[worst], ep["state"][worst], "action", ep["actions"][worst[0], worst[1]], "mismatch", mism[:3], "flagged", int(tg.sum()))
