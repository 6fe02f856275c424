"""Generate tests/golden/info_*.npz: CAVEnv.info() (reference library/environment.py:106-117 — body_polygons and
road_angles) along runs of the UNMODIFIED reference imported through oracle/refload.py — build container only.

    python -m oracle.gen_info_golden            # regenerate
    python -m oracle.gen_info_golden --check    # regenerate in memory and compare with the committed files

Per recorded step: state [M, 4], body_polygons [M, 8] (x of rear_left, front_left, front_right, rear_right, then y), and
road_angles [M] with NaN where the reference returns None (the body's box intersects the major road).
"""
import argparse
import json
import os
import sys

import numpy as np

from . import refload
from .trace import _state_row

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = {
    "info_pedestrians2_rc_seed12": dict(scenario="pedestrians", tester="random-constrained", epsilon=0.05, seed=12, num_pedestrians=2),
    "info_pelican_random_seed13": dict(scenario="pelican-crossing", tester="random", epsilon=0.05, seed=13, collisions="none", zones=False),
    "info_crossroads_random_seed14": dict(scenario="crossroads", tester="random", epsilon=0.05, seed=14, collisions="none", zones=False,
                                          ego="random", ego_epsilon=0.05),
}
MAX_STEPS = 400


def build(name):
    mods = refload.load()
    cfg = refload.stock_config_dict(**CASES[name])
    config = mods["config"].make_config(json.loads(json.dumps(cfg)))
    _, env, agents, _ = config.setup()
    for agent in agents:   # a RandomAgent ego is built on an unseeded import-time generator (config.py:305-310): seed it, like trace.record
        own = getattr(agent, "np_random", None)
        if own is not None and own is not env.np_random:
            agent.np_random = np.random.RandomState(10_000 + int(cfg.get("seed") or 0))
    state = env.reset()
    info = env.info()
    for agent in agents:
        agent.reset()
    states, polygons, angles = [], [], []

    def take(info):
        states.append([_state_row(body.state) for body in env.bodies])
        polygons.append([[x for x, _ in polygon] + [y for _, y in polygon] for polygon in info["body_polygons"]])
        angles.append([float("nan") if a is None else float(a) for a in info["road_angles"]])

    take(info)
    for _ in range(MAX_STEPS):
        joint_action = [agent.choose_action(state, space, info) for agent, space in zip(agents, env.action_space)]
        previous = state
        state, reward, done, info = env.step(joint_action)
        for agent, action, r in zip(agents, joint_action, reward):
            agent.process_feedback(previous, action, state, r)
        take(info)
        if done:
            break
    meta = {"config": cfg, "n_bodies": len(env.bodies), "body_classes": [type(b).__name__ for b in env.bodies]}
    return {"meta": np.frombuffer(json.dumps(meta, sort_keys=True).encode(), dtype=np.uint8), "state": np.array(states),
            "body_polygons": np.array(polygons), "road_angles": np.array(angles)}


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--check", action="store_true")
    args = parser.parse_args(argv)
    failures = 0
    for name in CASES:
        payload = build(name)
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        angles = payload["road_angles"]
        summary = f"steps={angles.shape[0]} bodies={angles.shape[1]} off-road samples={int(np.isfinite(angles).sum())}"
        if args.check:
            old = dict(np.load(path)) if os.path.isfile(path) else {}
            ok = set(old) == set(payload) and all(np.array_equal(old[k], payload[k], equal_nan=payload[k].dtype.kind == "f") for k in payload)
            failures += not ok
            print(("OK   " if ok else "DIFF ") + name, summary)
        else:
            np.savez_compressed(path, **payload)
            print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)", summary)
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
