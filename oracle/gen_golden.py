"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference through oracle/refload.py) — build container only.

    python -m oracle.gen_golden            # regenerate every fixture
    python -m oracle.gen_golden --check    # regenerate in memory and compare with the committed files

The fixtures pin (i) oracle/cavgym_oracle.c and (ii) the CUDA path, on replayed
joint actions.  They depend on the stand-ins' restatement of gym 0.17.2 seeding
and of Shapely's predicates (see oracle/standins/): they pin OUR oracle to the
reference's source, not to upstream gym/GEOS binaries.
"""
import argparse
import json
import math
import os
import sys

import numpy as np

from . import refload, trace

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

M2PX = 16

# name -> kwargs of refload.stock_config_dict
CASES = {
    # BASELINE C1 (x10 episodes): seed 0, eps 0.01 -> 901,901,901,901,715*,901,901,778*,901,901
    "pedestrians_rc_seed0": dict(scenario="pedestrians", tester="random-constrained", epsilon=0.01, seed=0, episodes=10),
    # experiments.py:15 epsilon
    "pedestrians_rc_eps05_seed1": dict(scenario="pedestrians", tester="random-constrained", epsilon=0.5, seed=1, episodes=4),
    "pedestrians3_rc_seed2": dict(scenario="pedestrians", tester="random-constrained", epsilon=0.02, seed=2, episodes=3,
                                  num_pedestrians=3),
    "pedestrians_proximity_seed3": dict(scenario="pedestrians", tester="proximity", threshold=M2PX * 34, seed=3, episodes=4),
    "pedestrians_random_all_seed4": dict(scenario="pedestrians", tester="random", epsilon=0.05, seed=4, episodes=3,
                                         collisions="all", offroad=True, num_pedestrians=2),
    "pedestrians_random_none_seed5": dict(scenario="pedestrians", tester="random", epsilon=0.1, seed=5, episodes=2,
                                          collisions="none", zones=False, ego="random", ego_epsilon=0.05),
    "crossroads_random_all_seed6": dict(scenario="crossroads", tester="random", epsilon=0.05, seed=6, episodes=3,
                                        collisions="all", offroad=True),
    "crossroads_random_ego_seed7": dict(scenario="crossroads", tester="random", epsilon=0.02, seed=7, episodes=2,
                                        ego="random", ego_epsilon=0.02),
    "busstop_random_all_seed8": dict(scenario="bus-stop", tester="random", epsilon=0.01, seed=8, episodes=3,
                                     collisions="all", offroad=True),
    "busstop_noop_seed9": dict(scenario="bus-stop", tester="noop", seed=9, episodes=1, collisions="all"),
    "pelican_random_all_seed10": dict(scenario="pelican-crossing", tester="random", epsilon=0.05, seed=10, episodes=3,
                                      collisions="all", offroad=True),
    "pelican_random_ego_seed11": dict(scenario="pelican-crossing", tester="random", epsilon=0.01, seed=11, episodes=2),
}


def build_case(name):
    cfg = refload.stock_config_dict(**CASES[name])
    meta, episodes = trace.record(cfg)
    payload = {"meta": np.frombuffer(json.dumps(meta, sort_keys=True).encode(), dtype=np.uint8)}
    payload["n_episodes"] = np.asarray(len(episodes))
    for e, ep in enumerate(episodes):
        for key, value in ep.items():
            payload[f"ep{e}_{key}"] = value
    return meta, payload


def build_geometry():
    rows, zones = trace.geometry_vectors(seed=0, count=240)
    return {"pairs": rows, "zones": zones}


def same(a, b):
    if set(a.keys()) != set(b.keys()):
        return False
    return all(np.array_equal(a[k], b[k], equal_nan=True) if a[k].dtype.kind == "f" else np.array_equal(a[k], b[k])
               for k in a.keys())


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--check", action="store_true")
    parser.add_argument("--only", nargs="*", default=None)
    args = parser.parse_args(argv)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    failures = 0
    names = list(CASES) if args.only is None else args.only
    for name in names + ([] if args.only else ["geometry_kat"]):
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        if name == "geometry_kat":
            payload, summary = build_geometry(), ""
        else:
            meta, payload = build_case(name)
            lens = [int(payload[f"ep{e}_done"].shape[0]) for e in range(int(payload["n_episodes"]))]
            winners = [int(payload[f"ep{e}_winner"][-1]) for e in range(int(payload["n_episodes"]))]
            summary = f"bodies={meta['body_classes']} lengths={lens} winners={winners}"
        if args.check:
            ok = os.path.isfile(path) and same(dict(np.load(path)), payload)
            failures += not ok
            print(("OK   " if ok else "DIFF ") + name, summary)
        else:
            np.savez_compressed(path, **payload)
            print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)", summary)
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
