"""Generate tests/golden/learn_*.npz: runs of the UNMODIFIED reference (imported through oracle/refload.py) with its
host-side learning / arbitration agents — build container only.

    python -m oracle.gen_learning_golden            # regenerate
    python -m oracle.gen_learning_golden --check    # regenerate in memory and compare with the committed files

  learn_qego_*      QLearningEgoAgent (examples/agents/ego.py:16-145) as the ego, `config.json`'s own ego_config block;
                    extra[t] = the feature weights (opponent-major, feature names sorted) followed by alpha, after step t
  learn_election_*  ElectionAgent testers arbitrated by Election (examples/agents/pedestrian.py:94-116,
                    examples/election.py:4-57); extra[t] = voting flag, crossing flag per body, then active_player (-1 = None)

The q-learning TESTER (pedestrian.py:119-251) has no fixture: the reference raises TypeError at its first
process_feedback (`self.alpha` is the config's LinSpace object, pedestrian.py:128, 247) — reproduced by this script.
"""
import argparse
import copy
import json
import os
import sys

import numpy as np

from . import refload, trace
from .gen_golden import GOLDEN_DIR, same

M2PX = 16


def stock_q_learning():
    """ego_config of the reference's own config.json (config.json:20-40)."""
    with open(os.path.join(refload.REFERENCE_ROOT, "config.json")) as fh:
        return json.load(fh)["ego_config"]


def all_features(block):
    block = copy.deepcopy(block)
    for name in ("distance_x", "distance_y", "distance", "relative_angle", "heading"):
        block["feature_config"][name] = True
    block["epsilon"] = 0.3
    block["alpha"] = {"start": 0.8, "stop": 0.2, "num_steps": 300}
    return block


def cases():
    q = stock_q_learning()
    base = refload.stock_config_dict
    out = {
        # config.json as shipped except mode -> headless (and 3 episodes): the stock scenario with its q-learning ego
        "learn_qego_rc_seed0": dict(base(tester="random-constrained", epsilon=0.01, seed=0, episodes=3), ego_config=q),
        "learn_qego3_rc_seed21": dict(base(tester="random-constrained", epsilon=0.05, seed=21, episodes=2, num_pedestrians=3,
                                           max_timesteps=400), ego_config=all_features(q)),
        "learn_election3_seed22": base(tester="election", threshold=M2PX * 34, seed=22, episodes=4, num_pedestrians=3),
        "learn_election5_seed23": base(tester="election", threshold=M2PX * 60, seed=23, episodes=3, num_pedestrians=5),
    }
    return out


def q_extra(env, agents, simulation):
    ego = agents[0]
    weights = [ego.feature_weights[i][f] for i in ego.opponent_indexes for f in sorted(ego.feature_bounds)]
    return weights + [ego.alpha]


def election_extra(env, agents, simulation):
    flags = []
    for agent in agents:
        flags += [float(getattr(agent, "voting", False)), float(getattr(agent, "crossing", False))]
    active = simulation.election.active_player
    return flags + [-1.0 if active is None else float(active)]


def build(name, cfg):
    meta, episodes = trace.record(cfg, extra=q_extra if name.startswith("learn_qego") else election_extra)
    payload = {"meta": np.frombuffer(json.dumps(meta, sort_keys=True).encode(), dtype=np.uint8), "n_episodes": np.asarray(len(episodes))}
    for e, ep in enumerate(episodes):
        for key, value in ep.items():
            payload[f"ep{e}_{key}"] = value
    return meta, payload


def tester_q_learning_is_broken():
    cfg = dict(refload.stock_config_dict(seed=0, episodes=1, max_timesteps=5), tester_config=stock_q_learning())
    try:
        trace.record(cfg)
    except TypeError as exc:
        return "LinSpace" in str(exc)
    return False


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--check", action="store_true")
    args = parser.parse_args(argv)
    failures = 0
    for name, cfg in cases().items():
        meta, payload = build(name, cfg)
        n = int(payload["n_episodes"])
        summary = (f"agents={meta['agent_classes']} lengths={[int(payload[f'ep{e}_done'].shape[0]) for e in range(n)]} "
                   f"winners={[int(payload[f'ep{e}_winner'][-1]) for e in range(n)]}")
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        if args.check:
            ok = os.path.isfile(path) and same(dict(np.load(path)), payload)
            failures += not ok
            print(("OK   " if ok else "DIFF ") + name, summary)
        else:
            np.savez_compressed(path, **payload)
            print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)", summary)
    print("reference q-learning tester raises TypeError at its first process_feedback:", tester_q_learning_is_broken())
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
