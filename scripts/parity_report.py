"""Parity numbers per golden fixture and precision, as measured (not as asserted): strict and trajectory-scale state error,
reward error, near-tangent flags, event and liveness agreement of cavgym_replay against the reference traces.

    python scripts/parity_report.py > profiles/r2_parity_report.txt

  strict      max over steps of |dp| / max(1, |p|), |dv| / max(1, |v|), angle difference / max(1, |theta|)   (helpers.state_err)
  trajectory  the same with the position / velocity scale taken as the largest magnitude reached so far in the episode
              (helpers.state_err_trajectory)
  flagged     steps the engine marked near-tangent (|decision margin| < tau: 1e-7 px in fp64, 0.05 px in fp32)
  mismatch    steps whose done / winner differ from the reference (all must be flagged); the comparison of an episode ends
              at its first mismatch (after an event diverges the two runs are different episodes)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import GOLDEN_CASES, LEARN_CASES, compile_from_meta, load_golden, rel_err, soa, state_err, state_err_trajectory  # noqa: E402


def main():
    from cavgym_b200 import BatchedCAVEnv
    print(f"{'fixture':34s} {'dtype':8s} {'steps':>6s} {'flagged':>8s} {'frac':>8s} {'mismatch':>8s} {'unflagged':>9s} "
          f"{'strict':>10s} {'trajectory':>10s} {'reward':>10s} {'liveness':>9s}")
    worst = {}
    for name in GOLDEN_CASES + LEARN_CASES:
        meta, episodes = load_golden(name)
        for dtype in ("float64", "float32"):
            env = BatchedCAVEnv(None, None, None, num_envs=1, dtype=dtype, compiled=compile_from_meta(meta), device="cuda:0")
            steps = flagged = mismatches = unflagged = live_equal = 0
            strict = traj_err = reward_err = 0.0
            for ep in episodes:
                env.reset(init_state=soa(ep["init_state"][None]))
                env.set_global_timestep(int(ep["t_global_start"]))
                out = env.replay(ep["actions"][..., None])
                got = {k: (v.double() if v.dtype.is_floating_point else v).cpu().numpy()[..., 0] for k, v in out.items()}
                t_len = ep["actions"].shape[0]
                tangent = got["tangent"].astype(bool)
                mismatch = (got["done"] != ep["done"]) | (got["winner"] != ep["winner"])
                steps += t_len
                flagged += int(tangent.sum())
                mismatches += int(mismatch.sum() > 0)
                unflagged += int((mismatch & ~tangent).sum())
                if mismatch.any():
                    t_len = int(np.nonzero(mismatch)[0][0])
                else:
                    live_equal += int(np.array_equal(env.episode_liveness.cpu().numpy()[:, 0], ep["liveness"][-1]))
                strict = max(strict, state_err(got["state"][:t_len], ep["state"][:t_len]))
                traj_err = max(traj_err, state_err_trajectory(got["state"][:t_len], ep["state"][:t_len]))
                reward_err = max(reward_err, rel_err(got["reward"][:t_len], ep["reward"][:t_len]))
            env.close()
            print(f"{name:34s} {dtype:8s} {steps:6d} {flagged:8d} {flagged / steps:8.4f} {mismatches:8d} {unflagged:9d} "
                  f"{strict:10.2e} {traj_err:10.2e} {reward_err:10.2e} {live_equal:4d}/{len(episodes) - mismatches:<4d}")
            w = worst.setdefault(dtype, {"strict": 0.0, "trajectory": 0.0, "reward": 0.0, "frac": 0.0, "unflagged": 0})
            w["strict"], w["trajectory"], w["reward"] = max(w["strict"], strict), max(w["trajectory"], traj_err), max(w["reward"], reward_err)
            w["frac"], w["unflagged"] = max(w["frac"], flagged / steps), w["unflagged"] + unflagged
    for dtype, w in worst.items():
        print(f"worst {dtype}: strict {w['strict']:.3e}  trajectory {w['trajectory']:.3e}  reward {w['reward']:.3e}  "
              f"flagged fraction {w['frac']:.4f}  unflagged mismatches {w['unflagged']}")


if __name__ == "__main__":
    main()
