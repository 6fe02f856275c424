"""The CPU oracle (oracle/cavgym_oracle.c) against the golden traces recorded from the unmodified reference
(tests/golden/*.npz, made by oracle/gen_golden.py).  State must be BIT-EQUAL (same operation order, same libm),
events exact; rewards within 1e-12 relative (the clipped-area quotient is the only inexact quantity)."""
import numpy as np
import pytest

from helpers import GOLDEN_CASES, compile_from_meta, load_golden, soa
from oracle.oracle import Oracle


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_replays_reference_trace(name):
    meta, episodes = load_golden(name)
    oracle = Oracle(compile_from_meta(meta), 1)
    for ep in episodes:
        oracle.reset(init_state=soa(ep["init_state"][None]))
        oracle.set_global_timestep(int(ep["t_global_start"]))
        state, reward, done, winner, _ = oracle.replay(ep["actions"][..., None])
        assert np.array_equal(state[..., 0], ep["state"]), "state must be bit-equal to the reference"
        assert np.array_equal(done[:, 0], ep["done"])
        assert np.array_equal(winner[:, 0], ep["winner"])
        assert np.array_equal(oracle.liveness[:, 0], ep["liveness"][-1])
        scale = np.maximum(1.0, np.abs(ep["reward"]))
        assert np.max(np.abs(reward[..., 0] - ep["reward"]) / scale) < 1e-12


def test_first_reset_and_rewards_match_survey_probe():
    """SURVEY §8c known answers for seed 0: first reset state and the step-1 rewards."""
    _, episodes = load_golden("pedestrians_rc_seed0")
    ep = episodes[0]
    assert ep["init_state"].tolist() == [[0.0, 29.2, 108.0, 0.0], [1166.0896011871268, 94.1116259080894, 22.4, 0.0]]
    assert ep["reward"][0].tolist() == [0.004545454545454408, 3.9954545454545456]
    assert [len(e["done"]) for e in episodes] == [901, 901, 901, 901, 715, 901, 901, 778, 901, 901]
    assert [int(e["winner"][-1]) for e in episodes] == [0, 0, 0, 0, 1, 0, 0, 1, 0, 0]


def test_oracle_agents_follow_reference_with_replayed_draws():
    """On-'device' agent logic of the oracle (crossing state machine, steering inverse, RandomAgent) fed with the
    MT19937 draws the reference consumed: actions, agent state and events must follow the reference trace."""
    for name in ("pedestrians_rc_seed0", "pedestrians_rc_eps05_seed1", "pedestrians3_rc_seed2", "pedestrians_proximity_seed3",
                 "pedestrians_random_all_seed4", "pelican_random_all_seed10", "crossroads_random_all_seed6"):
        meta, episodes = load_golden(name)
        oracle = Oracle(compile_from_meta(meta, mode="device"), 1)
        for ep in episodes:
            oracle.reset(init_state=soa(ep["init_state"][None]))
            oracle.set_global_timestep(int(ep["t_global_start"]))
            for t in range(ep["actions"].shape[0]):
                oracle.set_uniform_override(np.nan_to_num(ep["draws"][t], nan=0.5)[..., None])
                state, reward, done, winner, _ = oracle.step(None)
                assert np.array_equal(oracle.actions_taken[..., 0], ep["actions"][t]), (name, t)
                assert np.array_equal(state[..., 0], ep["state"][t]), (name, t)
                assert bool(done[0]) == bool(ep["done"][t]) and int(winner[0]) == int(ep["winner"][t])
                crossing = ~np.isnan(ep["agent_state"][t]).all(axis=1)
                got = oracle.agent_state[..., 0]
                assert np.array_equal(got[crossing], ep["agent_state"][t][crossing], equal_nan=True), (name, t)


def test_philox_known_answers():
    from oracle import oracle as o
    assert o.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert o.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert o.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


@pytest.mark.parametrize("name", __import__("helpers").INFO_CASES)
def test_oracle_info_matches_reference(name):
    """CAVEnv.info() (environment.py:106-117) recorded from the unmodified reference: body polygons bit-equal, road angles
    bit-equal where defined, None (NaN) for exactly the same bodies."""
    from helpers import load_info_golden
    meta, state, polygons, angles = load_info_golden(name)
    t_len, m = state.shape[0], meta["n_bodies"]
    oracle = Oracle(compile_from_meta(meta), t_len)           # one env per recorded step
    oracle.reset(init_state=np.ascontiguousarray(np.transpose(state, (1, 2, 0))))
    got_polygons, got_angles = oracle.info()
    assert np.array_equal(np.transpose(got_polygons, (2, 0, 1)), polygons)
    assert np.array_equal(np.isnan(got_angles.T), np.isnan(angles))
    assert np.array_equal(got_angles.T, angles, equal_nan=True)
    assert np.isfinite(angles).any() and np.isnan(angles).any()
