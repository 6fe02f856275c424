"""config.json -> Config.setup() -> Simulation.run() on the single-environment compat view, against the reference's own
run of the same config (golden fixture produced by the unmodified reference, oracle/gen_golden.py), and the batched
counterpart against the CPU oracle."""
import copy

import numpy as np
import pytest

from helpers import compile_from_meta, load_golden, state_err
from test_config import STOCK

pytestmark = pytest.mark.gpu


def test_simulation_run_reproduces_the_reference_episodes():
    """Stock config.json (ego noop, headless), seed 0: same RNG stream (seeding.np_random), same host-side agents, every
    transition on the GPU -> the reference's episode lengths, interesting flags, scores and final states."""
    from cavgym_b200.config import make_config
    from cavgym_b200.simulation import Simulation
    meta, episodes = load_golden("pedestrians_rc_seed0")
    assert meta["config"]["seed"] == 0 and meta["config"]["tester_config"]["epsilon"] == 0.01
    n_episodes = 3
    config = make_config(dict(copy.deepcopy(STOCK), episodes=n_episodes))
    np_seed, env, agents, keyboard_agent = config.setup()
    assert np_seed == 0 and keyboard_agent is None
    assert [type(a).__name__ for a in agents] == ["NoopAgent", "RandomConstrainedAgent"]
    results, summary = Simulation(env, agents, config, keyboard_agent).run()
    assert summary.episodes == n_episodes
    for row, ep in zip(results, episodes):
        assert row.time.timesteps == ep["actions"].shape[0]
        winner = int(ep["winner"][-1])
        assert row.interesting == (winner > 0)
        if row.interesting:
            assert row.score == -int(ep["liveness"][-1][1:].sum())
    # the bodies end where the reference's bodies ended in the last episode run
    last = episodes[n_episodes - 1]["state"][-1]
    got = np.array([list(body.state) for body in env.bodies])
    assert state_err(got, last) < 1e-9


def test_batched_simulation_matches_oracle_counters():
    from cavgym_b200.config import make_config
    from cavgym_b200.simulation import BatchedSimulation
    from oracle.oracle import Oracle
    n = 2048
    config = make_config(dict(copy.deepcopy(STOCK), episodes=n, seed=3, tester_config={"option": "random-constrained", "epsilon": 0.5}))
    sim = BatchedSimulation(config, n, chunk=200)
    summary = sim.run()
    got = sim.env.stats()
    assert got["episodes"] >= n and summary.episodes == got["episodes"] and summary.interesting == got["interesting"]
    meta, _ = load_golden("pedestrians_rc_eps05_seed1")
    oracle = Oracle(compile_from_meta(meta, mode="device"), n, seed=3, threads=8)
    oracle.reset()
    oracle.rollout(sim.steps_run, auto_reset=True)
    want = oracle.stats()
    if got["tangent"] == 0 and want["tangent"] == 0:
        for key in ("episodes", "interesting", "sum_t", "sum_t2", "sum_score", "sum_score2", "env_steps"):
            assert got[key] == want[key], key
    assert "interesting test(s)" in summary.console_message()
