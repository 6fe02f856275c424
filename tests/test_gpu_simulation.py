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
        for key in ("episodes", "interesting", "sum_t", "sum_t2", "sum_score", "sum_score2", "env_steps", "sum_t_interesting", "sum_t2_interesting"):
            assert got[key] == want[key], key
    assert "interesting test(s)" in summary.console_message()


def test_compat_view_info_matches_reference():
    """env.info() of the single-environment compat view (one cavgym_info launch) against the reference's own info()
    along the recorded run: polygons as ConvexQuadrilateral objects, road angles as float / None."""
    from helpers import load_info_golden
    from cavgym_b200.config import make_config
    meta, state, polygons, angles = load_info_golden("info_pedestrians2_rc_seed12")
    config = make_config(copy.deepcopy(meta["config"]))
    _, env, agents, _ = config.setup()
    observation = env.reset()
    info = env.info()
    for agent in agents:
        agent.reset()
    for t in range(60):
        assert state_err(np.array(observation), state[t]) < 1e-9
        got = np.array([[x for x, _ in polygon] + [y for _, y in polygon] for polygon in info["body_polygons"]])
        assert np.max(np.abs(got - polygons[t])) < 1e-8
        for mine, theirs in zip(info["road_angles"], angles[t]):
            assert (mine is None) == bool(np.isnan(theirs))
            if mine is not None:
                assert abs(mine - theirs) < 1e-9
        joint_action = [agent.choose_action(observation, space, info) for agent, space in zip(agents, env.action_space)]
        previous = observation
        observation, reward, done, info = env.step(joint_action)
        for agent, action, r in zip(agents, joint_action, reward):
            agent.process_feedback(previous, action, observation, r)


def test_batched_simulation_writes_one_episode_row_per_episode(tmp_path):
    """episode.log at batch scale (reporting.py:157-158): the rows the device ring collected are exactly the episodes the
    counters saw — same count, same interesting set, same sums — and every row is an episode the oracle also finished,
    env by env and episode by episode."""
    from cavgym_b200.config import make_config
    from cavgym_b200.simulation import BatchedSimulation
    from oracle.oracle import Oracle
    n = 1024
    log = tmp_path / "episode.log"
    config = make_config(dict(copy.deepcopy(STOCK), episodes=600, seed=5, episode_log=str(log),
                              tester_config={"option": "random-constrained", "epsilon": 0.5}))
    sim = BatchedSimulation(config, n, chunk=150)
    sim.keep_rows = True
    summary = sim.run()
    stats = sim.env.stats()
    rows = sim.episode_rows
    assert sim.dropped_rows == 0 and len(rows) == stats["episodes"] == summary.episodes
    assert sum(r.interesting for _, _, r in rows) == stats["interesting"]
    assert sum(r.time.timesteps for _, _, r in rows) == stats["sum_t"]
    assert sum(int(r.score) for _, _, r in rows if r.interesting) == stats["sum_score"]
    assert len({(env, episode) for env, episode, _ in rows}) == len(rows)          # no episode twice
    lines = log.read_text().strip().splitlines()
    assert len(lines) == len(rows) and lines[0].split(",")[0] == "1" and len(lines[0].split(",")) == 5
    # the oracle on the same Philox stream, stepped as far: per (env, episode) the same length and winner class
    meta, _ = load_golden("pedestrians_rc_eps05_seed1")
    oracle = Oracle(compile_from_meta(meta, mode="device"), n, seed=5, threads=8)
    oracle.reset()
    done_at, episode_no = {}, np.ones(n, np.int64)      # the first reset starts every env's episode 1
    for _ in range(sim.steps_run):
        oracle.rollout(1, auto_reset=False)
        ended = np.nonzero(oracle.done_latch)[0]
        for e in ended:
            done_at[(int(e), int(episode_no[e]))] = (int(oracle.timestep[e]), int(oracle.winner_latch[e]) > 0)
        if len(ended):
            mask = np.zeros(n, np.uint8)
            mask[ended] = 1
            oracle.reset(mask=mask)
            episode_no[ended] += 1
    if stats["tangent"] == 0:
        for env, episode, result in rows:
            assert done_at[(env, episode)] == (result.time.timesteps, result.interesting), (env, episode)


def test_episode_ring_reports_what_it_dropped():
    """A full ring drops rows but not counts: drained + dropped = episodes scored since the last drain, and the ring is empty
    and usable again afterwards."""
    from helpers import compile_from_meta as compile_, load_golden as load
    from cavgym_b200 import BatchedCAVEnv
    meta, _ = load("pedestrians_rc_eps05_seed1")
    env = BatchedCAVEnv(None, None, None, num_envs=512, dtype="float64", compiled=compile_(meta, mode="device"), seed=2)
    env.set_episode_log(100)
    env.reset()
    env.rollout(1000, auto_reset=True)
    episodes = env.stats()["episodes"]
    rows, dropped = env.drain_episodes()
    assert episodes > 512 and len(rows) == 100 and dropped == episodes - 100
    assert np.all(rows["episode"] >= 1) and np.all((rows["timesteps"] >= 1) & (rows["timesteps"] <= 1000))
    assert np.all((rows["env"] >= 0) & (rows["env"] < 512))
    env.rollout(50, auto_reset=True)
    rows2, dropped2 = env.drain_episodes()
    assert len(rows2) + dropped2 == env.stats()["episodes"] - episodes
    env.set_episode_log(0)
    env.rollout(50, auto_reset=True)      # ring off: scoring continues, nothing is appended
    with pytest.raises(Exception):
        env.drain_episodes()
    env.close()
