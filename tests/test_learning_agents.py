"""Host-side learning / arbitration agents against runs of the UNMODIFIED reference
(tests/golden/learn_*.npz, made by oracle/gen_learning_golden.py).

CPU: QLearningEgoAgent is fed the states, rewards and RandomState draws the reference's agent saw and must choose the same
actions and end every step with the same weights — bit for bit; ElectionAgent + Election likewise reproduce the executed
joint actions, flags and the active player.  The oracle replays the recorded episodes (4 and 6 bodies) bit-equal.
GPU: the same configs run end to end through Config.setup + Simulation on the compat view (every transition a CUDA launch,
the shared RandomState stream reproduced by seeding.np_random), and the tensor-API learner is checked against the host class.
"""
import copy
import json

import numpy as np
import pytest

from helpers import LEARN_CASES, compile_from_meta, load_golden, soa, state_err

Q_CASES = [c for c in LEARN_CASES if c.startswith("learn_qego")]
ELECTION_CASES = [c for c in LEARN_CASES if c.startswith("learn_election")]


class ReplayRandom:
    """Serves recorded [0, 1) draws through the RandomState methods the agents call (oracle/trace.py logs integer draws as
    (index + 0.5) / n)."""

    def __init__(self):
        self.queue = []

    def load(self, draws):
        assert not self.queue, "the previous step left draws unused"
        self.queue = [float(d) for d in draws if not np.isnan(d)]

    def uniform(self, low=0.0, high=1.0):
        return low + (high - low) * self.queue.pop(0)

    def choice(self, options):
        options = list(options)
        return options[int(self.queue.pop(0) * len(options))]


def build(meta, np_random=None):
    """Config.setup() of the fixture's config without touching a GPU: the compat env only launches when it is stepped."""
    from cavgym_b200.config import make_config
    config = make_config(copy.deepcopy(meta["config"]))
    _, env, agents, _ = config.setup()
    if np_random is not None:
        for agent in agents:
            if getattr(agent, "np_random", None) is not None:
                agent.np_random = np_random
    return config, env, agents


@pytest.mark.parametrize("name", Q_CASES)
def test_q_learning_ego_follows_the_reference_bit_for_bit(name):
    meta, episodes = load_golden(name)
    rng = ReplayRandom()
    _, env, agents, = build(meta, rng)
    ego = agents[0]
    assert type(ego).__name__ == meta["agent_classes"][0] == "QLearningEgoAgent"
    names = sorted(ego.feature_bounds)
    for ep in episodes:
        state = ep["init_state"].tolist()
        ego.reset()
        for t in range(ep["actions"].shape[0]):
            rng.load(ep["draws"][t][0])
            action = ego.choose_action(state, env.action_space[0])
            assert [float(action[0]), float(action[1])] == ep["actions"][t][0].tolist(), (name, t)
            previous, state = state, ep["state"][t].tolist()
            ego.process_feedback(previous, action, state, float(ep["reward"][t][0]))
            weights = [ego.feature_weights[i][f] for i in ego.opponent_indexes for f in names]
            assert weights + [ego.alpha] == ep["extra"][t].tolist(), (name, t)
    assert any(w != 0.0 for w in weights)


@pytest.mark.parametrize("name", ELECTION_CASES)
def test_election_reproduces_the_reference_arbitration(name):
    from cavgym_b200.examples.election import Election
    from cavgym_b200.library.bodies import DynamicBodyState
    from cavgym_b200.library.geometry import Point
    meta, episodes = load_golden(name)
    _, env, agents = build(meta)
    env.np_random = ReplayRandom()            # Election draws from env.np_random; a single winner never indexes the draw
    env.np_random.choice = lambda options: list(options)[0] if len(options) == 1 else pytest.fail("tie between voters")
    election = Election(env, agents)
    assert election.electorate == list(range(1, len(agents)))
    for ep in episodes:
        state = ep["init_state"].tolist()
        for agent in agents:
            agent.reset()
        for t in range(ep["actions"].shape[0]):
            for body, row in zip(env.bodies, state):   # Election reads positions from the bodies (election.py:44)
                body.state = DynamicBodyState(Point(row[0], row[1]), row[2], row[3])
            joint = [agent.choose_action(state, space) for agent, space in zip(agents, env.action_space)]
            joint = election.result(state, joint)
            assert [[float(a[0]), float(a[1])] for a in joint] == ep["actions"][t].tolist(), (name, t)
            previous, state = state, ep["state"][t].tolist()
            for agent, action, reward in zip(agents, joint, ep["reward"][t]):
                agent.process_feedback(previous, action, state, float(reward))
            flags = []
            for agent in agents:
                flags += [float(getattr(agent, "voting", False)), float(getattr(agent, "crossing", False))]
            active = -1.0 if election.active_player is None else float(election.active_player)
            assert flags + [active] == ep["extra"][t].tolist(), (name, t)


@pytest.mark.parametrize("name", LEARN_CASES)
def test_oracle_replays_the_learning_runs(name):
    """More reference steps for the oracle (and, on the GPU, for the engine): pedestrians scenarios of 2, 4 and 6 bodies."""
    from oracle.oracle import Oracle
    meta, episodes = load_golden(name)
    oracle = Oracle(compile_from_meta(meta), 1)
    for ep in episodes:
        oracle.reset(init_state=soa(ep["init_state"][None]))
        oracle.set_global_timestep(int(ep["t_global_start"]))
        state, reward, done, winner, _ = oracle.replay(ep["actions"][..., None])
        assert np.array_equal(state[..., 0], ep["state"])
        assert np.array_equal(done[:, 0], ep["done"]) and np.array_equal(winner[:, 0], ep["winner"])
        assert np.max(np.abs(reward[..., 0] - ep["reward"]) / np.maximum(1.0, np.abs(ep["reward"]))) < 1e-12


def test_q_learning_tester_and_keyboard_are_refused_with_the_reason():
    from cavgym_b200.config import make_config
    meta, _ = load_golden(Q_CASES[0])
    cfg = copy.deepcopy(meta["config"])
    cfg["tester_config"], cfg["ego_config"] = cfg["ego_config"], {"option": "noop"}
    with pytest.raises(NotImplementedError, match="LinSpace"):
        make_config(cfg).setup()
    cfg = copy.deepcopy(meta["config"])
    cfg["ego_config"] = {"option": "keyboard"}
    with pytest.raises(NotImplementedError, match="interactive"):
        make_config(cfg).setup()


def test_stock_config_json_sets_up(capsys):
    """The reference's config.json AS SHIPPED (q-learning ego, render mode): parses, round-trips and sets up; render mode
    runs headless with a warning."""
    from cavgym_b200.config import make_config
    from test_config import STOCK
    meta, _ = load_golden("learn_qego_rc_seed0")
    shipped = dict(copy.deepcopy(STOCK), ego_config=meta["config"]["ego_config"],
                   mode_config={"option": "render", "episode_condition": 5, "video_dir": None}, verbosity="info")
    config = make_config(copy.deepcopy(shipped))
    assert json.loads(json.dumps(config.to_data())) == shipped
    _, env, agents, keyboard = config.setup()
    assert [type(a).__name__ for a in agents] == ["QLearningEgoAgent", "RandomConstrainedAgent"] and keyboard is None
    assert agents[0].np_random is env.np_random
    assert len(agents[0].available_actions) == 5 and sorted(agents[0].feature_bounds) == ["distance", "heading", "relative_angle"]


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["learn_qego_rc_seed0", "learn_election3_seed22"])
def test_simulation_with_learning_agents_reproduces_the_reference_run(name):
    """End to end on the compat view: same seed -> same spawn draws, exploration draws, actions, weights and episodes."""
    from cavgym_b200.simulation import Simulation
    meta, episodes = load_golden(name)
    config, env, agents = build(dict(meta, config=dict(meta["config"], episodes=2)))
    results, summary = Simulation(env, agents, config).run()
    assert [row.time.timesteps for row in results] == [ep["actions"].shape[0] for ep in episodes[:2]]
    assert [row.interesting for row in results] == [int(ep["winner"][-1]) > 0 for ep in episodes[:2]]
    got = np.array([list(body.state) for body in env.bodies])
    assert state_err(got, episodes[1]["state"][-1]) < 1e-9
    if name.startswith("learn_qego"):
        ego = agents[0]
        weights = [ego.feature_weights[i][f] for i in ego.opponent_indexes for f in sorted(ego.feature_bounds)]
        assert np.allclose(weights, episodes[1]["extra"][-1][:-1], rtol=1e-7, atol=1e-9)


def check_tensor_learner(name, dev):
    """BatchedQLearningEgoAgent with N = 1, fed the reference's states, rewards and draws: the reference's actions and,
    step for step, its weights (torch's cos / sin / atan2 differ from libm by ulps: 1e-8 relative, not bitwise)."""
    import torch
    from cavgym_b200.config import make_config
    from cavgym_b200.examples.agents.ego import BatchedQLearningEgoAgent
    from cavgym_b200.examples.constants import car_constants
    from cavgym_b200.examples.environments import pedestrians
    meta, episodes = load_golden(name)
    config = make_config(copy.deepcopy(meta["config"]))
    m = meta["n_bodies"]
    learner = BatchedQLearningEgoAgent(config.ego_config, car_constants, 1.0 / 60, m - 1, pedestrians.env_constants.viewer_width,
                                       pedestrians.env_constants.viewer_height, 1, dev)
    order = [learner.names.index(f) for f in sorted(learner.names)]
    steps = 0
    for ep in episodes:
        state = torch.tensor(ep["init_state"], dtype=torch.float64, device=dev).unsqueeze(-1)
        for t in range(ep["actions"].shape[0]):     # whole episodes: the weights carry over from one to the next
            draws = [d for d in ep["draws"][t][0] if not np.isnan(d)]
            explore = torch.tensor([draws[0]], dtype=torch.float64, device=dev)
            pick = torch.tensor([draws[1] if len(draws) > 1 else 0.0], dtype=torch.float64, device=dev)
            index, rows = learner.choose_action(state, explore, pick)
            assert rows[:, 0].tolist() == ep["actions"][t][0].tolist(), (name, t)
            previous, state = state, torch.tensor(ep["state"][t], dtype=torch.float64, device=dev).unsqueeze(-1)
            learner.process_feedback(previous, index, state, torch.tensor([ep["reward"][t][0]], dtype=torch.float64, device=dev))
            want = ep["extra"][t][:-1].reshape(m - 1, -1)
            assert np.allclose(learner.weights[:, order].cpu().numpy(), want, rtol=1e-8, atol=1e-9), (name, t)
            assert learner.alpha == ep["extra"][t][-1]
            steps += 1
        if steps > 1200:
            break
    assert steps > 500


def test_independent_learners_are_side_by_side_copies_of_the_reference_agent():
    """shared=False: every env is its own learner with its own alpha schedule, gamma and epsilon (the grid of experiments.py in
    one batch).  Three envs fed the inputs of one recorded run — two with the run's hyper-parameters, one with another
    gamma: the first two end every step with the reference's weights and learning rate, the third does not."""
    import torch
    from cavgym_b200.config import make_config
    from cavgym_b200.examples.agents.ego import BatchedQLearningEgoAgent
    from cavgym_b200.examples.constants import car_constants
    from cavgym_b200.examples.environments import pedestrians
    meta, episodes = load_golden("learn_qego3_rc_seed21")
    config = make_config(copy.deepcopy(meta["config"]))
    q, m, n = config.ego_config, meta["n_bodies"], 3
    learner = BatchedQLearningEgoAgent(q, car_constants, 1.0 / 60, m - 1, pedestrians.env_constants.viewer_width,
                                       pedestrians.env_constants.viewer_height, n, torch.device("cpu"), shared=False,
                                       alpha=([q.alpha.start] * n, [q.alpha.stop] * n, [q.alpha.num_steps] * n),
                                       gamma=[q.gamma, q.gamma, 0.5 * q.gamma], epsilon=[q.epsilon] * n)
    order = [learner.names.index(f) for f in sorted(learner.names)]
    for ep in episodes:
        state = torch.tensor(ep["init_state"]).unsqueeze(-1).expand(-1, -1, n).contiguous()
        for t in range(ep["actions"].shape[0]):
            draws = [d for d in ep["draws"][t][0] if not np.isnan(d)]
            index, rows = learner.choose_action(state, torch.tensor([draws[0]] * n), torch.tensor([draws[1] if len(draws) > 1 else 0.0] * n))
            assert rows[:, 0].tolist() == rows[:, 1].tolist() == ep["actions"][t][0].tolist()
            index[2] = index[0]      # keep the third learner on the recorded trajectory: only its update differs
            previous, state = state, torch.tensor(ep["state"][t]).unsqueeze(-1).expand(-1, -1, n).contiguous()
            learner.process_feedback(previous, index, state, torch.tensor([ep["reward"][t][0]] * n))
            want = ep["extra"][t][:-1].reshape(m - 1, -1)
            for e in (0, 1):
                assert np.allclose(learner.weights[e][:, order].numpy(), want, rtol=1e-8, atol=1e-9), (t, e)
            assert np.allclose(learner.alpha.numpy(), ep["extra"][t][-1], rtol=1e-12)
    assert not np.allclose(learner.weights[2][:, order].numpy(), want, rtol=1e-3)


@pytest.mark.parametrize("name", Q_CASES)
def test_tensor_api_learner_is_the_host_learner_at_one_env_cpu_tensors(name):
    import torch
    check_tensor_learner(name, torch.device("cpu"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", Q_CASES)
def test_tensor_api_learner_is_the_host_learner_at_one_env(name):
    import torch
    check_tensor_learner(name, torch.device("cuda", 0))


@pytest.mark.gpu
def test_batched_simulation_runs_the_stock_q_learning_ego():
    """python -m cavgym_b200 config.json --envs N with the reference's stock ego option: the learner drives the ego of
    every env through cavgym_step, testers act on the device, finished envs are reset, the weights move."""
    from cavgym_b200.config import make_config
    from cavgym_b200.simulation import BatchedSimulation
    meta, _ = load_golden("learn_qego_rc_seed0")
    config = make_config(dict(copy.deepcopy(meta["config"]), episodes=64, max_timesteps=200,
                              tester_config={"option": "random-constrained", "epsilon": 0.2}))
    sim = BatchedSimulation(config, 256, chunk=50)
    summary = sim.run()
    stats = sim.env.stats()
    assert stats["episodes"] >= 64 and stats["errors"] == 0 and summary.episodes == stats["episodes"]
    assert stats["env_steps"] == 256 * sim.steps_run     # every env is live on every step (finished ones are reset at once)
    table = sim.learner.feature_weights()
    assert any(abs(w) > 0 for w in table[1].values()) and all(np.isfinite(list(table[1].values())))


@pytest.mark.gpu
def test_experiments_grid_runs_as_one_batch(tmp_path):
    """cavgym_b200.experiments: grid points of (alpha, gamma, epsilon) as independent learners in one batch, testers on the
    device; every grid point gets the reference's config.json / episode.log / run.log."""
    from cavgym_b200 import experiments
    from cavgym_b200.config import AgentType
    grid = [(0.1, 0.9, 0.1), (0.5, 0.5, 0.5), (0.9, 0.1, 0.9), (0.5, 0.9, 0.0)]
    summaries, learner = experiments.run_tester_type(AgentType.RANDOM_CONSTRAINED, grid, runs=3, episodes=3, log_root=str(tmp_path))
    assert set(summaries) == set(grid) and learner.weights.shape[0] == 12
    for point, summary in summaries.items():
        assert summary.episodes == 9 and 9 <= summary.timesteps <= 9 * 1000
        log_dir = tmp_path / f"tester=random-constrained/alpha={point[0]}/gamma={point[1]}/epsilon={point[2]}"
        rows = (log_dir / "episode.log").read_text().strip().splitlines()
        assert len(rows) == 9 and all(len(r.split(",")) == 5 for r in rows)
        assert len((log_dir / "run.log").read_text().strip().split(",")) == 10
        assert json.loads((log_dir / "config.json").read_text())["ego_config"]["gamma"] == point[1]
    weights = learner.weights.cpu().numpy()
    assert np.isfinite(weights).all() and np.abs(weights).max() > 0
    assert not np.allclose(weights[0], weights[3])          # different hyper-parameters learn different tables



@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", ELECTION_CASES)
def test_device_election_agents_follow_the_reference(name, dtype):
    """CAV_AGENT_ELECTION: the ElectionAgents and the arbitration of examples/election.py inside the step kernel (no random
    draws are involved: proximity triggers, closest voter wins).  Consecutive reference episodes in ONE engine, because the
    reference's Election object — and its active player — outlives episodes: executed joint actions, crossing-agent state,
    body state and events must be the reference's."""
    from cavgym_b200 import BatchedCAVEnv
    meta, episodes = load_golden(name)
    env = BatchedCAVEnv(None, None, None, num_envs=1, dtype=dtype, compiled=compile_from_meta(meta, mode="device"), device="cuda:0")
    env.set_action_logging(True)
    tol = 1e-9 if dtype == "float64" else 1e-4
    compared = 0
    for ep in episodes:
        env.reset(init_state=soa(ep["init_state"][None]))
        env.set_global_timestep(int(ep["t_global_start"]))
        for t in range(ep["actions"].shape[0]):
            state, reward, done, winner, tangent = env.step(None)
            if bool(tangent[0]) and (bool(done[0]) != bool(ep["done"][t]) or int(winner[0]) != int(ep["winner"][t])):
                break      # a flagged near-tangent divergence (fp32): different episodes from here on
            assert bool(done[0]) == bool(ep["done"][t]) and int(winner[0]) == int(ep["winner"][t]), (name, t)
            got_actions = env.actions_taken.double().cpu().numpy()[..., 0]
            assert np.max(np.abs(got_actions - ep["actions"][t])) < (1e-9 if dtype == "float64" else 2e-3), (name, t)
            from helpers import state_err_trajectory
            assert state_err(state.double().cpu().numpy()[..., 0], ep["state"][t]) < (tol if dtype == "float64" else 5e-4), (name, t)
            if dtype == "float64":
                crossing = ~np.isnan(ep["agent_state"][t]).all(axis=1)
                got = env.agent_state.cpu().numpy()[..., 0]
                assert np.array_equal(np.isnan(got), np.isnan(ep["agent_state"][t])), (name, t)
                assert np.allclose(got[crossing], ep["agent_state"][t][crossing], rtol=1e-9, atol=1e-9, equal_nan=True), (name, t)
            compared += 1
    assert compared > 500
    env.close()


@pytest.mark.gpu
def test_batched_simulation_with_election_testers():
    """tester_config election at batch scale: three election pedestrians per env on the device, rollout + scoring in-kernel."""
    from cavgym_b200.config import make_config
    from cavgym_b200.simulation import BatchedSimulation
    meta, _ = load_golden("learn_election3_seed22")
    config = make_config(dict(copy.deepcopy(meta["config"]), episodes=2000))
    sim = BatchedSimulation(config, 2048, chunk=250)
    summary = sim.run()
    stats = sim.env.stats()
    assert stats["episodes"] >= 2000 and stats["errors"] == 0 and summary.interesting == stats["interesting"] > 0
