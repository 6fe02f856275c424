"""`experiments.py` of the reference (experiments.py:11-122) at batch scale: its grid over (tester type, alpha, gamma,
epsilon) of Q-learning-ego runs — one process per grid point there — is ONE batch here: every grid point is an environment
with its own learner (BatchedQLearningEgoAgent(shared=False): own weight table, own alpha / gamma / epsilon), the testers
act on the device, and the per-run results come from the device's per-episode rows.

    python -m cavgym_b200.experiments [--runs R] [--episodes E] [--log-dir logs]

One engine per tester type (the tester is part of the compiled scenario); `--runs R` repeats every grid point R times with
different environments (the reference runs each point once, with seed 0).  For every grid point the reference's files are
written: `logs/tester=<t>/alpha=<a>/gamma=<g>/epsilon=<e>/{config.json, episode.log, run.log}` with its row formats
(reporting.py:157-158, 223-224); with R > 1 the logs hold the episodes of all repetitions.

The reference's script is stale at HEAD (it passes a float where QLearningConfig wants a LinSpace); the
learning rate of a grid point is taken as constant, which is what a float alpha meant.
"""
import argparse
import itertools
import os

from . import reporting
from .config import (AgentType, CollisionType, Config, FeatureConfig, HeadlessConfig, LinSpace, PedestriansConfig, ProximityConfig,
                     QLearningConfig, RandomConfig, RandomConstrainedConfig)
from .examples.constants import M2PX

TESTER_TYPES = (AgentType.RANDOM, AgentType.RANDOM_CONSTRAINED, AgentType.PROXIMITY)   # experiments.py:112
ALPHAS = GAMMAS = EPSILONS = (0.1, 0.5, 0.9)                                           # experiments.py:113-115


def make_tester_config(agent_type):
    """experiments.py:11-37 (the on-device testers)."""
    if agent_type is AgentType.RANDOM:
        return RandomConfig(epsilon=0.01)
    if agent_type is AgentType.RANDOM_CONSTRAINED:
        return RandomConstrainedConfig(epsilon=0.5)
    if agent_type is AgentType.PROXIMITY:
        return ProximityConfig(threshold=float(M2PX * 34))
    raise NotImplementedError(agent_type)


def make_config(tester_type, alpha, gamma, epsilon, log_root="logs", episodes=10):
    """experiments.py:40-82: the config of one grid point and its log directory."""
    log_dir = f"{log_root}/tester={tester_type}/alpha={alpha}/gamma={gamma}/epsilon={epsilon}"
    features = FeatureConfig(distance_x=False, distance_y=False, distance=True, relative_angle=True, heading=True, on_road=False,
                             inverse_distance=False)
    return log_dir, Config(
        verbosity=reporting.Verbosity.SILENT, episode_log=f"{log_dir}/episode.log", run_log=f"{log_dir}/run.log", seed=0,
        episodes=episodes, max_timesteps=1000, terminate_collisions=CollisionType.EGO, terminate_ego_zones=True,
        terminate_ego_offroad=False, reward_win=6000.0, reward_draw=2000.0, cost_step=4.0,
        scenario_config=PedestriansConfig(num_pedestrians=1, outbound_pavement=1.0, inbound_pavement=1.0),
        ego_config=QLearningConfig(alpha=LinSpace(start=alpha, stop=alpha, num_steps=2), gamma=gamma, epsilon=epsilon, features=features,
                                   log=None),
        tester_config=make_tester_config(tester_type), mode_config=HeadlessConfig())


def run_tester_type(tester_type, grid, runs=1, episodes=10, log_root="logs", device=None, dtype="float64", max_steps=200000):
    """All grid points of one tester type in one batch.  Returns {(alpha, gamma, epsilon): RunSummary}."""
    import timeit
    import torch
    from .examples.agents.ego import BatchedQLearningEgoAgent
    grid = list(grid)
    points = [point for point in grid for _ in range(runs)]          # env e is a run of points[e]
    n = len(points)
    configs = {point: make_config(tester_type, *point, log_root=log_root, episodes=episodes) for point in grid}
    template = configs[grid[0]][1]
    env = template.batched(n, device=device, dtype=dtype)
    learner = BatchedQLearningEgoAgent(template.ego_config, env.bodies[0].constants, env.time_resolution, env.num_bodies - 1,
                                       env.constants.viewer_width, env.constants.viewer_height, n, env.device, dtype=env.dtype,
                                       shared=False, alpha=([p[0] for p in points], [p[0] for p in points], 2.0),
                                       gamma=[p[1] for p in points], epsilon=[p[2] for p in points])
    env.set_episode_log(4 * n + 1024)
    env.reset()
    joint = torch.zeros((env.num_bodies, 2, n), dtype=env.dtype, device=env.device)
    previous = torch.empty_like(env.state)
    rows_of = [[] for _ in range(n)]
    start, steps = timeit.default_timer(), 0
    while min(len(rows) for rows in rows_of) < episodes and steps < max_steps:
        for _ in range(100):      # Simulation.run's timestep loop (simulation.py:69-93) for every run at once
            previous.copy_(env.state)
            index, ego_rows = learner.choose_action(previous)
            joint[0] = ego_rows
            state, reward, _, _, _ = env.step(joint)
            learner.process_feedback(previous, index, state, reward[0])
            env.reset(mask=env.done_latch != 0)
        steps += 100
        drained, dropped = env.drain_episodes()
        assert dropped == 0, "episode ring too small"
        for row in drained:
            rows_of[int(row["env"])].append(row)
    runtime_ms = (timeit.default_timer() - start) * 1000
    nan = float("nan")
    out = {}
    for point in grid:
        log_dir, config = configs[point]
        os.makedirs(log_dir, exist_ok=True)
        config.write_json(f"{log_dir}/config.json")
        episode_file, run_file = reporting.get_episode_file_logger(config.episode_log), reporting.get_run_file_logger(config.run_log)
        results = []
        for e, p in enumerate(points):
            if p != point:
                continue
            for row in rows_of[e][:episodes]:      # the first `episodes` episodes of the run, as the reference stops there
                interesting = int(row["winner"]) > 0
                results.append(reporting.EpisodeResults(len(results) + 1, reporting.TimeResults(int(row["timesteps"]), nan, nan, env.time_resolution),
                                                        completed=int(row["timesteps"]) == config.max_timesteps, interesting=interesting,
                                                        score=-int(row["liveness_sum"]) if interesting else nan))
        for result in results:
            episode_file.info(result.file_message())
        hits = [r for r in results if r.interesting]
        summary = reporting.RunSummary(len(results), sum(r.time.timesteps for r in results), runtime_ms, env.time_resolution, len(hits),
                                       reporting.confidence_interval([r.time.timesteps for r in hits]),
                                       reporting.confidence_interval([r.score for r in hits]))
        run_file.info(summary.file_message())
        out[point] = summary
    env.close()
    return out, learner


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--runs", type=int, default=1, help="repetitions of every grid point (the reference: 1)")
    parser.add_argument("--episodes", type=int, default=10)
    parser.add_argument("--log-dir", default="logs")
    parser.add_argument("-p", "--processes", type=int, default=None, help="accepted for compatibility; the batch replaces the process pool")
    args = parser.parse_args(argv)
    grid = list(itertools.product(ALPHAS, GAMMAS, EPSILONS))
    for tester_type in TESTER_TYPES:
        print(f"starting: tester={tester_type}, {len(grid)} grid point(s) x {args.runs} run(s)")
        summaries, _ = run_tester_type(tester_type, grid, runs=args.runs, episodes=args.episodes, log_root=args.log_dir)
        hits = sum(s.interesting for s in summaries.values())
        print(f"finished: tester={tester_type}: {sum(s.episodes for s in summaries.values())} episode(s), {hits} interesting")


if __name__ == "__main__":
    main()
