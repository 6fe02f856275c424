"""CPU timing of the UNMODIFIED Python reference (SURVEY 8d "CPU reference timing") — build container only: it imports
/root/reference through oracle/refload.py over the stand-ins for gym / Shapely (neither is installable here), so the figure
is "reference Python + stand-in geometry"; README.md:27 of the reference quotes 2,557 env-steps/s with the real dependencies.

    python scripts/time_python_reference.py [--episodes 10] [--processes 0]

(i) one process: Simulation.run on BASELINE config C1 (stock config.json, ego noop, headless), seed 0;
(ii) multiprocessing.Pool(P) over seeds, one run per process, in the style of experiments.py:95-122.
"""
import argparse
import copy
import multiprocessing
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(args):
    seed, episodes = args
    from oracle import refload
    mods = refload.load()
    cfg = mods["config"].make_config(copy.deepcopy(refload.stock_config_dict(seed=seed, episodes=episodes)))
    _, env, agents, keyboard_agent = cfg.setup()
    simulation = mods["simulation"].Simulation(env, agents, config=cfg, keyboard_agent=keyboard_agent)
    steps = [0]
    step = env.step

    def counted(joint_action):
        steps[0] += 1
        return step(joint_action)

    env.step = counted
    start = time.perf_counter()
    simulation.run()
    return steps[0], time.perf_counter() - start, len(env.bodies)


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--episodes", type=int, default=10)
    parser.add_argument("--processes", type=int, default=0, help="0 = os.cpu_count()")
    opts = parser.parse_args()
    steps, seconds, bodies = run((0, opts.episodes))
    print(f"one process: {steps} env-steps in {seconds:.2f} s = {steps / seconds:,.0f} env-steps/s = {steps * bodies / seconds:,.0f} body-steps/s "
          f"(C1, {opts.episodes} episodes, seed 0)")
    p = opts.processes or os.cpu_count()
    start = time.perf_counter()
    with multiprocessing.Pool(p) as pool:
        results = pool.map(run, [(seed, opts.episodes) for seed in range(p)])
    wall = time.perf_counter() - start
    total = sum(r[0] for r in results)
    print(f"{p} processes (one seed each): {total} env-steps in {wall:.2f} s = {total / wall:,.0f} env-steps/s = {total * bodies / wall:,.0f} body-steps/s")
