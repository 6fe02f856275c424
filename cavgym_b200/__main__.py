"""`python -m cavgym_b200 [config.json] [--envs N]` — the reference's entry point (cavgym.py:1-40: read a config, set the
scenario up, run the simulation).  Without --envs: the reference's own loop on one environment (host-side agents, one
CUDA launch per transition).  With --envs N: N concurrent copies with the on-device agents (BatchedSimulation)."""
import json
import sys

from .config import ConfigParser, make_config
from .simulation import BatchedSimulation, Simulation


def main(argv=None):
    parser = ConfigParser()
    parser.add_argument("--envs", type=int, default=0, help="run this many concurrent environments on the GPU (0: single environment)")
    parser.add_argument("--dtype", default="float64", choices=["float64", "float32"])
    args = parser.parse_args(argv)
    config = make_config(json.load(args.input))
    if args.envs > 0:
        BatchedSimulation(config, args.envs, dtype=args.dtype).run()
    else:
        _, env, agents, keyboard_agent = config.setup()
        Simulation(env, agents, config, keyboard_agent).run()


if __name__ == "__main__":
    sys.exit(main())
