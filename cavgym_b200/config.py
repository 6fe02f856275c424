"""config.json scenarios (reference config.py:23-531, config.json, schema.json): the same keys, options, value checks and
error behaviour, so a reference config file drives this engine unchanged.

  make_config(data) / ConfigParser      JSON -> Config                         (reference config.py:417-531)
  Config.setup()                        seed, single-env compat view + host-side agents   (reference config.py:272-415)
  Config.agent_specs(bodies)            which ON-DEVICE agent drives each body (same body/agent compatibility rules)
  Config.batched(num_envs, ...)         BatchedCAVEnv over N copies of the scenario with the on-device agents
  Config.write_json(path)               Config -> JSON, the inverse of make_config    (reference config.py:417-437)

Sections are described by one table (OPTIONS) instead of a class per option; the reference's class names
(PedestriansConfig, RandomConstrainedConfig, HeadlessConfig ...) are kept as constructors because callers in the style
of experiments.py:95-119 build configs with them.

The Q-learning ego (examples/agents/ego.py) and the election tester (examples/agents/pedestrian.py, examples/election.py)
run host-side on the single-environment view like in the reference; at batch scale the Q-learning ego has a tensor-API
form (BatchedQLearningEgoAgent, driven by BatchedSimulation) and the election testers are an on-device agent kind
(CAV_AGENT_ELECTION: the arbitration of examples/election.py runs inside the step kernel).  Not provided: the keyboard ego (interactive) and the
Q-learning TESTER (the reference's own raises TypeError at its first update, pedestrian.py:128, 247).  Render mode has no
viewer here (pyglet is not part of the engine): a render config runs headless and says so once.  Every option still
parses, validates and round-trips; building an unavailable one raises NotImplementedError, as the reference does for
combinations it does not support (config.py:343, 396).
"""
import json
import pathlib
from argparse import ArgumentParser, FileType
from dataclasses import asdict, dataclass, field, make_dataclass
from enum import Enum
from typing import Optional

from .reporting import Verbosity, get_console, pretty_str_list


class _StrEnum(Enum):
    def __str__(self):
        return self.value


class Scenario(_StrEnum):
    BUS_STOP = "bus-stop"
    CROSSROADS = "crossroads"
    PEDESTRIANS = "pedestrians"
    PELICAN_CROSSING = "pelican-crossing"


class AgentType(_StrEnum):
    NOOP = "noop"
    KEYBOARD = "keyboard"
    RANDOM = "random"
    RANDOM_CONSTRAINED = "random-constrained"
    PROXIMITY = "proximity"
    ELECTION = "election"
    Q_LEARNING = "q-learning"


class Mode(_StrEnum):
    HEADLESS = "headless"
    RENDER = "render"


class CollisionType(Enum):
    NONE = "none"
    EGO = "ego"
    ALL = "all"


# ---------------------------------------------------------------- value checks (messages as the reference words them)
def _unit_interval(name):
    def check(value):
        if value < 0 or value > 1:
            raise ValueError(f"{name} must be in [0, 1]" if name not in ("outbound_pavement", "inbound_pavement") else f"{name} must be in [0,1]")
    return check


def _positive(name):
    def check(value):
        if value <= 0:
            raise ValueError(f"{name} must be >= 0")   # sic: the reference's wording for `threshold <= 0`
    return check


def _at_least(name, bound):
    def check(value):
        if value < bound:
            raise ValueError(f"{name} must be >= {bound}")
    return check


def _section(name, tag_field, tag, fields, extra_check=None):
    """A frozen dataclass with typed fields, per-field checks and a class attribute naming its option."""
    def post_init(self):
        for field_name, field_type, check in fields:
            value = getattr(self, field_name)
            optional = getattr(field_type, "__origin__", None) is not None   # Optional[...]
            expected = field_type.__args__[0] if optional else field_type
            if value is None and optional:
                continue
            if expected is float and isinstance(value, int) and not isinstance(value, bool):
                object.__setattr__(self, field_name, float(value))
            elif not isinstance(value, expected) or (expected is int and isinstance(value, bool)):
                raise TypeError(f"{name}.{field_name} must be {expected.__name__}, not {type(value).__name__}")
            if check is not None:
                check(getattr(self, field_name))
        if extra_check is not None:
            extra_check(self)
    cls = make_dataclass(name, [(f, t) for f, t, _ in fields], frozen=True, namespace={"__post_init__": post_init, tag_field: tag})
    return cls


BusStopConfig = _section("BusStopConfig", "scenario", Scenario.BUS_STOP, [])
CrossroadsConfig = _section("CrossroadsConfig", "scenario", Scenario.CROSSROADS, [])
PelicanCrossingConfig = _section("PelicanCrossingConfig", "scenario", Scenario.PELICAN_CROSSING, [])
PedestriansConfig = _section("PedestriansConfig", "scenario", Scenario.PEDESTRIANS, [
    ("num_pedestrians", int, _at_least("num_pedestrians", 0)),
    ("outbound_pavement", float, _unit_interval("outbound_pavement")),
    ("inbound_pavement", float, _unit_interval("inbound_pavement"))])

NoopConfig = _section("NoopConfig", "agent", AgentType.NOOP, [])
KeyboardConfig = _section("KeyboardConfig", "agent", AgentType.KEYBOARD, [])
RandomConfig = _section("RandomConfig", "agent", AgentType.RANDOM, [("epsilon", float, _unit_interval("epsilon"))])
RandomConstrainedConfig = _section("RandomConstrainedConfig", "agent", AgentType.RANDOM_CONSTRAINED,
                                   [("epsilon", float, _unit_interval("epsilon"))])
ProximityConfig = _section("ProximityConfig", "agent", AgentType.PROXIMITY, [("threshold", float, _positive("threshold"))])
ElectionConfig = _section("ElectionConfig", "agent", AgentType.ELECTION, [("threshold", float, _positive("threshold"))])

FeatureConfig = make_dataclass("FeatureConfig", [(name, bool) for name in (
    "distance_x", "distance_y", "distance", "relative_angle", "heading", "on_road", "inverse_distance")], frozen=True)


def _check_linspace(self):
    if self.stop > self.start:
        raise ValueError("start must be greater than stop")


LinSpace = _section("LinSpace", "kind", "linspace", [("start", float, _unit_interval("start")), ("stop", float, _unit_interval("stop")),
                                                     ("num_steps", int, _at_least("num_steps", 2))], _check_linspace)
QLearningConfig = _section("QLearningConfig", "agent", AgentType.Q_LEARNING, [
    ("alpha", LinSpace, None), ("gamma", float, _unit_interval("gamma")), ("epsilon", float, _unit_interval("epsilon")),
    ("features", FeatureConfig, None), ("log", Optional[str], None)])

HeadlessConfig = _section("HeadlessConfig", "mode", Mode.HEADLESS, [])
RenderConfig = _section("RenderConfig", "mode", Mode.RENDER, [("episode_condition", int, _at_least("episode_condition", 1)),
                                                               ("video_dir", Optional[str], None)])

# option string -> constructor, per config.json section
OPTIONS = {
    "scenario_config": {"bus-stop": BusStopConfig, "crossroads": CrossroadsConfig, "pedestrians": PedestriansConfig,
                        "pelican-crossing": PelicanCrossingConfig},
    "ego_config": {"noop": NoopConfig, "keyboard": KeyboardConfig, "random": RandomConfig, "q-learning": QLearningConfig},
    "tester_config": {"noop": NoopConfig, "random": RandomConfig, "random-constrained": RandomConstrainedConfig,
                      "proximity": ProximityConfig, "election": ElectionConfig, "q-learning": QLearningConfig},
    "mode_config": {"headless": HeadlessConfig, "render": RenderConfig},
}
ENV_IDS = {Scenario.PELICAN_CROSSING: "PelicanCrossing-v0", Scenario.BUS_STOP: "BusStop-v0", Scenario.CROSSROADS: "Crossroads-v0",
           Scenario.PEDESTRIANS: "Pedestrians-v0"}
_HOST_ONLY = {AgentType.KEYBOARD: "the keyboard agent is interactive (render mode)",
              AgentType.Q_LEARNING: "Q-learning agents update their weights on the host (ego: BatchedQLearningEgoAgent on the tensor API)"}
_UNAVAILABLE = {"ego": {AgentType.KEYBOARD: _HOST_ONLY[AgentType.KEYBOARD]},
                "tester": {AgentType.Q_LEARNING: "the reference's Q-learning tester fails at its first process_feedback "
                                                 "(pedestrian.py:128, 247: LinSpace * float), there is no behaviour to reproduce"}}


@dataclass(frozen=True)
class Config:
    verbosity: Verbosity
    episode_log: Optional[str]
    run_log: Optional[str]
    seed: Optional[int]
    episodes: int
    max_timesteps: int
    terminate_collisions: CollisionType
    terminate_ego_zones: bool
    terminate_ego_offroad: bool
    reward_win: float
    reward_draw: float
    cost_step: float
    scenario_config: object
    ego_config: object
    tester_config: object
    mode_config: object

    def __post_init__(self):
        if self.episodes <= 0:
            raise ValueError("seed must be > 0")       # sic (reference config.py:265)
        if self.seed and self.seed < 0:
            raise ValueError("seed must be >= 0")
        if self.max_timesteps <= 0:
            raise ValueError("seed must be > 0")       # sic (reference config.py:269)
        for name, allowed in (("scenario_config", OPTIONS["scenario_config"]), ("ego_config", OPTIONS["ego_config"]),
                              ("tester_config", OPTIONS["tester_config"]), ("mode_config", OPTIONS["mode_config"])):
            if type(getattr(self, name)) not in allowed.values():
                raise TypeError(f"{name} is not one of {sorted(allowed)}")

    # ---- scenario ----------------------------------------------------------------------------
    def env_kwargs(self):
        sc = self.scenario_config
        if sc.scenario is Scenario.PEDESTRIANS:
            return {"num_pedestrians": sc.num_pedestrians, "outbound_percentage": sc.outbound_pavement,
                    "inbound_percentage": sc.inbound_pavement}
        return {}

    def make_env(self, np_random):
        """The compat single-environment view, as gym.make(id, env_config=self, np_random=np_random, ...) builds it."""
        from . import examples
        return examples.make(ENV_IDS[self.scenario_config.scenario], env_config=self, np_random=np_random, **self.env_kwargs())

    # ---- agents -------------------------------------------------------------------------------
    def _unsupported(self, agent_type):
        raise NotImplementedError(f"agent option {agent_type.value!r} has no on-device form in cavgym_b200: {_HOST_ONLY[agent_type]}")

    def agent_specs(self, bodies):
        """AgentSpec (on-device agent) per body, under the reference's compatibility rules (config.py:343-410):
        crossing agents only drive Pedestrian bodies, a PelicanCrossing only takes noop / random."""
        from .library import bodies as body_lib
        from .scenario import AgentSpec
        ego, tester = self.ego_config, self.tester_config
        if ego.agent is AgentType.Q_LEARNING:
            specs = [AgentSpec("external")]   # the learner chooses the ego's action on the tensor API, every step
        elif ego.agent in _HOST_ONLY:
            self._unsupported(ego.agent)
        else:
            specs = [AgentSpec("noop") if ego.agent is AgentType.NOOP else AgentSpec("random", epsilon=ego.epsilon)]
        for body in bodies[1:]:
            if tester.agent in _HOST_ONLY:
                self._unsupported(tester.agent)
            if tester.agent is AgentType.NOOP:
                specs.append(AgentSpec("noop"))
            elif tester.agent is AgentType.RANDOM:
                specs.append(AgentSpec("random", epsilon=tester.epsilon))
            elif isinstance(body, body_lib.Pedestrian) and tester.agent is AgentType.RANDOM_CONSTRAINED:
                specs.append(AgentSpec("random-constrained", epsilon=tester.epsilon))
            elif isinstance(body, body_lib.Pedestrian) and tester.agent is AgentType.PROXIMITY:
                specs.append(AgentSpec("proximity", threshold=tester.threshold))
            elif isinstance(body, body_lib.Pedestrian) and tester.agent is AgentType.ELECTION:
                specs.append(AgentSpec("election", threshold=tester.threshold))   # arbitrated inside the step kernel
            else:
                raise NotImplementedError   # config.py:396: e.g. random-constrained on a Car
        return specs

    def setup(self):
        """(np_seed, env, agents, keyboard_agent) like the reference: one RandomState seeds the env, the spawners and
        every tester agent, in agent-index order (config.py:275-415)."""
        from . import seeding
        from .examples.agents.ego import QLearningEgoAgent
        from .examples.agents.pedestrian import ElectionAgent, ProximityAgent, RandomConstrainedAgent
        from .examples.agents.template import NoopAgent, RandomAgent
        from .library import bodies as body_lib
        from .library.actions import TrafficLightAction
        console = get_console(self.verbosity)
        np_random, np_seed = seeding.np_random(self.seed)
        console.info(f"seed={np_seed}")
        if self.mode_config.mode is Mode.RENDER:
            console.warning("render mode: cavgym_b200 has no viewer (pyglet is not part of the engine), running headless")
        env = self.make_env(np_random)
        console.info(f"bodies={pretty_str_list(body.__class__.__name__ for body in env.bodies)}")
        ego, tester = self.ego_config, self.tester_config
        if ego.agent in _UNAVAILABLE["ego"]:
            self._unsupported(ego.agent)
        if ego.agent is AgentType.NOOP:
            agents = [NoopAgent(index=0, noop_action=env.bodies[0].noop_action)]
        elif ego.agent is AgentType.Q_LEARNING:   # config.py:311-322: shares the env's RandomState, five throttle actions
            agents = [QLearningEgoAgent(index=0, np_random=np_random, q_learning_config=ego, body=env.bodies[0],
                                        time_resolution=env.time_resolution, width=env.constants.viewer_width,
                                        height=env.constants.viewer_height, num_actions=5, num_opponents=len(env.bodies) - 1)]
        else:   # the reference builds the ego's RandomAgent WITHOUT the shared np_random (config.py:305-310)
            agents = [RandomAgent(index=0, noop_action=env.bodies[0].noop_action, epsilon=ego.epsilon)]
        road_centre = env.constants.road_map.major_road.bounding_box().longitudinal_line()
        for i, body in enumerate(env.bodies[1:], start=1):
            if tester.agent in _UNAVAILABLE["tester"]:
                raise NotImplementedError(f"tester option {tester.agent.value!r}: {_UNAVAILABLE['tester'][tester.agent]}")
            if isinstance(body, body_lib.DynamicBody):
                pedestrian = isinstance(body, body_lib.Pedestrian)
                if tester.agent is AgentType.NOOP:
                    agent = NoopAgent(index=i, noop_action=body.noop_action)
                elif tester.agent is AgentType.RANDOM:
                    agent = RandomAgent(index=i, noop_action=body.noop_action, epsilon=tester.epsilon, np_random=np_random)
                elif tester.agent is AgentType.RANDOM_CONSTRAINED and pedestrian:
                    agent = RandomConstrainedAgent(index=i, body=body, time_resolution=env.time_resolution, road_centre=road_centre,
                                                   epsilon=tester.epsilon, np_random=np_random)
                elif tester.agent is AgentType.PROXIMITY and pedestrian:
                    agent = ProximityAgent(index=i, body=body, time_resolution=env.time_resolution, road_centre=road_centre,
                                           distance_threshold=tester.threshold)
                elif tester.agent is AgentType.ELECTION and pedestrian:
                    agent = ElectionAgent(index=i, body=body, time_resolution=env.time_resolution, road_centre=road_centre,
                                          distance_threshold=tester.threshold)
                else:
                    raise NotImplementedError
            else:   # PelicanCrossing: with any other tester the reference re-appends the previous loop's agent (:397-410)
                if tester.agent is AgentType.NOOP:
                    agent = NoopAgent(index=i, noop_action=TrafficLightAction.NOOP.value)
                elif tester.agent is AgentType.RANDOM:
                    agent = RandomAgent(index=i, noop_action=TrafficLightAction.NOOP.value, epsilon=tester.epsilon, np_random=np_random)
                else:
                    agent = agents[-1]
            agents.append(agent)
        console.info(f"agents={pretty_str_list(agent.__class__.__name__ for agent in agents)}")
        console.info(f"ego=({env.bodies[0].__class__.__name__}, {agents[0].__class__.__name__})")
        return np_seed, env, agents, None

    def batched(self, num_envs, device=None, dtype="float64", seed=None, env_offset=0):
        """N copies of this config's scenario on the GPU with the on-device agents; `seed` defaults to the config's."""
        import numpy as np
        from .engine import BatchedCAVEnv
        template = self.make_env(np.random.RandomState(0))   # scenario description only: spawns are re-drawn on the device
        seed = (self.seed or 0) if seed is None else seed
        return BatchedCAVEnv(template.bodies, template.constants, self, num_envs=num_envs, agents=self.agent_specs(template.bodies),
                             device=device, dtype=dtype, seed=seed, env_offset=env_offset)

    # ---- JSON ----------------------------------------------------------------------------------
    def to_data(self):
        def section(value, tag):
            return {"option": str(getattr(value, tag)), **asdict(value)}
        data = asdict(self)
        data["verbosity"] = self.verbosity.value
        data["terminate_collisions"] = self.terminate_collisions.value
        data["scenario_config"] = section(self.scenario_config, "scenario")
        data["ego_config"] = section(self.ego_config, "agent")
        data["tester_config"] = section(self.tester_config, "agent")
        data["mode_config"] = section(self.mode_config, "mode")
        for key in ("ego_config", "tester_config"):      # config.json spells the Q-learning feature block feature_config
            if "features" in data[key]:
                data[key]["feature_config"] = data[key].pop("features")
        return data

    def write_json(self, path):
        pathlib.Path(path).parent.mkdir(parents=True, exist_ok=True)
        with open(path, "w", encoding="utf-8") as handle:
            json.dump(self.to_data(), handle, ensure_ascii=False, indent=2)


def _make_section(section, data):
    data = dict(data)
    option = data.pop("option")
    if option not in OPTIONS[section]:
        raise NotImplementedError(f"{section} option {option!r}")
    cls = OPTIONS[section][option]
    if cls is QLearningConfig:
        data["alpha"] = LinSpace(**data.pop("alpha"))
        data["features"] = FeatureConfig(**data.pop("feature_config"))
    fields = getattr(cls, "__dataclass_fields__", {})
    unexpected = {k: v for k, v in data.items() if k not in fields}
    if unexpected:
        raise ValueError(f"unexpected parameters {unexpected}")
    return cls(**data)


def make_config(data):
    """dict (parsed config.json) -> Config."""
    data = dict(data)
    data["verbosity"] = Verbosity(str(data["verbosity"]))
    data["terminate_collisions"] = CollisionType(str(data["terminate_collisions"]))
    for section in OPTIONS:
        data[section] = _make_section(section, data[section])
    for key in ("reward_win", "reward_draw", "cost_step"):
        data[key] = float(data[key])
    return Config(**data)


class ConfigParser(ArgumentParser):
    """`cavgym.py [INPUT]`: read a config from a file, or from stdin when none is given (reference config.py:520-531)."""

    def __init__(self):
        super().__init__()
        self.add_argument("input", metavar="INPUT", nargs="?", type=FileType("r"), default="-",
                          help="read config from %(metavar)s file, or from stdin if no file is provided")

    def parse_config(self, argv=None):
        return make_config(json.load(self.parse_args(argv).input))
