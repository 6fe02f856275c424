"""Scenario compiler: reference-style `bodies` / `constants` / `env_config` objects ->
the flat CavScenario tables of include/cavgym.h.

It walks exactly what CAVEnv.__init__ receives in the reference
(library/environment.py:59-92): the Body list (type, init state, per-type constants),
`constants.road_map` (road rectangles, obstacle), PelicanCrossing traffic lights
(environment.py:94-101), SpawnPedestrian spawn boxes (bodies.py:283-312) and the
env_config fields read inside step (environment.py:136-213).
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

from . import _abi

AGENT_CODES = {"external": _abi.CAV_AGENT_EXTERNAL, "noop": _abi.CAV_AGENT_NOOP, "random": _abi.CAV_AGENT_RANDOM,
               "random-constrained": _abi.CAV_AGENT_RANDOM_CONSTRAINED, "proximity": _abi.CAV_AGENT_PROXIMITY,
               "election": _abi.CAV_AGENT_ELECTION}
COLLISION_CODES = {"none": _abi.CAV_COLLISIONS_NONE, "ego": _abi.CAV_COLLISIONS_EGO, "all": _abi.CAV_COLLISIONS_ALL}


def is_a(obj, *class_names):
    """isinstance by CLASS NAME along the MRO: the compiler reads reference-style objects by the protocol the reference
    defines (library/bodies.py class names and attributes), so bodies built from an unmodified copy of the reference's own
    `library.bodies` compile exactly like the mirrors in cavgym_b200.library.bodies."""
    return any(cls.__name__ in class_names for cls in type(obj).__mro__)


@dataclass
class AgentSpec:
    """Which agent drives a body: 'external' (actions tensor) or an on-device agent."""
    kind: str = "external"
    epsilon: float = 0.0
    threshold: float = 0.0


@dataclass
class CompiledScenario:
    struct: _abi.CavScenario
    n_bodies: int
    body_kinds: List[int]
    body_classes: List[str]
    keepalive: list = field(default_factory=list)

    def pointer(self):
        return C.byref(self.struct)

    def tables(self):
        """The scenario as position-independent bytes: the CavScenario header with its two pointers blanked, the body
        rows and the spawn rows.  Two compilations describe the same scenario iff their tables are equal."""
        header = _abi.CavScenario.from_buffer_copy(self.struct)
        header.bodies = C.cast(None, C.POINTER(_abi.CavBody))
        header.spawns = C.cast(None, C.POINTER(_abi.CavSpawn))
        rows, spawns = self.keepalive
        return {"header": bytes(header), "bodies": bytes(rows), "spawns": bytes(spawns)[:C.sizeof(_abi.CavSpawn) * self.struct.n_spawns]}


def _quad(shape):
    q = _abi.CavQuad()
    for i, (x, y) in enumerate(shape):
        q.x[i], q.y[i] = float(x), float(y)
    return q


def _collision_code(value):
    key = getattr(value, "value", value)
    if key not in COLLISION_CODES:
        raise ValueError(f"terminate_collisions must be one of {sorted(COLLISION_CODES)}, not {value!r}")
    return COLLISION_CODES[key]


def compile_scenario(bodies: Sequence, constants, env_config, agents: Optional[Sequence[AgentSpec]] = None,
                     time_resolution: float = 1.0 / 60) -> CompiledScenario:
    m = len(bodies)
    if not 1 <= m <= _abi.CAV_MAX_BODIES:
        raise ValueError(f"a scenario needs 1..{_abi.CAV_MAX_BODIES} bodies, got {m}")
    agents = list(agents) if agents is not None else [AgentSpec() for _ in bodies]
    if len(agents) != m:
        raise AssertionError("each body must be assigned an agent and vice versa")  # simulation.py:11
    ego = bodies[0]
    if not is_a(ego, "DynamicBody"):
        raise ValueError("the ego (bodies[0]) must be a DynamicBody")

    sc = _abi.CavScenario()
    road_map = constants.road_map
    roads = list(road_map.roads)
    if len(roads) > _abi.CAV_MAX_ROADS:
        raise ValueError(f"at most {_abi.CAV_MAX_ROADS} roads are supported")
    for i, road in enumerate(roads):
        sc.roads[i] = _quad(road.bounding_box())
    sc.n_roads = len(roads)
    centre = road_map.major_road.bounding_box().longitudinal_line()  # config.py:363
    sc.centre_line[:] = [centre.start.x, centre.start.y, centre.end.x, centre.end.y]

    statics = []
    for body in bodies:  # environment.py:94-101
        if is_a(body, "PelicanCrossing"):
            statics += [body.outbound_traffic_light.bounding_box(), body.inbound_traffic_light.bounding_box()]
    if road_map.obstacle is not None:
        statics.append(road_map.obstacle.bounding_box())
    if len(statics) > _abi.CAV_MAX_STATICS:
        raise ValueError(f"at most {_abi.CAV_MAX_STATICS} static collidables are supported")
    for i, box in enumerate(statics):
        sc.statics[i] = _quad(box)
    sc.n_statics = len(statics)

    type_rows, spawn_rows = [], []
    body_rows = (_abi.CavBody * m)()
    kinds, classes = [], []
    for i, (body, agent) in enumerate(zip(bodies, agents)):
        row = body_rows[i]
        classes.append(type(body).__name__)
        if agent.kind not in AGENT_CODES:
            raise NotImplementedError(f"agent kind {agent.kind!r} has no on-device implementation")
        row.agent = AGENT_CODES[agent.kind]
        row.agent_epsilon, row.agent_threshold = float(agent.epsilon), float(agent.threshold)
        row.spawn_id = -1
        if is_a(body, "PelicanCrossing"):
            row.kind = _abi.CAV_BODY_PELICAN
            row.init_state[:] = [float(body.init_state.value), 0.0, 0.0, 0.0]
            row.static_box = _quad(body.bounding_box())
            if agent.kind in ("random-constrained", "proximity", "election"):
                raise NotImplementedError("crossing agents need a Pedestrian body")  # config.py:358-396
        elif is_a(body, "DynamicBody"):
            row.kind = _abi.CAV_BODY_DYNAMIC
            k = body.constants
            key = (k.length, k.width, k.wheelbase, k.min_velocity, k.max_velocity, k.min_throttle, k.max_throttle,
                   k.min_steering_angle, k.max_steering_angle)
            if key not in type_rows:
                type_rows.append(key)
            row.type_id = type_rows.index(key)
            if is_a(body, "Pedestrian"):
                row.flags |= _abi.CAV_FLAG_PEDESTRIAN
            elif agent.kind in ("random-constrained", "proximity", "election"):
                raise NotImplementedError("crossing agents need a Pedestrian body")  # config.py:358-396
            row.init_state[:] = [float(v) for v in body.init_state]
            if is_a(body, "SpawnPedestrian"):
                row.flags |= _abi.CAV_FLAG_SPAWN
                sp = body.spawn_init_state
                if not 1 <= len(sp.position_boxes) <= _abi.CAV_MAX_SPAWN_BOXES or not 1 <= len(sp.orientations) <= _abi.CAV_MAX_SPAWN_ORIENT:
                    raise ValueError("unsupported spawn description")
                spawn = _abi.CavSpawn()
                spawn.n_boxes, spawn.n_orientations = len(sp.position_boxes), len(sp.orientations)
                for j, box in enumerate(sp.position_boxes):
                    spawn.boxes[j] = _quad(box)
                for j, orientation in enumerate(sp.orientations):
                    spawn.orientations[j] = float(orientation)
                spawn.velocity = float(sp.velocity)
                spawn_rows.append(spawn)
                row.spawn_id = len(spawn_rows) - 1
        else:
            raise NotImplementedError(f"body class {type(body).__name__} is not supported by the engine")
        kinds.append(row.kind)
    if len(type_rows) > _abi.CAV_MAX_TYPES:
        raise ValueError(f"at most {_abi.CAV_MAX_TYPES} distinct DynamicBodyConstants are supported")
    for i, key in enumerate(type_rows):
        sc.types[i] = _abi.CavBodyType(*[float(v) for v in key])
    spawns = (_abi.CavSpawn * max(1, len(spawn_rows)))(*spawn_rows)

    sc.n_bodies, sc.n_types, sc.n_spawns = m, len(type_rows), len(spawn_rows)
    sc.bodies = C.cast(body_rows, C.POINTER(_abi.CavBody))
    sc.spawns = C.cast(spawns, C.POINTER(_abi.CavSpawn))
    sc.terminate_collisions = _collision_code(env_config.terminate_collisions)
    sc.terminate_ego_zones = int(bool(env_config.terminate_ego_zones))
    sc.terminate_ego_offroad = int(bool(env_config.terminate_ego_offroad))
    sc.max_timesteps = int(env_config.max_timesteps)
    sc.reward_win, sc.reward_draw, sc.cost_step = float(env_config.reward_win), float(env_config.reward_draw), float(env_config.cost_step)
    sc.viewer_width = float(constants.viewer_width)
    sc.time_resolution = float(time_resolution)
    v0 = ego.init_state.velocity  # environment.py:87-88
    sc.ego_maintenance_velocity = float(v0)
    sc.ego_max_velocity_offset = float(max(abs(ego.constants.max_velocity - v0), abs(ego.constants.min_velocity - v0)))
    return CompiledScenario(sc, m, kinds, classes, keepalive=[body_rows, spawns])
