"""Single-environment compat view of the batched engine.

`CAVEnv(bodies, constants, env_config, np_random)` keeps the reference's multi-agent Gym
protocol (library/environment.py:15-44 MarkovGameEnv, :54-243 CAVEnv): `reset()` returns the
joint observation as `[[x, y, v, theta], ...]`, `step(joint_action)` returns
`(joint_observation, joint_reward, done, info)`, and `bodies`, `constants`, `env_config`,
`np_random`, `frequency`, `time_resolution`, `action_space`, `observation_space`,
`episode_liveness`, `run_liveness`, `ego`, `current_timestep` behave as callers of the reference
expect (simulation.py, reporting.py:231-238, config.py:295-321).  Underneath, every transition is
one launch of the CUDA engine on a batch of one environment; for throughput use
`cavgym_b200.BatchedCAVEnv` directly.
"""
from dataclasses import dataclass

from .. import spaces
from .assets import RoadMap, Occlusion  # noqa: F401  (re-exported like the reference module)
from .bodies import PelicanCrossing, Pedestrian, DynamicBody, DynamicBodyState, TrafficLightState  # noqa: F401
from .geometry import Point


class MarkovGameEnv:
    """Multi-agent Gym-style environment: all agents act simultaneously each timestep."""
    metadata = {'render.modes': ['human', 'rgb_array']}
    reward_range = (-float('inf'), float('inf'))
    spec = None

    def step(self, joint_action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode='human'):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return

    @property
    def unwrapped(self):
        return self


@dataclass(frozen=True)
class CAVEnvConstants:
    viewer_width: int
    viewer_height: int
    road_map: RoadMap


class _LazyInfo(dict):
    """`info` dict whose 'body_polygons' / 'road_angles' (reference :106-117) are only
    computed if somebody reads them; nothing on the step path consumes them."""

    def __init__(self, env):
        super().__init__()
        self._env = env

    def __missing__(self, key):
        if key not in ('body_polygons', 'road_angles'):
            raise KeyError(key)
        # one launch of the engine's info kernel (cavgym_info) on the batch of one; reference-shaped values out
        from .geometry import ConvexQuadrilateral
        out = self._env._batched().info()
        corners = out['body_polygons'][:, :, 0].double().cpu().tolist()
        angles = out['road_angles'][:, 0].double().cpu().tolist()
        self['body_polygons'] = [ConvexQuadrilateral(*[(row[i], row[4 + i]) for i in range(4)]) for row in corners]
        self['road_angles'] = [None if a != a else a for a in angles]
        return self[key]

    def __contains__(self, key):
        return key in ('body_polygons', 'road_angles') or super().__contains__(key)


class CAVEnv(MarkovGameEnv):
    def __init__(self, bodies, constants, env_config, np_random=None, device=None, dtype="float64"):
        import numpy as np
        self.bodies = bodies
        self.constants = constants
        self.env_config = env_config
        self.np_random = np_random if np_random is not None else np.random.RandomState()
        self.frequency = 60
        self.time_resolution = 1.0 / self.frequency

        self.action_space = spaces.Tuple([body.action_space() for body in bodies])
        self.observation_space = spaces.Tuple([body.observation_space() for body in bodies])
        for space in list(self.action_space) + list(self.observation_space):
            space.np_random = self.np_random

        self.episode_liveness = [0 for _ in bodies]
        self.run_liveness = [0 for _ in bodies]
        self.ego = bodies[0]
        self.ego_maintenance_velocity = self.ego.init_state.velocity
        self.ego_max_velocity_offset = max(abs(self.ego.constants.max_velocity - self.ego_maintenance_velocity),
                                           abs(self.ego.constants.min_velocity - self.ego_maintenance_velocity))
        self.current_timestep = 0
        self.viewer = None
        self._device, self._dtype = device, dtype
        self._engine = None

    # ---- engine plumbing -------------------------------------------------------------
    def _batched(self):
        if self._engine is None:
            from ..engine import BatchedCAVEnv
            self._engine = BatchedCAVEnv(self.bodies, self.constants, self.env_config, num_envs=1,
                                         device=self._device, dtype=self._dtype)
        return self._engine

    def _pull_state(self, rows):
        for body, row in zip(self.bodies, rows):
            if isinstance(body, DynamicBody):
                body.state = DynamicBodyState(Point(row[0], row[1]), row[2], row[3])
            else:
                body.state = TrafficLightState(int(row[0]))
                body.outbound_traffic_light.state = body.inbound_traffic_light.state = body.state

    def collidable_entities(self):
        entities = [body for body in self.bodies if isinstance(body, Occlusion)]
        for body in self.bodies:
            if isinstance(body, PelicanCrossing):
                entities += [body.outbound_traffic_light, body.inbound_traffic_light]
        if self.constants.road_map.obstacle is not None:
            entities.append(self.constants.road_map.obstacle)
        return entities

    # ---- reference protocol ----------------------------------------------------------
    def state(self):
        return [list(body.state) for body in self.bodies]

    def info(self):
        return _LazyInfo(self)

    def reset(self):
        for body in self.bodies:
            body.reset()  # SpawnPedestrian re-draws with the caller's np_random (host RNG, seed-compatible)
        self.episode_liveness = [0 for _ in self.bodies]
        rows = [[float(v.value) if isinstance(v, TrafficLightState) else float(v) for v in body.state] for body in self.bodies]
        self._batched().reset_to(rows)
        return self.state()

    def step(self, joint_action):
        assert self.action_space.contains(joint_action), f"{joint_action} ({type(joint_action)}) invalid"
        engine = self._batched()
        engine.set_global_timestep(self.current_timestep)
        rows, rewards, done, winner, liveness, taken = engine.step_single(joint_action)
        self._pull_state(rows)
        for body, action in zip(self.bodies, taken):
            if isinstance(body, DynamicBody):
                body.throttle, body.steering_angle = action
        for i, count in enumerate(liveness):
            self.run_liveness[i] += count - self.episode_liveness[i]
            self.episode_liveness[i] = count
        info = self.info()
        if winner >= 0:
            info['winner'] = winner
        self.current_timestep += 1
        return self.state(), rewards, done, info

    def render(self, mode='human'):
        raise NotImplementedError("rendering (pyglet) is outside the batched stepping engine; see DESIGN.md 'Out of scope'")

    def close(self):
        self.viewer = None
