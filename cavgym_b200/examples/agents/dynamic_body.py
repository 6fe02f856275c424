"""Steering/throttle controllers for dynamic bodies (reference examples/agents/dynamic_body.py:11-124).
`make_steering_action` is the host statement of the inverse bicycle arc that csrc/agents.cuh
(`steering_towards`) evaluates on the device for the crossing agents."""
import math

from ...library.bodies import DynamicBodyState
from ...library.geometry import Point
from ..targets import TargetOrientation, TargetVelocity
from .template import NoopAgent

TARGET_ERROR = 0.000000000000001
ACTION_ERROR = 0.000000000000001


def make_body_state(env_state, index):
    x, y, velocity, orientation = env_state[index]
    return DynamicBodyState(position=Point(float(x), float(y)), velocity=float(velocity), orientation=float(orientation))


def _clamp(value, low, high):
    return min(high, max(low, value))


def make_throttle_action(body_state, body_constants, time_resolution, target_velocity, noop_action):
    throttle = noop_action[0] if target_velocity is None else (target_velocity - body_state.velocity) / time_resolution
    return _clamp(throttle, body_constants.min_throttle, body_constants.max_throttle)


def _wrapped_difference(target, orientation):
    return math.atan2(math.sin(target - orientation), math.cos(target - orientation))


def make_steering_action(body_state, body_constants, time_resolution, target_orientation, noop_action):
    """Steering angle that turns the body towards target_orientation in one step, saturating at full lock."""
    k, v = body_constants, body_state.velocity
    if v == 0 or target_orientation is None:
        return _clamp(noop_action[1], k.min_steering_angle, k.max_steering_angle)
    wanted = _wrapped_difference(target_orientation, body_state.orientation)
    lock = k.min_steering_angle if wanted < 0 else k.max_steering_angle
    reachable = (-1 if lock < 0 else 1) * 2 * time_resolution * v / math.sqrt(k.wheelbase**2 * (1 + 4 / math.tan(lock)**2))
    turn = reachable if wanted / reachable > 1 else wanted
    steering = (-1 if turn < 0 else 1) * math.atan(
        2 * k.wheelbase * math.sqrt(turn**2 / (4 * v**2 * time_resolution**2 - k.wheelbase**2 * turn**2)))
    return _clamp(steering, k.min_steering_angle, k.max_steering_angle)


class TargetAgent(NoopAgent):
    def __init__(self, body, time_resolution, **kwargs):
        super().__init__(noop_action=body.noop_action, **kwargs)
        self.body = body
        self.time_resolution = time_resolution
        k = body.constants
        self.target_velocity_mapping = {TargetVelocity.MIN: k.min_velocity, TargetVelocity.MID: (k.min_velocity + k.max_velocity) / 2,
                                        TargetVelocity.MAX: k.max_velocity}

    def reset(self):
        self.body.target_velocity = None
        self.body.target_orientation = None

    def choose_action(self, state, action_space, info=None):
        body_state, k = make_body_state(state, self.index), self.body.constants
        throttle = make_throttle_action(body_state, k, self.time_resolution, self.body.target_velocity, self.noop_action)
        steering = make_steering_action(body_state, k, self.time_resolution, self.body.target_orientation, self.noop_action)
        return [_clamp(throttle, action_space.low[0], action_space.high[0]), _clamp(steering, action_space.low[1], action_space.high[1])]

    def process_feedback(self, previous_state, action, state, reward):
        body_state = make_body_state(state, self.index)
        if self.body.target_velocity is not None and abs(self.body.target_velocity - body_state.velocity) < TARGET_ERROR:
            self.body.target_velocity = None
        if self.body.target_orientation is not None and \
                abs(_wrapped_difference(self.body.target_orientation, body_state.orientation)) < TARGET_ERROR:
            self.body.target_orientation = None

    def device_spec(self):
        return None  # interactive / host-side only


_VELOCITY_KEYS = {65365: TargetVelocity.MAX, 65366: TargetVelocity.MID, 65367: TargetVelocity.MIN}
_ORIENTATION_KEYS = {65361: TargetOrientation.WEST, 65362: TargetOrientation.NORTH, 65363: TargetOrientation.EAST,
                     65364: TargetOrientation.SOUTH, 65457: TargetOrientation.SOUTH_WEST, 65458: TargetOrientation.SOUTH,
                     65459: TargetOrientation.SOUTH_EAST, 65460: TargetOrientation.WEST, 65462: TargetOrientation.EAST,
                     65463: TargetOrientation.NORTH_WEST, 65464: TargetOrientation.NORTH, 65465: TargetOrientation.NORTH_EAST}
key_target_velocity, key_target_orientation = _VELOCITY_KEYS, _ORIENTATION_KEYS


class KeyboardAgent(TargetAgent):
    def key_press(self, key, _mod):
        if key in _VELOCITY_KEYS:
            self.body.target_velocity = self.target_velocity_mapping[_VELOCITY_KEYS[key]]
        elif key in _ORIENTATION_KEYS:
            self.body.target_orientation = _ORIENTATION_KEYS[key].value
