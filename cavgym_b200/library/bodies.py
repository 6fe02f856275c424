"""Body plugin surface (reference library/bodies.py:23-461): the Body subclasses a
scenario is written with.  In this engine a Body is a DESCRIPTOR — init_state,
constants and class (Car / Pedestrian / SpawnPedestrian / ...) are compiled by
cavgym_b200/scenario.py into the device tables; the per-step integration
(reference DynamicBody.step :214-275) runs in the CUDA kernels (csrc/agents.cuh: body_step,
shared by every kernel through csrc/transition.cuh) for all environments at once.  `body.state` is refreshed from the device by the
single-environment compat view (library/environment.py) after every step.
"""
import math
from copy import copy
from dataclasses import dataclass
from enum import Enum

import numpy as np

from .. import spaces
from . import geometry
from .actions import TrafficLightAction
from .assets import Occlusion, Road
from .geometry import Point

REACTION_TIME = 0.675
STEERING_ERROR = 0.0000000000001
DISTANCE_ERROR = 0.001


class Body:
    def __init__(self, init_state, constants, **kwargs):
        super().__init__(**kwargs)
        self.init_state = init_state
        self.constants = constants
        self.state = copy(init_state)

    def reset(self):
        self.state = copy(self.init_state)

    def bounding_box(self):
        raise NotImplementedError

    def step(self, action, time_resolution):
        raise NotImplementedError

    def observation_space(self):
        raise NotImplementedError

    def action_space(self):
        raise NotImplementedError


@dataclass
class DynamicBodyState:
    position: Point
    velocity: float
    orientation: float

    def __copy__(self):
        return DynamicBodyState(copy(self.position), self.velocity, self.orientation)

    def __iter__(self):
        yield from self.position
        yield self.velocity
        yield self.orientation


@dataclass(frozen=True)
class DynamicBodyConstants:
    length: float
    width: float
    wheelbase: float
    track: float
    min_velocity: float
    max_velocity: float
    min_throttle: float
    max_throttle: float
    min_steering_angle: float
    max_steering_angle: float


class DynamicBody(Body, Occlusion):
    def __init__(self, init_state, constants):
        super().__init__(init_state=init_state, constants=constants)
        self.shape = geometry.make_rectangle(constants.length, constants.width)
        self.wheels = geometry.make_rectangle(constants.wheelbase, constants.track)
        self.wheelbase_offset = constants.wheelbase / 2.0
        self.noop_action = [0.0, 0.0]
        self.throttle, self.steering_angle = self.noop_action
        self.target_velocity = None
        self.target_orientation = None
        self.target_spline = None
        self.planner_spline = None

    def observation_space(self):
        k = self.constants
        return spaces.Box(low=np.array([-math.inf, -math.inf, k.min_velocity, -math.pi], dtype=np.float64),
                          high=np.array([math.inf, math.inf, k.max_velocity, math.pi], dtype=np.float64), dtype=np.float64)

    def action_space(self):
        k = self.constants
        return spaces.Box(low=np.array([k.min_throttle, k.min_steering_angle], dtype=np.float64),
                          high=np.array([k.max_throttle, k.max_steering_angle], dtype=np.float64), dtype=np.float64)

    def reset(self):
        super().reset()
        self.throttle, self.steering_angle = self.noop_action

    def bounding_box(self):
        return self.shape.transform(self.state.orientation, self.state.position)

    def wheel_positions(self):
        return self.wheels.transform(self.state.orientation, self.state.position)

    def stopping_zones(self):
        """Braking and reaction zones ahead of the body (reference :122-135); the device
        evaluates the same construction every step (csrc/geometry.cuh: EgoFrame, ego_margins)."""
        braking = (self.state.velocity ** 2) / (2 * -self.constants.min_throttle)
        total = braking + self.state.velocity * REACTION_TIME
        if total == 0 or self.steering_angle != 0:
            return None, None
        front = Point(self.constants.length * 0.5, 0).transform(self.state.orientation, self.state.position)
        zone = geometry.make_rectangle(total, self.constants.width, rear_offset=0).transform(self.state.orientation, front)
        return zone.split_longitudinally(braking / total)

    def line_anchor(self, road):
        closest = road.bounding_box().longitudinal_line().closest_point_from(self.state.position)
        return geometry.Line(self.state.position, closest)

    def line_anchor_relative_angle(self, road):
        return geometry.normalise_angle(self.line_anchor(road).orientation() - self.state.orientation)

    def step(self, action, time_resolution):
        """One kinematic-bicycle step of THIS body alone, executed by the CUDA kinematics
        kernel (cavgym_bodies_step); there is no host implementation."""
        from ..engine import bodies_step
        self.throttle, self.steering_angle = action
        if abs(self.steering_angle) < STEERING_ERROR:
            self.steering_angle = 0.0
        x, y, v, theta = bodies_step(self.constants, [list(self.state)], [list(action)], time_resolution)[0]
        self.state = DynamicBodyState(Point(x, y), v, theta)


class Pedestrian(DynamicBody):
    pass


@dataclass(frozen=True)
class SpawnPedestrianState:
    position_boxes: list
    velocity: float
    orientations: list


class SpawnPedestrian(Pedestrian):
    """Pedestrian whose initial state is re-drawn at every reset (reference :290-312).
    With a caller-supplied `np_random` (numpy RandomState, as the reference's Config.setup
    passes) the draw happens on the host with that generator, for seed compatibility;
    in batched runs the engine draws on the device from its Philox stream."""

    def __init__(self, spawn_init_state, constants, np_random=None):
        self.spawn_init_state = spawn_init_state
        self.np_random = np_random if np_random is not None else np.random.RandomState()
        super().__init__(self.spawn(), constants)

    def reset(self):
        self.init_state = self.spawn()
        super().reset()

    def spawn(self):
        boxes = self.spawn_init_state.position_boxes
        areas = [box.area() for box in boxes]
        total_area = sum(areas)
        box = self.np_random.choice(boxes, p=[area / total_area for area in areas])
        position = box.random_point(self.np_random)
        orientation = self.np_random.choice(self.spawn_init_state.orientations)
        return DynamicBodyState(position=position, velocity=self.spawn_init_state.velocity, orientation=orientation)


class Vehicle(DynamicBody):
    def __init__(self, init_state, constants):
        super().__init__(init_state, constants)
        self.indicators_shape = geometry.make_rectangle(constants.length * 0.8, constants.width)
        self.longitudinal_lights_shape = geometry.make_rectangle(constants.length, constants.width * 0.6)

    def indicators(self):
        return self.indicators_shape.transform(self.state.orientation, self.state.position)

    def longitudinal_lights(self):
        return self.longitudinal_lights_shape.transform(self.state.orientation, self.state.position)


class Car(Vehicle):
    def __init__(self, init_state, constants):
        super().__init__(init_state, constants)
        self.roof_shape = geometry.make_rectangle(constants.length * 0.5, constants.width)

    def roof(self):
        return self.roof_shape.transform(self.state.orientation, self.state.position)


class Bus(Vehicle):
    pass


class Bicycle(DynamicBody):
    pass


class TrafficLightState(Enum):
    RED = 0
    AMBER = 1
    GREEN = 2

    def __copy__(self):
        return TrafficLightState(self)

    def __iter__(self):
        yield self


_LIGHT_AFTER = {TrafficLightAction.TURN_RED: TrafficLightState.RED, TrafficLightAction.TURN_AMBER: TrafficLightState.AMBER,
                TrafficLightAction.TURN_GREEN: TrafficLightState.GREEN}


@dataclass(frozen=True)
class TrafficLightConstants:
    width: int
    height: int
    position: Point
    orientation: float


class TrafficLight(Body, Occlusion):
    def __init__(self, init_state, constants):
        super().__init__(init_state=init_state, constants=constants)
        self.static_bounding_box = geometry.make_rectangle(constants.width, constants.height).transform(
            constants.orientation, constants.position)
        self.red_light, self.amber_light, self.green_light = (
            Point(0.0, y).transform(constants.orientation, constants.position)
            for y in (constants.height * 0.25, 0.0, -constants.height * 0.25))

    def observation_space(self):
        return list(TrafficLightState)

    def action_space(self):
        return list(TrafficLightAction)

    def bounding_box(self):
        return self.static_bounding_box

    def step(self, action, time_resolution):
        self.state = _LIGHT_AFTER.get(TrafficLightAction(action), self.state)


@dataclass(frozen=True)
class PelicanCrossingConstants:
    road: Road
    width: int
    x_position: int


class PelicanCrossing(Body):
    def __init__(self, init_state, constants):
        super().__init__(init_state, constants)
        road = constants.road
        heading, origin = road.constants.orientation, road.constants.position
        position = Point(constants.x_position, 0.0).transform(heading, origin)
        self.static_bounding_box = geometry.make_rectangle(constants.width, road.width).transform(heading, position)
        self.outbound_intersection_bounding_box = geometry.make_rectangle(constants.width, road.outbound.width).transform(
            heading, Point(0, (road.inbound.width - position.y) * 0.5).translate(position))
        self.inbound_intersection_bounding_box = geometry.make_rectangle(constants.width, road.inbound.width).transform(
            heading, Point(0, -(road.outbound.width - position.y) * 0.5).translate(position))
        box = self.static_bounding_box
        self.outbound_traffic_light = TrafficLight(init_state, TrafficLightConstants(
            width=10, height=20, position=Point(box.rear_left.x, box.rear_left.y + 20.0), orientation=heading))
        self.inbound_traffic_light = TrafficLight(init_state, TrafficLightConstants(
            width=10, height=20, position=Point(box.front_right.x, box.front_right.y - 20.0), orientation=heading))
        lane = road.constants.lane_width
        self.outbound_spawn = Point(constants.x_position + (constants.width * 0.15), (road.width / 2.0) + (lane / 2.0)).transform(heading, origin)
        self.inbound_spawn = Point(constants.x_position - (constants.width * 0.15), -(road.width / 2.0) - (lane / 2.0)).transform(heading, origin)

    def observation_space(self):
        return spaces.Discrete(len(TrafficLightState))

    def action_space(self):
        return spaces.Discrete(len(TrafficLightAction))

    def bounding_box(self):
        return self.static_bounding_box

    def step(self, action, time_resolution):
        self.state = _LIGHT_AFTER.get(TrafficLightAction(action), self.state)
        self.outbound_traffic_light.step(action, time_resolution)
        self.inbound_traffic_light.step(action, time_resolution)
