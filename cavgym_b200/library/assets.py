"""Static world description (reference library/assets.py:45-175): roads, lanes, road map,
obstacle, bus stop.  Pure data — the scenario compiler (cavgym_b200/scenario.py) flattens the
rectangles into the engine's constant tables.  `Occlusion.occlusion_zone` (reference :18-42)
is render-only and is not provided."""
import math
from dataclasses import dataclass

from . import geometry
from .geometry import Point


class Occlusion:
    """Marker base: entities that take part in all-pairs collision (environment.py:94-101)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def bounding_box(self):
        raise NotImplementedError


class _StaticBox:
    static_bounding_box = None

    def bounding_box(self):
        return self.static_bounding_box


class Lane(_StaticBox):
    def __init__(self, bounding_box):
        self.static_bounding_box = bounding_box
        self.spawn = bounding_box.rear_centre()


class Direction(_StaticBox):
    def __init__(self, bounding_box, num_lanes, lane_width, orientation):
        self.static_bounding_box = bounding_box
        self.num_lanes, self.lane_width, self.orientation = num_lanes, lane_width, orientation
        self.width = lane_width * num_lanes
        self.lanes = [Lane(box) for box in bounding_box.divide_laterally(num_lanes)] if num_lanes > 0 else []
        self.bus_stop = None

    def set_bus_stop(self, bus_stop):
        self.bus_stop = bus_stop


@dataclass(frozen=True)
class RoadConstants:
    length: int
    num_outbound_lanes: int
    num_inbound_lanes: int
    lane_width: int
    position: Point
    orientation: float


class Road(_StaticBox):
    def __init__(self, constants):
        self.constants = constants
        self.num_lanes = constants.num_outbound_lanes + constants.num_inbound_lanes
        self.width = self.num_lanes * constants.lane_width
        self.static_bounding_box = geometry.make_rectangle(constants.length, self.width, rear_offset=0).transform(
            constants.orientation, constants.position)
        left, right = self.static_bounding_box.split_laterally(left_percentage=constants.num_outbound_lanes / self.num_lanes)
        self.outbound = Direction(left, constants.num_outbound_lanes, constants.lane_width, constants.orientation)
        self.inbound = Direction(right.flip(), constants.num_inbound_lanes, constants.lane_width,
                                 constants.orientation + math.radians(180.0))

    def spawn_position(self, relative_position):
        return relative_position.rotate(self.constants.orientation).translate(self.constants.position)

    def spawn_position_outbound(self, relative_x):
        return Point(relative_x, 0).rotate(self.constants.orientation).translate(self.static_bounding_box.rear_right)

    def spawn_position_inbound(self, relative_x):
        return Point(relative_x, 0).rotate(self.constants.orientation).translate(self.static_bounding_box.rear_left)

    def spawn_orientation(self, relative_orientation):
        return self.constants.orientation + relative_orientation


class RoadMap:
    def __init__(self, major_road, minor_roads=None):
        self.major_road, self.minor_roads = major_road, minor_roads
        self.roads = [major_road] + (minor_roads if minor_roads is not None else [])
        self.obstacle = None
        if minor_roads is not None:
            def partition(boxes, fraction_of):
                return [box.split_longitudinally(rear_percentage=fraction_of(road)) for box, road in zip(boxes, minor_roads)]

            def near(road):
                return major_road.width / road.constants.length

            whole = partition([r.static_bounding_box for r in minor_roads], near)
            outbound = partition([r.outbound.static_bounding_box for r in minor_roads], near)
            inbound = partition([r.inbound.static_bounding_box for r in minor_roads], lambda road: 1 - near(road))
            self.intersection_bounding_boxes = [a for a, _ in whole]
            self.difference_bounding_boxes = [b for _, b in whole]
            self.outbound_intersection_bounding_boxes = [a for a, _ in outbound]
            self.outbound_difference_bounding_boxes = [b for _, b in outbound]
            self.inbound_intersection_bounding_boxes = [b for _, b in inbound]
            self.inbound_difference_bounding_boxes = [a for a, _ in inbound]

    def set_obstacle(self, obstacle):
        self.obstacle = obstacle


@dataclass(frozen=True)
class ObstacleConstants:
    width: int
    height: int
    position: Point
    orientation: float


class Obstacle(_StaticBox, Occlusion):
    def __init__(self, constants, **kwargs):
        super().__init__(**kwargs)
        self.constants = constants
        self.static_bounding_box = geometry.make_rectangle(constants.width, constants.height).transform(
            constants.orientation, constants.position)


@dataclass(frozen=True)
class BusStopConstants:
    road_direction: Direction
    x_position: float
    length: float


class BusStop(_StaticBox):
    def __init__(self, constants):
        self.constants = constants
        direction = constants.road_direction
        self.static_bounding_box = geometry.make_rectangle(constants.length, direction.lane_width * 0.75, left_offset=0).translate(
            Point(constants.x_position, 0).translate(direction.static_bounding_box.rear_left))
