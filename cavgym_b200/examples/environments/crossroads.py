"""'Crossroads-v0': two cars on the major road, two minor roads forming a staggered crossroads,
one standing pedestrian (scenario data as reference examples/environments/crossroads.py:9-82).
`make_bodies()` builds fresh Body objects per environment; the module-level `bodies` list is
kept for source compatibility with the reference."""
import math

from ...library import geometry
from ...library.assets import Road, RoadConstants, RoadMap
from ...library.bodies import Car, DynamicBodyState, Pedestrian
from ...library.environment import CAVEnv, CAVEnvConstants
from ..constants import M2PX, car_constants, pedestrian_constants

LANE = M2PX * 3.65


def _two_lane_road(length, position, orientation):
    return Road(RoadConstants(length=length, num_outbound_lanes=1, num_inbound_lanes=1, lane_width=LANE,
                              position=position, orientation=orientation))


major_road = _two_lane_road(M2PX * 99, geometry.Point(0.0, 0.0), 0.0)
_half = major_road.constants.length * 0.5
_quarter_lane = major_road.constants.lane_width * 0.25

road_map = RoadMap(
    major_road=major_road,
    minor_roads=[
        _two_lane_road(M2PX * 24.75, major_road.spawn_position_inbound(_half - _quarter_lane),
                       major_road.spawn_orientation(math.radians(270.0))),
        _two_lane_road(M2PX * 24.75, major_road.spawn_position_outbound(_half + _quarter_lane),
                       major_road.spawn_orientation(math.radians(90.0))),
    ])

env_constants = CAVEnvConstants(
    viewer_width=major_road.constants.length,
    viewer_height=sum(minor_road.constants.length for minor_road in road_map.minor_roads) - major_road.width,
    road_map=road_map)


def make_bodies():
    cruising = car_constants.max_velocity / 2.0
    return [
        Car(DynamicBodyState(major_road.outbound.lanes[0].spawn, cruising, major_road.outbound.orientation), car_constants),
        Car(DynamicBodyState(major_road.inbound.lanes[0].spawn, cruising, major_road.inbound.orientation), car_constants),
        Pedestrian(DynamicBodyState(geometry.Point(160, -20).translate(road_map.intersection_bounding_boxes[0].front_left),
                                    0.0, major_road.inbound.orientation), pedestrian_constants),
    ]


bodies = make_bodies()


class CrossroadsEnv(CAVEnv):
    def __init__(self, **kwargs):
        super().__init__(bodies=make_bodies(), constants=env_constants, **kwargs)
