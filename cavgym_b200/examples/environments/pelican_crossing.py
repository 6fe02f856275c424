"""'PelicanCrossing-v0': two cars approaching a signalled crossing with an obstacle beside it,
two pedestrians waiting at the lights (scenario data as reference
examples/environments/pelican_crossing.py:9-87)."""
import math

from ...library import geometry
from ...library.assets import Obstacle, ObstacleConstants, Road, RoadConstants, RoadMap
from ...library.bodies import (Car, DynamicBodyState, Pedestrian, PelicanCrossing, PelicanCrossingConstants,
                               TrafficLightState)
from ...library.environment import CAVEnv, CAVEnvConstants
from ..constants import M2PX, car_constants, pedestrian_constants

road_map = RoadMap(major_road=Road(RoadConstants(
    length=M2PX * 99, num_outbound_lanes=1, num_inbound_lanes=1, lane_width=M2PX * 3.65,
    position=geometry.Point(0.0, 0.0), orientation=0.0)))
major_road = road_map.major_road

env_constants = CAVEnvConstants(
    viewer_width=major_road.constants.length,
    viewer_height=major_road.width + ((M2PX * 3) * 2),
    road_map=road_map)


def make_pelican_crossing():
    return PelicanCrossing(init_state=TrafficLightState.GREEN, constants=PelicanCrossingConstants(
        road=major_road, width=major_road.constants.lane_width * 1.5, x_position=major_road.constants.length * 0.5))


pelican_crossing = make_pelican_crossing()

road_map.set_obstacle(Obstacle(ObstacleConstants(
    width=M2PX * 3, height=M2PX * 1.5,
    position=geometry.Point(-20, -20).transform(major_road.constants.orientation, pelican_crossing.static_bounding_box.rear_right),
    orientation=major_road.constants.orientation)))


def make_bodies(crossing=None):
    crossing = crossing if crossing is not None else make_pelican_crossing()
    cruising = car_constants.max_velocity / 2.0
    facing = major_road.outbound.orientation
    return [
        Car(DynamicBodyState(major_road.outbound.lanes[0].spawn, cruising, major_road.outbound.orientation), car_constants),
        Car(DynamicBodyState(major_road.inbound.lanes[0].spawn, cruising, major_road.inbound.orientation), car_constants),
        crossing,
        Pedestrian(DynamicBodyState(crossing.inbound_spawn, 0.0, facing + math.radians(90.0)), pedestrian_constants),
        Pedestrian(DynamicBodyState(crossing.outbound_spawn, 0.0, facing + math.radians(270.0)), pedestrian_constants),
    ]


bodies = make_bodies(pelican_crossing)


class PelicanCrossingEnv(CAVEnv):
    def __init__(self, **kwargs):
        super().__init__(bodies=make_bodies(), constants=env_constants, **kwargs)
