"""'BusStop-v0': three-lane one-way road with a bus stop; ego car, a bus, two cars and a bicycle
(scenario data as reference examples/environments/bus_stop.py:7-82)."""
from ...library import geometry
from ...library.assets import BusStop, BusStopConstants, Road, RoadConstants, RoadMap
from ...library.bodies import Bicycle, Bus, Car, DynamicBodyState
from ...library.environment import CAVEnv, CAVEnvConstants
from ..constants import M2PX, bicycle_constants, bus_constants, car_constants

road_map = RoadMap(major_road=Road(RoadConstants(
    length=M2PX * 99, num_outbound_lanes=3, num_inbound_lanes=0, lane_width=M2PX * 3.65,
    position=geometry.Point(0.0, 0.0), orientation=0.0)))
outbound = road_map.major_road.outbound
outbound.set_bus_stop(BusStop(BusStopConstants(
    road_direction=outbound, x_position=M2PX * 99 * 0.75, length=bus_constants.length * 1.25)))

env_constants = CAVEnvConstants(
    viewer_width=road_map.major_road.constants.length,
    viewer_height=road_map.major_road.width + ((M2PX * 3) * 2),
    road_map=road_map)


def make_bodies():
    heading = outbound.orientation
    lanes = outbound.lanes

    def ahead(distance):
        return geometry.Point(distance, 0).rotate(heading).translate(lanes[0].spawn)

    car_speed, bicycle_speed = car_constants.max_velocity / 2.0, bicycle_constants.max_velocity / 2.0
    return [
        Car(DynamicBodyState(ahead(200), car_speed, heading), car_constants),
        Bus(DynamicBodyState(ahead(400), car_speed, heading), bus_constants),
        Car(DynamicBodyState(lanes[0].spawn, car_speed, heading), car_constants),
        Car(DynamicBodyState(lanes[1].spawn, car_speed, heading), car_constants),
        Bicycle(DynamicBodyState(lanes[2].spawn, bicycle_speed, heading), bicycle_constants),
    ]


bodies = make_bodies()


class BusStopEnv(CAVEnv):
    def __init__(self, **kwargs):
        super().__init__(bodies=make_bodies(), constants=env_constants, **kwargs)
