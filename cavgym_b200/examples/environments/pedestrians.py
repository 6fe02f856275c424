"""'Pedestrians-v0': one ego car on a straight two-lane road, N pedestrians spawned on the
pavements (scenario data as reference examples/environments/pedestrians.py:11-96)."""
import math

from ...library import geometry
from ...library.assets import Road, RoadConstants, RoadMap
from ...library.bodies import Car, DynamicBodyState, SpawnPedestrian, SpawnPedestrianState
from ...library.environment import CAVEnv, CAVEnvConstants
from ..constants import M2PX, car_constants, pedestrian_constants

road_map = RoadMap(major_road=Road(RoadConstants(
    length=M2PX * 99, num_outbound_lanes=1, num_inbound_lanes=1, lane_width=M2PX * 3.65,
    position=geometry.Point(0.0, 0.0), orientation=0.0)))
major_road = road_map.major_road

pavement_width = M2PX * 3

env_constants = CAVEnvConstants(
    viewer_width=int(major_road.constants.length),
    viewer_height=int(major_road.width + (pavement_width * 2)),
    road_map=road_map)

bounding_box = major_road.bounding_box()


def _pavement(side_corner, offset):
    return geometry.make_rectangle(major_road.constants.length, pavement_width, rear_offset=0).transform(
        major_road.constants.orientation, geometry.Point(0, offset).translate(side_corner))


outbound_pavement = _pavement(bounding_box.rear_left, pavement_width / 2)
inbound_pavement = _pavement(bounding_box.rear_right, -(pavement_width / 2))
pedestrian_diameter = math.sqrt(pedestrian_constants.length ** 2 + pedestrian_constants.width ** 2)
x_scale = 1 - (pedestrian_diameter / major_road.constants.length)
y_scale = 1 - (pedestrian_diameter / pavement_width)
spawn_orientations = [major_road.outbound.orientation, major_road.inbound.orientation]


def make_spawn_position_boxes(outbound_percentage, inbound_percentage):
    """Spawn rectangles: the part `percentage` (from the far end) of each shrunken pavement."""
    assert 0 <= outbound_percentage <= 1
    assert 0 <= inbound_percentage <= 1
    assert outbound_percentage > 0 or inbound_percentage > 0
    boxes = []
    for pavement, percentage in ((outbound_pavement, outbound_percentage), (inbound_pavement, inbound_percentage)):
        if percentage == 0:
            continue
        shrunk = pavement.rescale(x_scale=x_scale, y_scale=y_scale)
        boxes.append(shrunk if percentage == 1 else shrunk.split_longitudinally(1 - percentage)[1])
    return boxes


def make_bodies(num_pedestrians, outbound_percentage, inbound_percentage, np_random=None):
    ego = Car(
        init_state=DynamicBodyState(
            position=major_road.outbound.lanes[0].spawn,
            velocity=car_constants.min_velocity + (car_constants.max_velocity - car_constants.min_velocity) * 0.75,
            orientation=major_road.outbound.orientation),
        constants=car_constants)
    pedestrians = [
        SpawnPedestrian(
            spawn_init_state=SpawnPedestrianState(
                position_boxes=make_spawn_position_boxes(outbound_percentage, inbound_percentage),
                velocity=M2PX * 1.4,
                orientations=spawn_orientations),
            constants=pedestrian_constants,
            np_random=np_random)
        for _ in range(num_pedestrians)]
    return [ego] + pedestrians


class PedestriansEnv(CAVEnv):
    def __init__(self, num_pedestrians, outbound_percentage, inbound_percentage, np_random=None, **kwargs):
        super().__init__(bodies=make_bodies(num_pedestrians, outbound_percentage, inbound_percentage, np_random),
                         constants=env_constants, np_random=np_random, **kwargs)
