"""'DenseTraffic': the dense-traffic stress scenario of BASELINE.json (config C4) — 64 cars and 256 spawned
pedestrians per environment, terminate_collisions = "all", so CAVEnv.step's all-pairs collision test
(reference library/environment.py:156-177) meets 320 * 319 / 2 = 51,040 box pairs per step.

SYNTHETIC: the reference has no such scenario.  It is assembled from the reference's own vocabulary only
(Road / RoadMap, Car, SpawnPedestrian, the stock body constants), in the style of
examples/environments/pedestrians.py, so the scenario compiler and every kernel see nothing new:

  * one straight road of `length_m` metres, `lanes_per_direction` lanes each way (lane width 3.65 m as in the stock
    scenarios), position (0, 0), orientation 0;
  * cars on a lane lattice: `num_cars` spread evenly over the lanes, equally spaced along each lane, heading along
    their lane's direction at 75 % of their maximum velocity (the stock ego's speed); body 0 — the ego — is the first car
    of the first outbound lane, at the road's rear edge like the stock ego;
  * pedestrians on two pavements, one per side, in `rows` rows per pavement: each SpawnPedestrian owns ONE small spawn
    box (a cell of the lattice, so freshly spawned pedestrians never overlap) and one orientation (its row's walking
    direction), and walks at 1.4 m/s.

The lattice is wide enough that an episode survives for O(100) steps once agents start to act (crossing pedestrians,
cars that brake or steer), i.e. the broad phase rejects almost everything and the narrow phase sees a few pairs per step.
"""
import math

from ...library import geometry
from ...library.assets import Road, RoadConstants, RoadMap
from ...library.bodies import Car, DynamicBodyState, SpawnPedestrian, SpawnPedestrianState
from ...library.environment import CAVEnv, CAVEnvConstants
from ..constants import M2PX, car_constants, pedestrian_constants


def make_world(length_m=400, lanes_per_direction=4, rows=2, row_pitch_m=2.5):
    road_map = RoadMap(major_road=Road(RoadConstants(
        length=M2PX * length_m, num_outbound_lanes=lanes_per_direction, num_inbound_lanes=lanes_per_direction,
        lane_width=M2PX * 3.65, position=geometry.Point(0.0, 0.0), orientation=0.0)))
    pavement_width = M2PX * row_pitch_m * (rows + 0.5)
    constants = CAVEnvConstants(
        viewer_width=int(road_map.major_road.constants.length),
        viewer_height=int(road_map.major_road.width + pavement_width * 2),
        road_map=road_map)
    return road_map, constants


def make_bodies(num_cars=64, num_pedestrians=256, length_m=400, lanes_per_direction=4, rows=2, row_pitch_m=2.5,
                spawn_box_m=(2.5, 0.6), np_random=None, road_map=None, order="class"):
    if road_map is None:
        road_map, _ = make_world(length_m, lanes_per_direction, rows, row_pitch_m)
    road = road_map.major_road
    length = road.constants.length
    lanes = [(lane, road.outbound.orientation) for lane in road.outbound.lanes] + \
            [(lane, road.inbound.orientation) for lane in road.inbound.lanes]
    per_lane = math.ceil(num_cars / len(lanes))
    cruise = car_constants.min_velocity + (car_constants.max_velocity - car_constants.min_velocity) * 0.75
    cars = []
    for i in range(num_cars):
        lane, orientation = lanes[i % len(lanes)]
        # distance from the lane's own rear edge: the ego starts ON the edge like the stock ego; every other car starts a
        # car length inside, so that no road share sits exactly on the 0.5 liveness threshold (environment.py:144)
        along = (i // len(lanes)) * (length / per_lane) + (0 if i == 0 else car_constants.length)
        position = geometry.Point(along, 0.0).rotate(orientation).translate(lane.spawn)
        cars.append(Car(init_state=DynamicBodyState(position=position, velocity=cruise, orientation=orientation),
                        constants=car_constants))

    box_length, box_width = M2PX * spawn_box_m[0], M2PX * spawn_box_m[1]
    row_pitch = M2PX * row_pitch_m
    tracks = [(side, row) for side in (1, -1) for row in range(rows)]   # +1: the outbound (left) side of the road
    per_track = math.ceil(num_pedestrians / len(tracks))
    pitch = length / per_track
    assert pitch >= box_length + 2 * pedestrian_constants.length, "pedestrian lattice too tight for this road length"
    pedestrians = []
    for i in range(num_pedestrians):
        side, row = tracks[i % len(tracks)]
        column = i // len(tracks)
        centre = geometry.Point((column + 0.5) * pitch, side * (road.width / 2 + row_pitch * (row + 0.75)))
        orientation = road.outbound.orientation if (row + (side < 0)) % 2 == 0 else road.inbound.orientation
        pedestrians.append(SpawnPedestrian(
            spawn_init_state=SpawnPedestrianState(
                position_boxes=[geometry.make_rectangle(box_length, box_width).transform(0.0, centre)],
                velocity=M2PX * 1.4,
                orientations=[orientation]),
            constants=pedestrian_constants,
            np_random=np_random))
    # Body order: the ego first (environment.py:86), then by position along the road, so that consecutive bodies are
    # neighbours in space (the engine's collision broad phase culls whole groups of 32 consecutive bodies).
    # order="class" (default): cars along the road, then pedestrians along the road — groups of 32 are also of one class, so
    # the lanes of a warp run the same agent code (measured 6 % faster than order="road", everything interleaved).
    if order == "class":
        return [cars[0]] + sorted(cars[1:], key=_along) + sorted(pedestrians, key=_along)
    rest = sorted(cars[1:] + pedestrians, key=lambda body: _along(body))
    return [cars[0]] + rest


def _along(body):
    if isinstance(body, SpawnPedestrian):
        box = body.spawn_init_state.position_boxes[0]
        return sum(x for x, _ in box) / 4
    return body.init_state.position.x


class DenseTrafficEnv(CAVEnv):
    def __init__(self, num_cars=64, num_pedestrians=256, np_random=None, **kwargs):
        road_map, constants = make_world()
        super().__init__(bodies=make_bodies(num_cars, num_pedestrians, np_random=np_random, road_map=road_map),
                         constants=constants, np_random=np_random, **kwargs)
