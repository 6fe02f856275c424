"""Body constants of the stock scenarios in pixels, 16 px per metre (values from reference
examples/constants.py:5-57: car 4.5x1.75 m, pedestrian 0.65625x0.875 m, bus 12x2.55 m,
bicycle 2.25x0.875 m).  Limits are symmetric: throttle in +-max_velocity per second,
steering in +-fraction*pi."""
import math

from ..library.bodies import DynamicBodyConstants

M2PX = 16  # pixels per metre


def _metres(length, width, wheelbase, max_velocity, steering_fraction):
    return DynamicBodyConstants(
        length=M2PX * length, width=M2PX * width, wheelbase=M2PX * wheelbase, track=M2PX * width,
        min_velocity=0, max_velocity=M2PX * max_velocity,
        min_throttle=-(M2PX * max_velocity), max_throttle=M2PX * max_velocity,
        min_steering_angle=-(math.pi * steering_fraction), max_steering_angle=math.pi * steering_fraction)


car_constants = _metres(4.5, 1.75, 3, 9, 0.2)
pedestrian_constants = _metres(0.65625, 0.875, 0.328125, 1.4, 0.4)
bus_constants = _metres(12, 2.55, 16.875, 6.75, 0.16)
bicycle_constants = _metres(2.25, 0.875, 2.025, 4.5, 0.5)
