"""Target vocabularies of TargetAgent / KeyboardAgent (reference examples/targets.py:5-19)."""
import math
from enum import Enum

TargetVelocity = Enum("TargetVelocity", [("MIN", 0), ("MID", 1), ("MAX", 2)])

TargetOrientation = Enum("TargetOrientation", [
    ("NORTH", math.pi * 0.5), ("NORTH_EAST", math.pi * 0.25), ("EAST", 0.0), ("SOUTH_EAST", -(math.pi * 0.25)),
    ("SOUTH", -(math.pi * 0.5)), ("SOUTH_WEST", -(math.pi * 0.75)), ("WEST", math.pi), ("NORTH_WEST", math.pi * 0.75)])
