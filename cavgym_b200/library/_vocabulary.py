"""Discrete action vocabulary of the non-kinematic bodies (reference library/actions.py:4-13): the value a
PelicanCrossing / TrafficLight body takes as its action, also what the engine reads from slot [.][0] of the actions tensor."""
from enum import Enum


class Action(Enum):
    """Base of the discrete action enums; prints as its bare name, as the reference's log lines show it."""

    def __repr__(self):
        return self.name


TRAFFIC_LIGHT_ACTIONS = ("NOOP", "TURN_RED", "TURN_AMBER", "TURN_GREEN")   # values 0..3 in this order (bodies.py:450-461)
TrafficLightAction = Action("TrafficLightAction", [(name, value) for value, name in enumerate(TRAFFIC_LIGHT_ACTIONS)],
                            module=__name__)
