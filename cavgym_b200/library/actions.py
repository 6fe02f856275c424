"""`library.actions` of the reference (actions.py:4-13) — the names live in `_vocabulary`."""
from ._vocabulary import TRAFFIC_LIGHT_ACTIONS, Action, TrafficLightAction  # noqa: F401
