"""Scenario registry: the four env ids the reference registers with gym (examples/__init__.py:3-21), resolvable
with `make(id, **kwargs)` exactly like `gym.make` (no TimeLimit wrapper, kwargs forwarded to the env class)."""
import importlib

_REGISTRY = {}


def register(id, entry_point):
    if id in _REGISTRY:
        raise ValueError(f"Cannot re-register id: {id}")
    _REGISTRY[id] = entry_point


def make(id, **kwargs):
    if id not in _REGISTRY:
        raise KeyError(f"No registered env with id: {id}")
    entry_point = _REGISTRY[id]
    if not callable(entry_point):
        module_name, attr = entry_point.split(":")
        entry_point = getattr(importlib.import_module(module_name), attr)
    env = entry_point(**kwargs)
    env.spec = id
    return env


def registered():
    return sorted(_REGISTRY)


for _id, _module, _cls in (("PelicanCrossing-v0", "pelican_crossing", "PelicanCrossingEnv"),
                           ("BusStop-v0", "bus_stop", "BusStopEnv"),
                           ("Crossroads-v0", "crossroads", "CrossroadsEnv"),
                           ("Pedestrians-v0", "pedestrians", "PedestriansEnv")):
    register(id=_id, entry_point=f"{__name__}.environments.{_module}:{_cls}")
