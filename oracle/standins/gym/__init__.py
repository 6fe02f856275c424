"""Minimal stand-in for the `gym==0.17.2` surface the CAV-Gym reference touches.

TEST INFRASTRUCTURE ONLY (oracle side).  gym is not installed in this image and
cannot be fetched, so the behaviours below are restated from the 0.17.2
release: `Env`, `register`/`make` (no TimeLimit wrapper when
`max_episode_steps` is None), `spaces.{Box,Discrete,Tuple}` and
`utils.seeding.np_random`.  Call sites in the reference: config.py:275-285,
library/environment.py:59-80,120, library/bodies.py:97-109,441-445,
examples/agents/template.py:54, examples/__init__.py:3-21.
"""
import importlib

from . import spaces  # noqa: F401
from . import utils  # noqa: F401
from .utils import seeding  # noqa: F401


class Env:
    metadata = {'render.modes': []}
    reward_range = (-float('inf'), float('inf'))
    spec = None
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode='human'):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return

    @property
    def unwrapped(self):
        return self


class _Spec:
    def __init__(self, id, entry_point, kwargs):
        self.id = id
        self.entry_point = entry_point
        self.kwargs = dict(kwargs or {})

    def make(self, **kwargs):
        merged = dict(self.kwargs)
        merged.update(kwargs)
        if callable(self.entry_point):
            cls = self.entry_point
        else:
            module_name, attr = self.entry_point.split(":")
            cls = getattr(importlib.import_module(module_name), attr)
        env = cls(**merged)
        env.unwrapped.spec = self
        return env


_registry = {}


def register(id, entry_point=None, kwargs=None, **_ignored):
    if id in _registry:
        raise RuntimeError(f"Cannot re-register id: {id}")
    _registry[id] = _Spec(id, entry_point, kwargs)


def make(id, **kwargs):
    if id not in _registry:
        raise KeyError(f"No registered env with id: {id}")
    return _registry[id].make(**kwargs)
