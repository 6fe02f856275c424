"""Action/observation space descriptors for the compat API.

The reference builds `gym.spaces.{Box,Discrete,Tuple}` (gym==0.17.2,
library/bodies.py:97-109,441-445; library/environment.py:76-80) and only uses
`contains` (inclusive bounds), `sample`, iteration/indexing of Tuple and the
`np_random` attribute.  gym is not a dependency of this engine, so the same
three descriptors are provided here with those semantics.
"""
import numpy as np


class Space:
    np_random = None

    def _rng(self):
        if self.np_random is None:
            self.np_random = np.random.RandomState()
        return self.np_random

    def __contains__(self, x):
        return self.contains(x)


class Box(Space):
    def __init__(self, low, high, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.low = np.asarray(low, dtype=self.dtype)
        self.high = np.asarray(high, dtype=self.dtype)
        if self.low.shape != self.high.shape:
            raise ValueError("low and high must have the same shape")
        self.shape = self.low.shape

    def is_bounded(self):
        return bool(np.all(np.isfinite(self.low)) and np.all(np.isfinite(self.high)))

    def sample(self):
        """uniform(low, high) per component; every action Box of the reference is bounded."""
        if not self.is_bounded():
            raise NotImplementedError("sampling an unbounded Box is not on the CAV-Gym path")
        return self._rng().uniform(low=self.low, high=self.high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self):
        return self._rng().randint(self.n)

    def contains(self, x):
        if isinstance(x, (np.generic, np.ndarray)):
            if x.shape != () or x.dtype.kind not in "iu":
                return False
            x = int(x)
        return isinstance(x, int) and not isinstance(x, bool) and 0 <= x < self.n


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = list(spaces)

    def sample(self):
        return tuple(space.sample() for space in self.spaces)

    def contains(self, x):
        return isinstance(x, (list, tuple)) and len(x) == len(self.spaces) and all(
            space.contains(part) for space, part in zip(self.spaces, x))

    def __getitem__(self, index):
        return self.spaces[index]

    def __len__(self):
        return len(self.spaces)

    def __iter__(self):
        return iter(self.spaces)
