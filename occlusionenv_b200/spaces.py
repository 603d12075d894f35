"""Minimal stand-ins for the two ``gym.spaces`` classes the reference touches (gym is not installed).

``environment.py:221-222`` builds ``Box(0, 1, shape=(4,S,S))`` and ``Box(low=-0.1, high=0.1, shape=(2,))``;
``SubProcVecEnv.py:54-61`` only asks whether a space is a ``Dict`` / ``Tuple`` and reads ``.shape`` / ``.dtype``.
"""
from __future__ import annotations

import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        super().__init__(shape, dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Dict(Space):
    def __init__(self, spaces):
        super().__init__()
        self.spaces = spaces


class Tuple(Space):
    def __init__(self, spaces):
        super().__init__()
        self.spaces = tuple(spaces)
