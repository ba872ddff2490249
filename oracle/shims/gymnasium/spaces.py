"""Inert space containers (the reference never samples or validates with them)."""
import numpy as np


class Space:
    pass


class Discrete(Space):
    def __init__(self, n, start=0, seed=None):
        self.n, self.start = int(n), int(start)

    def contains(self, x):
        return self.start <= int(x) < self.start + self.n


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec, self.dtype = np.asarray(nvec), dtype
        self.shape = self.nvec.shape


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class MultiBinary(Space):
    def __init__(self, n, seed=None):
        self.n = n
        self.shape = tuple(n) if isinstance(n, (tuple, list)) else (n,)


class Dict(Space):
    def __init__(self, spaces=None, seed=None):
        self.spaces = dict(spaces or {})

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()
