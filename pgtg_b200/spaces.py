"""Gymnasium spaces for `PGTGVectorEnv`. When the `gymnasium` package is importable its own classes
are used (so `isinstance` checks of consumers work); otherwise a minimal structural mirror of the
classes the reference touches (environment.py:415-441) keeps the same attributes."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - not installed in the build image
    from gymnasium.spaces import Box, Dict, Discrete, MultiBinary, MultiDiscrete  # type: ignore
    from gymnasium.vector.utils import batch_space  # type: ignore

    HAVE_GYMNASIUM = True
except Exception:  # ModuleNotFoundError in this image
    HAVE_GYMNASIUM = False

    class Space:
        shape: tuple = ()
        dtype = None

    class Discrete(Space):
        def __init__(self, n, start=0, seed=None):
            self.n, self.start, self.dtype = int(n), int(start), np.int64

        def contains(self, x):
            return self.start <= int(x) < self.start + self.n

        def __repr__(self):
            return f"Discrete({self.n}, start={self.start})" if self.start else f"Discrete({self.n})"

    class MultiDiscrete(Space):
        def __init__(self, nvec, dtype=np.int64, seed=None):
            self.nvec, self.dtype = np.asarray(nvec), dtype
            self.shape = self.nvec.shape

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape or ()), dtype

        def __repr__(self):
            return f"Box({self.low}, {self.high}, {self.shape}, {np.dtype(self.dtype).name})"

    class MultiBinary(Space):
        def __init__(self, n, seed=None):
            self.n = n
            self.shape = tuple(n) if isinstance(n, (tuple, list)) else (n,)
            self.dtype = np.int8

        def __repr__(self):
            return f"MultiBinary({self.n})"

    class Dict(Space):
        def __init__(self, spaces=None, seed=None):
            self.spaces = dict(spaces or {})

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {v!r}" for k, v in self.spaces.items()) + ")"

    def batch_space(space, n):
        if isinstance(space, Dict):
            return Dict({k: batch_space(v, n) for k, v in space.items()})
        if isinstance(space, Discrete):
            return MultiDiscrete([space.n] * n)
        if isinstance(space, MultiDiscrete):
            return MultiDiscrete(np.tile(space.nvec, (n, 1)), dtype=space.dtype)
        if isinstance(space, Box):
            return Box(space.low, space.high, (n,) + tuple(space.shape), space.dtype)
        if isinstance(space, MultiBinary):
            return MultiBinary((n,) + tuple(space.shape))
        raise TypeError(space)
