"""Stand-in for the `gymnasium` package (absent from this image, no network).

TEST INFRASTRUCTURE ONLY. It exists so that the UNMODIFIED reference
(`/root/reference/pgtg/environment.py`) can be imported and run in the build
container to record golden traces (see oracle/ref_runner.py). It is never
imported by the product package `pgtg_b200`.

Surface touched by the reference:
  * gym.Env base class with `reset(seed=...)` and the `np_random` property
    (environment.py:297, 591, 599) -- gymnasium 0.28.1 semantics:
    `np_random = numpy.random.Generator(PCG64(SeedSequence(seed)))`.
  * gymnasium.spaces.{Discrete, MultiDiscrete, Box, MultiBinary, Dict}
    (environment.py:415-441) -- inert containers here.
  * gymnasium.envs.registration.register (pgtg/__init__.py:1,7).

A hook (`RNG_FACTORY`) lets the trace recorder wrap the parent generator so the
five child streams spawned at environment.py:593-599 are recording proxies.
"""
import numpy as np

from . import spaces  # noqa: F401
from . import envs  # noqa: F401

# Optional callable(np.random.Generator) -> generator-like; set by oracle/ref_runner.py
RNG_FACTORY = None


class Env:
    metadata = {}
    render_mode = None
    _np_random = None

    @property
    def np_random(self):
        if self._np_random is None:
            g = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
            self._np_random = RNG_FACTORY(g) if RNG_FACTORY else g
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            g = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
            self._np_random = RNG_FACTORY(g) if RNG_FACTORY else g

    def step(self, action):
        raise NotImplementedError

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self


_REGISTRY = {}


def make(id, **kwargs):
    import importlib

    mod, cls = _REGISTRY[id].split(":")
    return getattr(importlib.import_module(mod), cls)(**kwargs)
