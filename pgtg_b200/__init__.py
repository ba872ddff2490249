"""pgtg_b200 -- B200-native batched PGTG (ProcGrid Traffic Gym) simulator.

`PGTGVectorEnv` keeps the constructor arguments, spaces and reward / terminated / truncated /
info semantics of the reference `PGTGEnv` (Inuri04/pgtg, pgtg/environment.py:297) but steps all
environments with hand-written sm_100a CUDA kernels behind the C ABI of include/pgtg_b200.h.
There is no CPU fallback: constructing an env without the built CUDA library raises.
"""
__version__ = "0.1.0"

from .config import MapPlan, make_config  # noqa: F401


def __getattr__(name):
    if name in ("PGTGVectorEnv", "make_vec"):
        from . import vector_env

        return getattr(vector_env, name)
    raise AttributeError(name)
