"""ctypes front-end of the CPU oracle (oracle/pgtg_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
`--impl reference` legs. Never imported by pgtg_b200.

`OracleVectorEnv` has the same array-level surface as the product's C ABI (same pgtg_config,
same output layouts) so parity tests compare arrays one to one.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pgtg_b200.config import HostConfig, PgtgConfig, make_config

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpgtg_oracle.so")
    src = os.path.join(_HERE, "pgtg_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ora_create.restype = C.c_void_p
        _LIB.ora_create.argtypes = [C.POINTER(PgtgConfig)]
        for name in ("ora_destroy", "ora_set_threads", "ora_stats"):
            getattr(_LIB, name).restype = None
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PgtgState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "agent", "flat_tire", "light_counter", "elapsed", "num_cars", "cars", "tiles", "plan", "used",
        "draw_cursor", "error")]


class OracleVectorEnv:
    def __init__(self, host_cfg: HostConfig | None = None, threads: int = 1, **kwargs):
        self.hc = host_cfg if host_cfg is not None else make_config(**kwargs)
        c = self.hc.pod
        self.N, self.C, self.P = c.num_envs, c.num_channels, self.hc.window
        self.T = c.map_w * c.map_h
        self.max_cars = max(1, c.max_cars)
        L = lib()
        self._h = C.c_void_p(L.ora_create(C.byref(c)))
        L.ora_set_threads(self._h, int(threads))
        if self.hc.map_plan is not None:
            mp = self.hc.map_plan
            from pgtg_b200._names import CARDINALS

            rc = L.ora_load_fixed_map(self._h, mp.packed_tiles(), mp.width, mp.height,
                                      int(mp.start[0]), int(mp.start[1]), CARDINALS.index(mp.start[2]),
                                      int(mp.goal[0]), int(mp.goal[1]), CARDINALS.index(mp.goal[2]))
            assert rc == 0
        N, Cc, P = self.N, self.C, self.P
        self.obs_map = np.zeros((N, Cc, P, P), np.int8)
        self.obs_position = np.zeros((N, 2), np.int32)
        self.obs_velocity = np.zeros((N, 2), np.int32)
        self.obs_nsd = np.zeros(N, np.int32)
        self.reward = np.zeros(N, np.float64)
        self.cost = np.zeros(N, np.float64)
        self.terminated = np.zeros(N, np.uint8)
        self.truncated = np.zeros(N, np.uint8)
        self.step_state = np.zeros((N, 4), np.int32)
        self.step_flags = np.zeros(N, np.uint8)
        self.final_obs_map = np.zeros((N, Cc, P, P), np.int8)
        self.final_obs_position = np.zeros((N, 2), np.int32)
        self.final_obs_velocity = np.zeros((N, 2), np.int32)
        self.final_obs_nsd = np.zeros(N, np.int32)
        self._tape = None

    def load_draws(self, values, tags, offsets):
        self._tape = (np.ascontiguousarray(values, np.float64), np.ascontiguousarray(tags, np.uint8),
                      np.ascontiguousarray(offsets, np.int64))
        rc = lib().ora_load_draws(self._h, _p(self._tape[0]), _p(self._tape[1]), _p(self._tape[2]))
        assert rc == 0, rc

    def reset(self, seeds=None, mask=None):
        s = None if seeds is None else np.ascontiguousarray(seeds, np.int64)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        rc = lib().ora_reset(self._h, _p(s), _p(m), _p(self.obs_map), _p(self.obs_position), _p(self.obs_velocity), _p(self.obs_nsd))
        assert rc == 0, rc

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.int32)
        assert a.shape == (self.N,)
        rc = lib().ora_step(self._h, _p(a), _p(self.obs_map), _p(self.obs_position), _p(self.obs_velocity), _p(self.obs_nsd),
                            _p(self.reward), _p(self.cost), _p(self.terminated), _p(self.truncated), _p(self.step_state),
                            _p(self.step_flags), _p(self.final_obs_map), _p(self.final_obs_position),
                            _p(self.final_obs_velocity), _p(self.final_obs_nsd))
        assert rc == 0, rc

    def observe(self):
        rc = lib().ora_observe(self._h, _p(self.obs_map), _p(self.obs_position), _p(self.obs_velocity), _p(self.obs_nsd))
        assert rc == 0

    def get_state(self) -> dict:
        N, T, MC = self.N, self.T, self.max_cars
        out = dict(
            agent=np.zeros((N, 4), np.int32), flat_tire=np.zeros(N, np.uint8), light_counter=np.zeros(N, np.int32),
            elapsed=np.zeros(N, np.int32), num_cars=np.zeros(N, np.int32), cars=np.zeros((N, MC, 7), np.int32),
            tiles=np.zeros((N, T), np.uint16), plan=np.zeros((N, 8), np.int32), used=np.zeros((N, T), np.uint8),
            draw_cursor=np.zeros(N, np.int64), error=np.zeros(N, np.int32))
        st = PgtgState(**{k: v.ctypes.data for k, v in out.items()})
        rc = lib().ora_get_state(self._h, C.byref(st))
        assert rc == 0
        return out

    def set_state(self, agent=None, flat_tire=None, num_cars=None, cars=None):
        keep = []
        kw = {}
        for name, arr, dt in (("agent", agent, np.int32), ("flat_tire", flat_tire, np.uint8), ("num_cars", num_cars, np.int32), ("cars", cars, np.int32)):
            if arr is not None:
                a = np.ascontiguousarray(arr, dt)
                keep.append(a)
                kw[name] = a.ctypes.data
        st = PgtgState(**kw)
        rc = lib().ora_set_state(self._h, C.byref(st))
        assert rc == 0

    def agent_direction(self):
        out = np.zeros(self.N, np.int32)
        assert lib().ora_agent_direction(self._h, _p(out)) == 0
        return out

    def stats(self, reset_after=False):
        out = np.zeros(8, np.float64)
        lib().ora_stats(self._h, _p(out), int(reset_after))
        return out

    def close(self):
        if self._h:
            lib().ora_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decompose_velocity(dx: int, dy: int):
    out = np.zeros((max(abs(dx), abs(dy)), 2), np.int32)
    n = lib().ora_decompose_velocity(int(dx), int(dy), _p(out))
    return out[:n]


def philox(ctr, key):
    c = np.array(ctr, np.uint32)
    lib().ora_philox(_p(c), C.c_uint32(key[0]), C.c_uint32(key[1]))
    return c
