"""Per-tick digests of every output of a vector env -- TEST INFRASTRUCTURE ONLY.

A conformance run at the BASELINE sizes (4096 envs x 256 ticks) would need a 700 MB trace; instead the
reference run is condensed into hashes (tests/golden/make_digests.py) and the parity tests recompute the
same hashes from the CUDA buffers. One *row* = every observable of one env at one tick in a fixed
byte layout; h[i, t] = blake2b-128(row). The fixture keeps
    tick_digest[t] = blake2b-256(h[0, t] || h[1, t] || ...)      which tick diverged
    env_digest[i]  = blake2b-256(h[i, 0] || h[i, 1] || ...)      which env diverged
Neither side of the comparison needs the other's arrays.

Row layout (little endian, fixed length for a configuration):
    reward f64 | cost f64 | terminated u8 | truncated u8 | step_state i32[4] | step_flags u8      (zeros at tick 0)
    final_obs_map i8[C,P,P] | final_obs_position i32[2] | final_obs_velocity i32[2] | final_obs_nsd i32
                                                            (zeros unless the env finished in this tick)
    obs_map i8[C,P,P] | obs_position i32[2] | obs_velocity i32[2] | obs_nsd i32
    agent i32[4] | num_cars i32 | cars i32[max_cars,7] (zeros past num_cars) | tiles u16[T] | plan i32[7]
    agent_direction i32
"""
from __future__ import annotations

import hashlib

import numpy as np


def _b(a, dtype, shape):
    a = np.ascontiguousarray(np.asarray(a, dtype=dtype))
    assert a.shape[1:] == shape, (a.shape, shape)
    return a.reshape(a.shape[0], -1).view(np.uint8)


def rows(N, C, P, T, max_cars, *, step=None, final=None, obs, state) -> np.ndarray:
    """-> uint8 [N, L]. `step` = dict(reward, cost, terminated, truncated, step_state, step_flags) or None
    (tick 0); `final` = dict(map, position, velocity, nsd) with rows of unfinished envs ignored;
    `obs` = dict(map, position, velocity, nsd); `state` = dict(agent, num_cars, cars, tiles, plan,
    agent_direction)."""
    parts = []
    if step is None:
        parts.append(np.zeros((N, 8 + 8 + 1 + 1 + 16 + 1), np.uint8))
        done = np.zeros(N, bool)
    else:
        parts += [_b(step["reward"], "<f8", ()), _b(step["cost"], "<f8", ()), _b(step["terminated"], np.uint8, ()),
                  _b(step["truncated"], np.uint8, ()), _b(step["step_state"], "<i4", (4,)), _b(step["step_flags"], np.uint8, ())]
        done = (np.asarray(step["terminated"]) | np.asarray(step["truncated"])).astype(bool)
    if final is None:
        parts.append(np.zeros((N, C * P * P + 8 + 8 + 4), np.uint8))
    else:
        f = np.concatenate([_b(final["map"], np.int8, (C, P, P)), _b(final["position"], "<i4", (2,)), _b(final["velocity"], "<i4", (2,)),
                            _b(final["nsd"], "<i4", ())], axis=1)
        f = np.where(done[:, None], f, 0).astype(np.uint8)
        parts.append(f)
    parts += [_b(obs["map"], np.int8, (C, P, P)), _b(obs["position"], "<i4", (2,)), _b(obs["velocity"], "<i4", (2,)), _b(obs["nsd"], "<i4", ())]
    cars = np.asarray(state["cars"], dtype="<i4")[:, :max_cars]
    if cars.shape[1] < max_cars:
        cars = np.concatenate([cars, np.zeros((N, max_cars - cars.shape[1], 7), "<i4")], axis=1)
    ncar = np.asarray(state["num_cars"], dtype="<i4")
    cars = np.where((np.arange(max_cars)[None, :] < ncar[:, None])[:, :, None], cars, 0).astype("<i4")
    parts += [_b(state["agent"], "<i4", (4,)), _b(ncar, "<i4", ()), _b(cars, "<i4", (max_cars, 7)), _b(state["tiles"], "<u2", (T,)),
              _b(np.asarray(state["plan"])[:, :7], "<i4", (7,)), _b(state["agent_direction"], "<i4", ())]
    return np.ascontiguousarray(np.concatenate(parts, axis=1))


def row_hashes(r: np.ndarray) -> np.ndarray:
    """uint8 [N, L] -> uint8 [N, 16]"""
    out = np.empty((r.shape[0], 16), np.uint8)
    for i in range(r.shape[0]):
        out[i] = np.frombuffer(hashlib.blake2b(r[i].tobytes(), digest_size=16).digest(), np.uint8)
    return out


def fold(h: np.ndarray):
    """h uint8 [N, T1, 16] -> (tick_digest uint8 [T1, 32], env_digest uint8 [N, 32])"""
    N, T1, _ = h.shape
    tick = np.stack([np.frombuffer(hashlib.blake2b(np.ascontiguousarray(h[:, t]).tobytes(), digest_size=32).digest(), np.uint8) for t in range(T1)])
    env = np.stack([np.frombuffer(hashlib.blake2b(np.ascontiguousarray(h[i]).tobytes(), digest_size=32).digest(), np.uint8) for i in range(N)])
    return tick, env


class LiveDigester:
    """Feeds on an env adapter (oracle / emu / CUDA, tests/native_env.py surface) tick by tick."""

    def __init__(self, env, N, C, P, T, max_cars):
        self.env, self.dims = env, (N, C, P, T, max_cars)
        self.h = []

    def _obs(self, final=False):
        e = self.env
        if final:
            return dict(map=e.final_obs_map, position=e.final_obs_position, velocity=e.final_obs_velocity, nsd=e.final_obs_nsd)
        return dict(map=e.obs_map, position=e.obs_position, velocity=e.obs_velocity, nsd=e.obs_nsd)

    def _state(self):
        st = self.env.get_state()
        st["agent_direction"] = self.env.agent_direction()
        return st

    def after_reset(self):
        self.h.append(row_hashes(rows(*self.dims, obs=self._obs(), state=self._state())))

    def after_step(self):
        e = self.env
        step = dict(reward=e.reward, cost=e.cost, terminated=e.terminated, truncated=e.truncated, step_state=e.step_state, step_flags=e.step_flags)
        self.h.append(row_hashes(rows(*self.dims, step=step, final=self._obs(True), obs=self._obs(), state=self._state())))

    def digests(self):
        return fold(np.stack(self.h, axis=1))
