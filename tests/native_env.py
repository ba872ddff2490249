"""Adapters giving the parity harness (tests/parity.py) one array-level surface over
  * the host emulation build of the kernel phases (tests/emu, CPU tests), and
  * the CUDA product through its C ABI (GPU tests)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pgtg_b200.config import make_config
from pgtg_b200.raw import RawEnv

_HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(_HERE, "emu")
EMU_LIB = os.path.join(EMU_DIR, "libpgtg_emu.so")

_NP = {"obs_map": np.int8, "obs_position": np.int32, "obs_velocity": np.int32, "obs_next_subgoal_direction": np.int32,
       "reward": np.float64, "cost": np.float64, "terminated": np.uint8, "truncated": np.uint8, "step_state": np.int32,
       "step_flags": np.uint8, "final_obs_map": np.int8, "final_obs_position": np.int32, "final_obs_velocity": np.int32,
       "final_obs_next_subgoal_direction": np.int32, "stats": np.float64}


def build_emu() -> str:
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])
    return EMU_LIB


class NativeAdapter:
    """backend="emu": host pointers viewed in place; backend="cuda": device buffers copied back
    through torch tensors made from the library's DLPack capsules."""

    def __init__(self, backend: str, **kwargs):
        lib_path = None
        if backend.startswith("emu"):  # "emu", or "emu_sat3": the variant whose per-square car counters saturate at 3
            lib_path = build_emu().replace("libpgtg_emu.so", f"libpgtg_{backend}.so")
            backend = "emu"
        self.backend = backend
        self.hc = make_config(**kwargs)
        self.raw = RawEnv(self.hc, device=0, lib_path=lib_path)
        self.N, self.C, self.P = self.raw.N, self.raw.C, self.raw.P
        self._torch = {}
        if backend == "cuda":
            import torch

            self._tmod = torch
            self._actions = torch.zeros(self.N, dtype=torch.int32, device="cuda:0")

    def _shape(self, name):
        N, Cc, P = self.N, self.C, self.P
        return {"obs_map": (N, Cc, P, P), "final_obs_map": (N, Cc, P, P), "obs_position": (N, 2), "obs_velocity": (N, 2),
                "final_obs_position": (N, 2), "final_obs_velocity": (N, 2), "step_state": (N, 4), "stats": (8,)}.get(name, (N,))

    def array(self, name):
        if self.backend == "emu":
            ptr = getattr(self.raw.bufs, name)
            assert ptr, name
            n = int(np.prod(self._shape(name)))
            buf = (C.c_char * (n * np.dtype(_NP[name]).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=_NP[name]).reshape(self._shape(name)).copy()
        if name not in self._torch:
            self._torch[name] = self._tmod.from_dlpack(self.raw.dlpack_capsule(name))
        self._tmod.cuda.synchronize()
        return self._torch[name].cpu().numpy()

    obs_map = property(lambda s: s.array("obs_map"))
    obs_position = property(lambda s: s.array("obs_position"))
    obs_velocity = property(lambda s: s.array("obs_velocity"))
    obs_nsd = property(lambda s: s.array("obs_next_subgoal_direction"))
    reward = property(lambda s: s.array("reward"))
    cost = property(lambda s: s.array("cost"))
    terminated = property(lambda s: s.array("terminated"))
    truncated = property(lambda s: s.array("truncated"))
    step_state = property(lambda s: s.array("step_state"))
    step_flags = property(lambda s: s.array("step_flags"))
    final_obs_map = property(lambda s: s.array("final_obs_map"))
    final_obs_position = property(lambda s: s.array("final_obs_position"))
    final_obs_velocity = property(lambda s: s.array("final_obs_velocity"))
    final_obs_nsd = property(lambda s: s.array("final_obs_next_subgoal_direction"))

    def load_draws(self, v, t, o):
        self.raw.load_draws(v, t, o)

    def reset(self, seeds=None, mask=None):
        self.raw.reset(seeds, mask)

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.int32)
        if self.backend == "emu":
            self._a = a
            self.raw.step_device(a.ctypes.data, 4)
        else:
            self._actions.copy_(self._tmod.from_numpy(a))
            self.raw.step_device(self._actions.data_ptr(), 4, self._tmod.cuda.current_stream().cuda_stream)

    def observe(self):
        self.raw.observe()

    def get_state(self):
        return self.raw.get_state()

    def agent_direction(self):
        return self.raw.get_info()["agent_direction"]

    def set_state(self, **kw):
        self.raw.set_state(**kw)

    def stats(self, reset_after=False):
        return self.raw.stats(reset_after)

    def close(self):
        self.raw.close()
