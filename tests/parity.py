"""Shared parity checker: replays a recorded reference trace through a vector env (the CPU oracle
or the CUDA product -- both expose the same arrays) and demands bit equality at every tick."""
from __future__ import annotations

import json

import numpy as np


def load_trace(path):
    z = np.load(path)
    tr = {k: z[k] for k in z.files}
    tr["meta"] = json.loads(bytes(tr["meta"]).decode())
    return tr


def trace_kwargs(tr):
    kw = dict(tr["meta"]["kwargs"])
    for k in ("random_map_start_position", "random_map_goal_position", "traffic_light_phases_duration"):
        if isinstance(kw.get(k), list):
            kw[k] = tuple(kw[k])
    if "map_plan" in kw:  # fixed maps travel inside the trace
        kw["map_plan"] = tr["meta"]["maps"][kw["map_plan"]]
    kw["num_envs"] = tr["meta"]["num_envs"]
    kw["max_episode_steps"] = tr["meta"]["max_episode_steps"]
    return kw


def golden_traces():
    import glob
    import os

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(glob.glob(os.path.join(here, "trace_*.npz")))


def trace_id(path):
    import os

    return os.path.basename(path)[len("trace_"):-len(".npz")]


def _eq(name, a, b, t):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b):
        bad = np.argwhere(a != b) if a.shape == b.shape else None
        raise AssertionError(f"tick {t}: {name} differs; first mismatches (index): {None if bad is None else bad[:5].tolist()}\n got {a[tuple(bad[0])] if bad is not None and len(bad) else a.shape} want {b[tuple(bad[0])] if bad is not None and len(bad) else b.shape}")


def check_live(env, tr, t, get_state=True):
    _eq("obs_map", env.obs_map, tr["obs_map"][t], t)
    _eq("obs_position", env.obs_position, tr["obs_position"][t], t)
    _eq("obs_velocity", env.obs_velocity, tr["obs_velocity"][t], t)
    _eq("obs_nsd", env.obs_nsd, tr["obs_nsd"][t], t)
    if get_state:
        st = env.get_state()
        _eq("state.agent", st["agent"], tr["agent"][t], t)
        _eq("state.num_cars", st["num_cars"], tr["num_cars"][t], t)
        mc = tr["cars"].shape[2]
        _eq("state.cars", st["cars"][:, :mc], tr["cars"][t], t)
        _eq("state.tiles", st["tiles"], tr["tiles"][t], t)
        _eq("state.plan[:7]", st["plan"][:, :7], tr["plan"][t][:, :7], t)
        if "agent_direction" in tr and hasattr(env, "agent_direction"):
            _eq("agent_direction", env.agent_direction(), tr["agent_direction"][t], t)
        assert not st["error"].any(), f"tick {t}: env error flags {st['error'][st['error'] != 0][:5]}"


def replay(env, tr, get_state=True, ticks=None, from_seeds=False):
    """env must be constructed with final_observation=True and the trace's kwargs, and either
    rng_mode=RNG_TAPE (the recorded draws are loaded) or, with from_seeds=True, rng_mode=RNG_NUMPY:
    then nothing but the reference's seeds (seed + i) goes in."""
    if from_seeds:
        env.reset(seeds=tr["meta"]["seed"] + np.arange(tr["meta"]["num_envs"], dtype=np.int64))
    else:
        env.load_draws(tr["tape_values"], tr["tape_tags"], tr["tape_offsets"])
        env.reset()
    check_live(env, tr, 0, get_state)
    T = tr["actions"].shape[0] if ticks is None else ticks
    for t in range(T):
        env.step(tr["actions"][t])
        _eq("reward", env.reward, tr["reward"][t], t)
        _eq("cost", env.cost, tr["cost"][t], t)
        _eq("terminated", env.terminated, tr["terminated"][t], t)
        _eq("truncated", env.truncated, tr["truncated"][t], t)
        _eq("step_state", env.step_state, tr["step_state"][t], t)
        _eq("step_flags", env.step_flags, tr["step_flags"][t], t)
        done = (tr["terminated"][t] | tr["truncated"][t]).astype(bool)
        if done.any():
            _eq("final_obs_map", env.final_obs_map[done], tr["final_obs_map"][t][done], t)
            _eq("final_obs_position", env.final_obs_position[done], tr["final_obs_position"][t][done], t)
            _eq("final_obs_velocity", env.final_obs_velocity[done], tr["final_obs_velocity"][t][done], t)
            _eq("final_obs_nsd", env.final_obs_nsd[done], tr["final_obs_nsd"][t][done], t)
        check_live(env, tr, t + 1, get_state)
    if ticks is None and not from_seeds:
        st = env.get_state()
        _eq("draw_cursor (all recorded draws consumed)", st["draw_cursor"], tr["tape_offsets"][1:], T)
