"""Fixtures the reference's own tests hold for the hot path (tests/golden/make_reference_fixtures.py):
the golden trajectory of tests/test_integration.py:66-93 (reproducibility_data.py:5-140, its live part) and the
next_subgoal_direction table of tests/test_environment.py:606-640 on the reference's map_with_all_directions.json.
Run from SEEDS ALONE (numpy-exact mode, gymnasium seeding) on the oracle, the host emulation of the kernels and
the CUDA kernels."""
import json
import os
import warnings

import numpy as np
import pytest

from oracle.oracle import OracleVectorEnv

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BACKENDS = ["oracle", "emu", pytest.param("cuda", marks=pytest.mark.gpu)]


def _make(backend, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if backend == "oracle":
            return OracleVectorEnv(**kw)
        from native_env import NativeAdapter

        return NativeAdapter(backend, **kw)


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(bytes(d["meta"]).decode())
    return d


@pytest.mark.parametrize("backend", BACKENDS)
def test_reference_golden_trajectory_live_part(backend):
    g = _load("ref_reproducibility.npz")
    meta = g["meta"]
    assert meta["live_steps"] == 18
    keys = meta["keys"]
    env = _make(backend, num_envs=1, rng_mode="numpy", ignore_traffic_collisions=True, **meta["kwargs"])
    env.reset(seeds=np.array([meta["seed"]], np.int64))
    planes = [keys.index(k) for k in meta["live_planes"]]

    def check(t):
        assert np.array_equal(env.obs_map[0][planes], g["obs_map"][t][planes]), f"planes differ at observation {t}"
        assert np.array_equal(env.obs_position[0], g["obs_position"][t]) and np.array_equal(env.obs_velocity[0], g["obs_velocity"][t])

    check(0)
    for t in range(meta["live_steps"]):
        env.step(g["actions"][t:t + 1])
        check(t + 1)
        assert env.reward[0] == g["reward"][t] and env.terminated[0] == g["terminated"][t] and env.truncated[0] == g["truncated"][t]
    assert sorted(set(g["reward"][:18].tolist())) == [0.0, 20.0]  # three subgoals of 100 / 5 were collected on the way
    env.close()


@pytest.mark.parametrize("backend", BACKENDS)
def test_reference_next_subgoal_direction_table(backend):
    g = _load("ref_next_subgoal_direction.npz")
    plan = g["meta"]["plan"]  # the reference's tests/test_data/map_with_all_directions.json
    assert g["meta"]["map_name"] == "map_with_all_directions.json" and (plan["width"], plan["height"]) == (5, 3)
    n = len(g["tile"])
    env = _make(backend, num_envs=n, rng_mode="numpy", map_plan=plan, use_next_subgoal_direction=True)
    env.reset(seeds=np.zeros(n, np.int64))  # every row is its own PGTGEnv reset with seed 0
    agent = np.zeros((n, 4), np.int32)
    agent[:, 0] = g["tile"][:, 0] * 9 + 4  # "move the agent to the center of the tile"
    agent[:, 1] = g["tile"][:, 1] * 9 + 4
    env.set_state(agent=agent)
    env.step(np.full(n, 4, np.int32))
    got = env.final_obs_nsd if False else env.obs_nsd
    done = (env.terminated | env.truncated).astype(bool)
    assert not done.any()  # tile centres are road squares
    assert np.array_equal(got, g["live_answer"])
    same = g["file_answer"] == g["live_answer"]
    assert same.sum() == 13 and np.array_equal(got[same], g["file_answer"][same])
    env.close()
