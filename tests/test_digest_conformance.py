"""Conformance at the BASELINE sizes against the LIVE reference: the unmodified reference was run
at 4096 envs x 256 ticks (config 2, on its own tests/test_data/map_with_all_directions.json), 4096 x 64 (config 3:
traffic + obstacles) and 512 x 100 (the consumer configuration of pgtg/train.py:21-38 with TimeLimit 100) by
tests/golden/make_digests.py, which kept a digest of EVERY output of every env at every tick (oracle/digest.py).
Here the same digests are recomputed from the oracle (CPU) and from the CUDA buffers (GPU): in conformance-tape
mode where the tape is small (config 2), and from seeds alone (numpy-exact mode) everywhere."""
import json
import os
import warnings

import numpy as np
import pytest

from oracle import digest
from oracle.oracle import OracleVectorEnv

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BACKENDS = ["oracle", pytest.param("cuda", marks=pytest.mark.gpu)]


def _load(name):
    path = os.path.join(GOLDEN, f"digest_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(bytes(d["meta"]).decode())
    a = d["actions_packed"]
    acts = np.empty((a.shape[0], a.shape[1] * 2), np.int32)
    acts[:, 0::2], acts[:, 1::2] = a & 15, a >> 4
    d["actions"] = acts[:, : d["meta"]["num_envs"]]
    return d


def _make(backend, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if backend == "oracle":
            return OracleVectorEnv(threads=os.cpu_count() or 1, **kw)
        from native_env import NativeAdapter

        return NativeAdapter(backend, **kw)


def _run(backend, d, mode):
    meta = d["meta"]
    kw = dict(meta["kwargs"])
    if "map_plan" in kw:
        kw["map_plan"] = meta["maps"][kw["map_plan"]]
    N, T = meta["num_envs"], meta["ticks"]
    env = _make(backend, num_envs=N, rng_mode=mode, final_observation=True, max_episode_steps=meta["max_episode_steps"], **kw)
    if mode == "tape":
        env.load_draws(d["tape_values_u8"].astype(np.float64), d["tape_tags"], d["tape_offsets"])
        env.reset()
    else:
        env.reset(seeds=meta["seed"] + np.arange(N, dtype=np.int64))
    st = env.get_state()
    dg = digest.LiveDigester(env, N, env.C, env.P, st["tiles"].shape[1], st["cars"].shape[1])
    dg.after_reset()
    for t in range(T):
        env.step(d["actions"][t])
        dg.after_step()
    tick, envd = dg.digests()
    bad_t = np.flatnonzero((tick != d["tick_digest"]).any(axis=1))
    bad_e = np.flatnonzero((envd != d["env_digest"]).any(axis=1))
    assert len(bad_t) == 0 and len(bad_e) == 0, f"first diverging tick {bad_t[:1]}, envs {bad_e[:8]} ({len(bad_e)} of {N})"
    if mode == "tape":
        assert np.array_equal(env.get_state()["draw_cursor"], d["tape_offsets"][1:]), "recorded draws left over / overrun"
    assert not env.get_state()["error"].any()
    env.close()
    return meta


@pytest.mark.parametrize("mode", ["tape", "numpy"])
@pytest.mark.parametrize("backend", BACKENDS)
def test_config2_4096_envs_256_ticks_on_the_reference_map(backend, mode):
    d = _load("config2")
    meta = d["meta"]
    assert meta["num_envs"] == 4096 and meta["ticks"] == 256 and "map_with_all_directions.json" in meta["maps"]
    _run(backend, d, mode)
    assert meta["totals"]["done"] > 100000  # random policy: ~40 % of the env-ticks end an episode


@pytest.mark.parametrize("backend", BACKENDS)
def test_config3_4096_envs_64_ticks_from_seeds(backend):
    d = _load("config3")
    meta = d["meta"]
    assert meta["num_envs"] == 4096 and meta["ticks"] == 64 and meta["kwargs"]["traffic_density"] == 0.05
    _run(backend, d, "numpy")
    assert meta["totals"]["cars_max"] > 10


@pytest.mark.parametrize("backend", BACKENDS)
def test_train_py_configuration_from_seeds(backend):
    d = _load("trainpy")
    meta = d["meta"]
    assert meta["kwargs"]["sliding_observation_window_size"] == 5 and meta["max_episode_steps"] == 100
    _run(backend, d, "numpy")
