"""GPU vs oracle in Philox mode, at the BASELINE.json configurations (sizes the oracle finishes in
seconds) and on ragged / extreme shapes. Bit-exact: integers, bytes and float64 rewards."""
import warnings

import numpy as np
import pytest

import philox_compare as pc
from oracle.oracle import OracleVectorEnv


def _pair(n, **kw):
    from native_env import NativeAdapter

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = NativeAdapter("cuda", num_envs=n, final_observation=True, **kw)
        ora = OracleVectorEnv(num_envs=n, final_observation=True, threads=8, **kw)
    return env, ora


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(pc.CONFIGS))
def test_cuda_matches_oracle_small(name):
    kw, n, ticks = pc.CONFIGS[name]
    env, ora = _pair(n, seed=4242, **kw)
    assert pc.compare(env, ora, ticks, state_every=10) > 0
    env.close(); ora.close()


@pytest.mark.gpu
def test_config2_4096_envs_fixed_map_256_ticks():
    """BASELINE config 2: 4096 envs on one fixed map, traffic off, no obstacles, 256 ticks."""
    import json, os

    tr = json.loads(bytes(np.load(os.path.join(os.path.dirname(__file__), "golden", "trace_config2_fixed_map.npz"))["meta"]).decode())
    plan = tr["maps"]["serpentine_5x3"]
    env, ora = _pair(4096, seed=7, map_plan=plan)
    assert pc.compare(env, ora, 256, state_every=64) > 1000
    env.close(); ora.close()


@pytest.mark.gpu
def test_config3_traffic_obstacles_4k():
    """BASELINE config 3 settings at 4096 envs x 48 ticks (the oracle's budget)."""
    env, ora = _pair(4096, seed=11, traffic_density=0.05, random_map_obstacle_probability=0.2)
    assert pc.compare(env, ora, 48, state_every=16, check_obs_every=4) > 1000
    env.close(); ora.close()


@pytest.mark.gpu
def test_config4_large_maps_dense_traffic():
    """BASELINE config 4 settings (8x8 tiles, 80 % connections, traffic 0.2, obstacles 0.5)."""
    env, ora = _pair(300, seed=3, random_map_width=8, random_map_height=8, random_map_percentage_of_connections=0.8,
                     traffic_density=0.2, random_map_obstacle_probability=0.5)
    assert pc.compare(env, ora, 16, state_every=8) > 0
    env.close(); ora.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 31, 129, 1000])
def test_ragged_sizes(n):
    env, ora = _pair(n, seed=5, traffic_density=0.05)
    pc.compare(env, ora, 20)
    env.close(); ora.close()


@pytest.mark.gpu
def test_default_64k_envs():
    """Default settings at 65536 envs: lock-step with the (8-thread) oracle for 12 ticks."""
    env, ora = _pair(65536, seed=99)
    assert pc.compare(env, ora, 12, check_obs_every=3) > 100000
    env.close(); ora.close()


@pytest.mark.gpu
def test_numpy_mode_config2_4096_envs_from_seeds():
    """BASELINE config 2 in numpy-exact mode: 4096 envs on the fixed map, seeds s+i only; the oracle runs
    the same numpy restatement (itself pinned against the reference's traces from seeds)."""
    import json, os

    tr = json.loads(bytes(np.load(os.path.join(os.path.dirname(__file__), "golden", "trace_config2_fixed_map.npz"))["meta"]).decode())
    env, ora = _pair(4096, rng_mode="numpy", map_plan=tr["maps"]["serpentine_5x3"])
    seeds = 1000 + np.arange(4096, dtype=np.int64)
    # both sides take the seeds through reset(seed) like the reference
    env_reset, ora_reset = env.reset, ora.reset
    env.reset = lambda: env_reset(seeds=seeds)
    ora.reset = lambda: ora_reset(seeds=seeds)
    assert pc.compare(env, ora, 128, state_every=64) > 1000
    env.close(); ora.close()


@pytest.mark.gpu
def test_numpy_mode_traffic_obstacles_procedural():
    env, ora = _pair(2048, rng_mode="numpy", traffic_density=0.05, random_map_obstacle_probability=0.3, seed=77)
    assert pc.compare(env, ora, 40, state_every=20) > 500
    env.close(); ora.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(pc.TRAFFIC_CONFIGS))
@pytest.mark.parametrize("final_obs", [True, False])
def test_traffic_tick_matches_oracle(name, final_obs):
    """The traffic tick kernel (pgtg_traffic.cu) against the oracle's sequential car loop on long-lived episodes:
    blocking chains, patience, push-through, lights, despawn / respawn bursts, occupancy counters."""
    from native_env import NativeAdapter

    kw, n, ticks, stay = pc.TRAFFIC_CONFIGS[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = NativeAdapter("cuda", num_envs=n, seed=77, final_observation=final_obs, **kw)
        ora = OracleVectorEnv(num_envs=n, seed=77, final_observation=final_obs, threads=8, **kw)
    assert "tick=traffic" in env.raw.kernel_info()
    pc.compare(env, ora, ticks, state_every=10, stay=stay, final_obs=final_obs)
    env.close(); ora.close()


@pytest.mark.gpu
def test_traffic_tick_equals_sequential_tick(monkeypatch):
    """Same handle configuration through the traffic tick and through the general (sequential) tick: bit-identical."""
    from native_env import NativeAdapter

    kw = dict(traffic_density=0.1, random_map_obstacle_probability=0.4, seed=5, final_observation=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = NativeAdapter("cuda", num_envs=3000, **kw)
        monkeypatch.setenv("PGTG_NO_TRAFFIC_KERNEL", "1")
        b = NativeAdapter("cuda", num_envs=3000, **kw)
    assert "tick=traffic" in a.raw.kernel_info() and "tick=general" in b.raw.kernel_info()
    pc.compare(a, b, 40, state_every=10, stay=0.6)
    a.close(); b.close()
