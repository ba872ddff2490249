"""Philox mode on the CPU: the product's kernel logic (host emulation build) against the oracle,
over configurations and sizes the recorded reference traces do not reach (ragged CTA tails,
16x16 maps, dense traffic, 1x1 maps)."""
import warnings

import pytest

import philox_compare as pc
from native_env import NativeAdapter
from oracle.oracle import OracleVectorEnv


@pytest.mark.parametrize("name", list(pc.CONFIGS))
def test_emulation_matches_oracle(name):
    kw, n, ticks = pc.CONFIGS[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = NativeAdapter("emu", num_envs=n, seed=4242, final_observation=True, **kw)
        ora = OracleVectorEnv(num_envs=n, seed=4242, final_observation=True, **kw)
    episodes = pc.compare(env, ora, ticks, state_every=10)
    assert episodes > 0
    env.close()
    ora.close()


@pytest.mark.parametrize("name", list(pc.TRAFFIC_CONFIGS))
@pytest.mark.parametrize("final_obs", [True, False])
def test_traffic_tick_matches_oracle(name, final_obs):
    """The traffic tick (pgtg_traffic.cuh: flat car phases + per-env ordered pass) against the oracle's sequential car loop."""
    kw, n, ticks, stay = pc.TRAFFIC_CONFIGS[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = NativeAdapter("emu", num_envs=n, seed=77, final_observation=final_obs, **kw)
        ora = OracleVectorEnv(num_envs=n, seed=77, final_observation=final_obs, **kw)
    pc.compare(env, ora, ticks, state_every=10, stay=stay, final_obs=final_obs)
    env.close()
    ora.close()


@pytest.mark.parametrize("name", ["crossing_full", "ring_dense_lights", "aggressive_push", "big_8x8_idle"])
def test_traffic_tick_counter_saturation_fallback(name):
    """The same against a build whose per-square counters saturate at 3 cars: squares with piled-up cars take the
    exact-count fallback of the ordered pass and of the agent's collision test."""
    kw, n, ticks, stay = pc.TRAFFIC_CONFIGS[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = NativeAdapter("emu_sat3", num_envs=n, seed=78, final_observation=True, **kw)
        ora = OracleVectorEnv(num_envs=n, seed=78, final_observation=True, **kw)
    pc.compare(env, ora, ticks, state_every=10, stay=stay)
    env.close()
    ora.close()


@pytest.mark.parametrize("name", ["carfree_sliding_nsd", "carfree_nsd_penalties"])
def test_carfree_configurations_on_the_traffic_tick(name, monkeypatch):
    """Car-free sliding-window / next_subgoal_direction configurations normally run the general tick; the traffic tick's
    parallel observation phases must give the same bits when forced onto them."""
    monkeypatch.setenv("PGTG_TRAFFIC_KERNEL_CARFREE", "1")
    kw, n, ticks = pc.CONFIGS[name]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = NativeAdapter("emu", num_envs=n, seed=4242, final_observation=True, **kw)
        ora = OracleVectorEnv(num_envs=n, seed=4242, final_observation=True, **kw)
    assert "tick=traffic" in env.raw.kernel_info()
    assert pc.compare(env, ora, ticks, state_every=10) > 0
    env.close()
    ora.close()
