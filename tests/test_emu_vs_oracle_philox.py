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
