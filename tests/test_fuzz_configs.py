"""Config fuzzing: deterministic pseudo-random PGTGEnv constructor arguments (tests/fuzz_configs.py --
map shapes from 1x1 to 6x6, start/goal modes, obstacle mixes, traffic and driver mixes, sliding windows,
literal feature planes, penalties), kernel logic vs oracle in Philox mode. The same configurations were
replayed against the unmodified reference with tools/fuzz_reference.py (260 configurations; that hunt
found the two 1-wide-map bugs now pinned by golden traces narrow_*)."""
import warnings

import pytest

import philox_compare as pc
from fuzz_configs import random_kwargs
from oracle.oracle import OracleVectorEnv

CASES = list(range(48))


def _make(backend, i, n):
    from native_env import NativeAdapter

    kw = random_kwargs(i)
    mes = None if i % 3 else 9
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = NativeAdapter(backend, num_envs=n, seed=900 + i, final_observation=True, max_episode_steps=mes, **kw)
        ora = OracleVectorEnv(num_envs=n, seed=900 + i, final_observation=True, max_episode_steps=mes, threads=4, **kw)
    return env, ora


@pytest.mark.parametrize("i", CASES)
def test_emulation_matches_oracle_on_random_config(i):
    env, ora = _make("emu", i, 24)
    pc.compare(env, ora, 20, state_every=10)
    env.close(); ora.close()


@pytest.mark.gpu
@pytest.mark.parametrize("i", CASES)
def test_cuda_matches_oracle_on_random_config(i):
    env, ora = _make("cuda", i, 200)
    pc.compare(env, ora, 16, state_every=8)
    env.close(); ora.close()
