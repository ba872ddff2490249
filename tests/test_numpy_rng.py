"""numpy-exact RNG mode: with rng_mode="numpy" the kernels restate numpy's
SeedSequence / PCG64 / Generator.random / integers / choice, so the golden traces of the unmodified
reference are reproduced from the SEEDS ALONE -- no recorded draws go in."""
import warnings

import numpy as np
import pytest

import parity
from pgtg_b200.config import RNG_NUMPY

TRACES = parity.golden_traces()


@pytest.mark.parametrize("path", TRACES, ids=parity.trace_id)
def test_kernel_logic_reproduces_reference_from_seeds(path):
    from native_env import NativeAdapter

    tr = parity.load_trace(path)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = NativeAdapter("emu", rng_mode=RNG_NUMPY, final_observation=True, **parity.trace_kwargs(tr))
    parity.replay(env, tr, from_seeds=True)
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", TRACES, ids=parity.trace_id)
def test_cuda_reproduces_reference_from_seeds(path):
    from native_env import NativeAdapter

    tr = parity.load_trace(path)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = NativeAdapter("cuda", rng_mode=RNG_NUMPY, final_observation=True, **parity.trace_kwargs(tr))
    parity.replay(env, tr, from_seeds=True)
    env.close()


def test_first_map_matches_numpy_driven_generation():
    """Independent of the traces: the default map of seed s equals what the recorded tape of the same
    seed produced (map_rng = first child of SeedSequence(s))."""
    from native_env import NativeAdapter

    tr = parity.load_trace([p for p in TRACES if p.endswith("trace_default.npz")][0])
    n = tr["meta"]["num_envs"]
    env = NativeAdapter("emu", num_envs=n, rng_mode=RNG_NUMPY)
    env.reset(seeds=tr["meta"]["seed"] + np.arange(n, dtype=np.int64))
    st = env.get_state()
    assert np.array_equal(st["tiles"], tr["tiles"][0]) and np.array_equal(st["agent"], tr["agent"][0])
    env.close()


@pytest.mark.parametrize("path", TRACES[:8], ids=parity.trace_id)
def test_oracle_reproduces_reference_from_seeds(path):
    from oracle.oracle import OracleVectorEnv

    tr = parity.load_trace(path)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = OracleVectorEnv(rng_mode=RNG_NUMPY, final_observation=True, **parity.trace_kwargs(tr))
    parity.replay(env, tr, from_seeds=True)
    env.close()
