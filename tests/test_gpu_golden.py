"""GPU parity proper: the sm_100a kernels, called through the C ABI (pgtg_create / pgtg_load_draws /
pgtg_reset / pgtg_step / pgtg_get_state / DLPack buffers), must reproduce the golden traces recorded
from the unmodified reference bit for bit -- observation planes, position, velocity, rewards
(float64), terminated / truncated, step info, terminal observations, agent and car state, map plans,
and the number of draws consumed."""
import warnings

import pytest

import parity
from pgtg_b200.config import RNG_TAPE

TRACES = parity.golden_traces()


@pytest.mark.gpu
@pytest.mark.parametrize("path", TRACES, ids=parity.trace_id)
def test_cuda_reproduces_reference(path):
    from native_env import NativeAdapter

    tr = parity.load_trace(path)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = NativeAdapter("cuda", rng_mode=RNG_TAPE, final_observation=True, **parity.trace_kwargs(tr))
    parity.replay(env, tr)
    env.close()
