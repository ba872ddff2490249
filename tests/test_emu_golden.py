"""The product's per-env logic and C-ABI host code (pgtg_b200/csrc), compiled for the host by
tests/emu (kernel phases run as loops), against the golden traces of the unmodified reference.
The same sources compiled by nvcc are checked on the GPU by tests/test_gpu_golden.py."""
import warnings

import pytest

import parity
from native_env import NativeAdapter
from pgtg_b200.config import RNG_TAPE

TRACES = parity.golden_traces()


@pytest.mark.parametrize("path", TRACES, ids=parity.trace_id)
def test_kernel_logic_reproduces_reference(path):
    tr = parity.load_trace(path)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = NativeAdapter("emu", rng_mode=RNG_TAPE, final_observation=True, **parity.trace_kwargs(tr))
    parity.replay(env, tr)
    env.close()
