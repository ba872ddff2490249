"""The reference's own unit-test answers (tests/reference_cases.py) on the oracle and on the host
emulation of the kernels (CPU), and on the CUDA kernels through the C ABI (GPU)."""
import warnings

import pytest

import reference_cases as rc


def _oracle(**kw):
    from oracle.oracle import OracleVectorEnv

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return OracleVectorEnv(num_envs=1, seed=5, **kw)


def _native(backend):
    def make(**kw):
        from native_env import NativeAdapter

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return NativeAdapter(backend, num_envs=1, seed=5, **kw)

    return make


@pytest.mark.parametrize("name,case", rc.ALL, ids=[n for n, _ in rc.ALL])
def test_oracle(name, case):
    case(_oracle)


@pytest.mark.parametrize("name,case", rc.ALL, ids=[n for n, _ in rc.ALL])
def test_kernel_logic_emulated(name, case):
    case(_native("emu"))


@pytest.mark.gpu
@pytest.mark.parametrize("name,case", rc.ALL, ids=[n for n, _ in rc.ALL])
def test_cuda(name, case):
    case(_native("cuda"))
