"""The CPU oracle (oracle/pgtg_oracle.c) against the golden traces recorded from the unmodified
reference: every observation plane, reward, flag, agent / car / map state at every tick, and the
exact number of random draws consumed."""
import warnings

import numpy as np
import pytest

import parity
from oracle.oracle import OracleVectorEnv, decompose_velocity
from pgtg_b200.config import RNG_TAPE

TRACES = parity.golden_traces()


def test_golden_traces_present():
    assert len(TRACES) >= 20


@pytest.mark.parametrize("path", TRACES, ids=parity.trace_id)
def test_oracle_reproduces_reference(path):
    tr = parity.load_trace(path)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = OracleVectorEnv(rng_mode=RNG_TAPE, final_observation=True, **parity.trace_kwargs(tr))
    parity.replay(env, tr)
    env.close()


def test_harness_detects_a_wrong_draw():
    """Sanity of the checker itself: flipping one recorded draw must break parity."""
    tr = parity.load_trace([p for p in TRACES if p.endswith("trace_obstacles.npz")][0])
    vals = tr["tape_values"].copy()
    idx = np.nonzero(tr["tape_tags"] == 0 * 8 + 1)[0]  # an index draw of the map stream
    vals[idx[3]] = (vals[idx[3]] + 1) % 2
    tr["tape_values"] = vals
    env = OracleVectorEnv(rng_mode=RNG_TAPE, final_observation=True, **parity.trace_kwargs(tr))
    with pytest.raises(AssertionError):
        parity.replay(env, tr)


# _decompose_velocity known answers: the five cases of the reference's own unit tests
# (tests/test_environment.py:1127-1153), then a few more of the same shape
DECOMPOSE_KAT = [
    ((3, 0), [(1, 0), (1, 0), (1, 0)]),            # test_decompose_velocity_easy
    ((0, 3), [(0, 1), (0, 1), (0, 1)]),
    ((3, -3), [(1, -1), (1, -1), (1, -1)]),        # test_decompose_velocity_complex
    ((3, 1), [(1, 0), (1, 1), (1, 0)]),
    ((-1, -3), [(0, -1), (-1, -1), (0, -1)]),
    ((0, 0), []),
    ((4, 2), [(1, 1), (1, 0), (1, 1), (1, 0)]),
]


@pytest.mark.parametrize("vel,want", DECOMPOSE_KAT)
def test_decompose_velocity(vel, want):
    got = [tuple(int(v) for v in row) for row in decompose_velocity(*vel)]
    assert got == want


def test_decompose_velocity_matches_float64_formula():
    """The reference rounds floor(i * (dminor / |dmajor|) + 0.5) in float64 (environment.py:29-30,
    725-738); numpy float64 arithmetic is the same IEEE arithmetic, so compare exhaustively."""
    for dx in range(-40, 41):
        for dy in range(-40, 41):
            got = decompose_velocity(dx, dy)
            n = max(abs(dx), abs(dy))
            pts = []
            for i in range(1, n + 1):
                if dx == 0:
                    pts.append((0, i * int(np.sign(dy))))
                elif dy == 0:
                    pts.append((i * int(np.sign(dx)), 0))
                elif abs(dx) >= abs(dy):
                    m = np.float64(dy) / np.float64(abs(dx))
                    pts.append((i * int(np.sign(dx)), int(np.floor(i * m + 0.5))))
                else:
                    m = np.float64(dx) / np.float64(abs(dy))
                    pts.append((int(np.floor(i * m + 0.5)), i * int(np.sign(dy))))
            want = np.diff(np.array([(0, 0)] + pts), axis=0) if pts else np.zeros((0, 2))
            assert np.array_equal(got, want), (dx, dy)
