#!/usr/bin/env python3
"""One-off hunt (build container only): random PGTGEnv constructor arguments, the UNMODIFIED reference
recorded with its draws, replayed through the oracle and the host emulation of the kernels.
    python tools/fuzz_reference.py [first] [count]"""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import parity  # noqa: E402
from fuzz_configs import random_kwargs  # noqa: E402
from native_env import NativeAdapter  # noqa: E402
from oracle import ref_runner  # noqa: E402
from oracle.oracle import OracleVectorEnv  # noqa: E402
from pgtg_b200.config import RNG_NUMPY, RNG_TAPE  # noqa: E402

first, count = (int(sys.argv[1]) if len(sys.argv) > 1 else 0), (int(sys.argv[2]) if len(sys.argv) > 2 else 40)
bad = 0
for i in range(first, first + count):
    kw = random_kwargs(i)
    mes = None if i % 3 else 9
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            tr = ref_runner.record_trace(kw, num_envs=3, ticks=40, seed=500 + i, max_episode_steps=mes, policy="seek" if i % 2 else "random")
        except Exception as ex:  # the reference itself rejects some combinations
            print(i, "reference raised", type(ex).__name__, str(ex)[:80])
            continue
        tr["meta"] = json.loads(bytes(tr["meta"]).decode())
        for name, make, seeds in (("oracle", lambda **k: OracleVectorEnv(rng_mode=RNG_TAPE, **k), False),
                                  ("emu", lambda **k: NativeAdapter("emu", rng_mode=RNG_TAPE, **k), False),
                                  ("emu-numpy", lambda **k: NativeAdapter("emu", rng_mode=RNG_NUMPY, **k), True)):
            try:
                env = make(final_observation=True, **parity.trace_kwargs(tr))
                parity.replay(env, tr, from_seeds=seeds)
            except Exception as ex:
                bad += 1
                print(i, name, "FAIL", type(ex).__name__, str(ex)[:300], "\n   kwargs:", kw)
    if i % 10 == 9:
        print("... up to", i, "failures so far:", bad, flush=True)
print("done; failures:", bad)
