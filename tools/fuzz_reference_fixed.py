#!/usr/bin/env python3
"""One-off hunt (build container only): random FIXED map plans (procedural plans perturbed with extra,
possibly one-sided exits and random obstacles) through the unmodified reference vs oracle / kernel logic.
    python tools/fuzz_reference_fixed.py [first] [count]"""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import parity  # noqa: E402
from native_env import NativeAdapter  # noqa: E402
from oracle import ref_runner  # noqa: E402
from oracle.oracle import OracleVectorEnv  # noqa: E402
from pgtg_b200._names import MASK_NAMES  # noqa: E402
from pgtg_b200.config import RNG_NUMPY, RNG_TAPE  # noqa: E402

first, count = (int(sys.argv[1]) if len(sys.argv) > 1 else 0), (int(sys.argv[2]) if len(sys.argv) > 2 else 30)
ref_runner.import_reference()
import map_generator  # noqa: E402  (the reference's)

bad = 0
for i in range(first, first + count):
    r = np.random.default_rng(77_000 + i)
    W, H = int(r.integers(1, 6)), int(r.integers(1, 6))
    plan = map_generator.generate_map(W, H, float(r.choice([0.3, 0.6, 1.0])), np.random.default_rng(i),
                                      start_position="random" if r.random() < 0.5 else (0, -1, "west"),
                                      goal_position="random" if r.random() < 0.5 else (-1, 0, "east"))
    for y in range(H):
        for x in range(W):
            t = plan.tiles[y][x]
            for d in range(4):
                if r.random() < 0.15:
                    t["exits"][d] = 1  # extra, possibly one-sided exit
            if r.random() < 0.4 and t["exits"] != [0, 0, 0, 0]:
                t["obstacle_type"] = str(r.choice(["ice", "broken road", "sand", "traffic_light"]))
                t["obstacle_mask"] = str(r.choice(MASK_NAMES))
    d = plan.to_dict()
    d["start"], d["goal"] = list(d["start"]), list(d["goal"])
    d["start"][:2] = [int(v) for v in d["start"][:2]]
    d["goal"][:2] = [int(v) for v in d["goal"][:2]]
    path = f"/tmp/pgtg_fuzz_fixed_{i}.json"
    json.dump(d, open(path, "w"))
    kw = dict(traffic_density=float(r.choice([0.0, 0.1, 0.4])), ignore_traffic_collisions=bool(r.random() < 0.5),
              use_next_subgoal_direction=bool(r.random() < 0.5), use_sliding_observation_window=bool(r.random() < 0.3),
              sliding_observation_window_size=int(r.integers(1, 6)),
              features_to_include_in_observation=["walls", "goals", "ice", "broken road", "sand", "traffic", "traffic_light", "start", "car_spawner"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            tr = ref_runner.record_trace({**kw, "map_path": path}, num_envs=3, ticks=40, seed=300 + i, policy="seek" if i % 2 else "random")
        except Exception as ex:
            print(i, "reference raised", type(ex).__name__, str(ex)[:80])
            continue
        tr["meta"] = json.loads(bytes(tr["meta"]).decode())
        tr["meta"]["kwargs"] = {**kw, "map_plan": "m"}
        tr["meta"]["maps"] = {"m": d}
        for name, make, seeds in (("oracle", lambda **k: OracleVectorEnv(rng_mode=RNG_TAPE, **k), False),
                                  ("emu", lambda **k: NativeAdapter("emu", rng_mode=RNG_TAPE, **k), False),
                                  ("emu-numpy", lambda **k: NativeAdapter("emu", rng_mode=RNG_NUMPY, **k), True)):
            try:
                env = make(final_observation=True, **parity.trace_kwargs(tr))
                parity.replay(env, tr, from_seeds=seeds)
            except Exception as ex:
                bad += 1
                print(i, name, "FAIL", type(ex).__name__, str(ex)[:300], f"\n   map {W}x{H} start {d['start']} goal {d['goal']} kw {kw}")
print("done; failures:", bad)
