#!/usr/bin/env python3
"""Full-size conformance fixtures (BASELINE configs 2 and 3) -- build container only.

Runs the UNMODIFIED reference PGTGEnv (/root/reference behind oracle/shims) at the stated sizes in
forked workers and keeps per-tick / per-env digests of every output (oracle/digest.py), the actions
and -- where it is small -- the recorded draw tape:

  digest_config2   4096 envs x 256 ticks on the reference's own tests/test_data/map_with_all_directions.json,
                   traffic off, no obstacles, env i seeded 0 + i, actions torch.randint(0, 9, (256, 4096),
                   generator=manual_seed(0)); tape = the start-square draws.
  digest_config3   4096 envs x 64 ticks, default 4x4 procedural maps, traffic 0.05, obstacles 0.2; no tape (20 M
                   draws): replayed from seeds alone in the numpy-exact mode.
  digest_trainpy   512 envs x 100 ticks of the reference's consumer configuration (pgtg/train.py:21-38) with
                   TimeLimit(100); from seeds alone.

    python tests/golden/make_digests.py [name ...]
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("PGTG_REFERENCE", "/root/reference")

JOBS = {
    "config2": dict(kwargs=dict(), map_json="tests/test_data/map_with_all_directions.json", n=4096, ticks=256, seed=0, tape=True, mes=None),
    "config3": dict(kwargs=dict(traffic_density=0.05, random_map_obstacle_probability=0.2), map_json=None, n=4096, ticks=64, seed=0, tape=False, mes=None),
    "trainpy": dict(kwargs=None, map_json=None, n=512, ticks=100, seed=0, tape=False, mes=100),
}


def train_py_kwargs():
    """The keyword arguments of pgtg/train.py:21-38, read from the reference file itself."""
    import ast

    src = open(os.path.join(REF, "pgtg", "train.py")).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.Call) and getattr(node.func, "id", getattr(node.func, "attr", "")) == "PGTGEnv":
            return {k.arg: ast.literal_eval(k.value) for k in node.keywords}
    raise RuntimeError("PGTGEnv(...) call not found in train.py")


def actions_for(ticks, n):
    import torch

    g = torch.Generator()
    g.manual_seed(0)
    return torch.randint(0, 9, (ticks, n), generator=g).numpy().astype(np.int32)


def main(argv):
    from oracle import ref_runner

    for name in argv or list(JOBS):
        job = JOBS[name]
        kw = dict(job["kwargs"]) if job["kwargs"] is not None else train_py_kwargs()
        ref_kw, maps = dict(kw), {}
        if job["map_json"]:
            path = os.path.join(REF, job["map_json"])
            maps[os.path.basename(path)] = json.load(open(path))
            ref_kw["map_path"] = path
            kw["map_plan"] = os.path.basename(path)
        acts = actions_for(job["ticks"], job["n"])
        t0 = time.time()
        out = ref_runner.record_digests(ref_kw, job["n"], job["ticks"], job["seed"], acts, max_episode_steps=job["mes"],
                                        workers=int(os.environ.get("WORKERS", os.cpu_count() or 1)), keep_tape=job["tape"])
        meta = dict(kwargs=kw, maps=maps, num_envs=job["n"], ticks=job["ticks"], seed=job["seed"], max_episode_steps=job["mes"],
                    totals=out["totals"], numpy=np.__version__, actions="torch.randint(0, 9, (ticks, n), generator=manual_seed(0))",
                    seconds=round(time.time() - t0, 1))
        arrays = dict(tick_digest=out["tick_digest"], env_digest=out["env_digest"],
                      actions_packed=(acts[:, 0::2] | acts[:, 1::2] << 4).astype(np.uint8),
                      meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
        if job["tape"]:
            v = out["tape_values"]
            assert np.all(v == np.floor(v)) and v.min() >= 0 and v.max() < 256, "config-2 tapes hold small indices only"
            arrays.update(tape_values_u8=v.astype(np.uint8), tape_tags=out["tape_tags"], tape_offsets=out["tape_offsets"])
        path = os.path.join(HERE, f"digest_{name}.npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: {time.time() - t0:.0f}s {out['totals']} -> {os.path.getsize(path) // 1024} KiB", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
