"""Forked vector runner over the UNMODIFIED reference `PGTGEnv` -- the CPU baseline BASELINE.md section 3 defines.

TEST / BENCH INFRASTRUCTURE ONLY. `gymnasium.vector.AsyncVectorEnv` is not installable here, so this is the
equivalent hand-rolled runner: one forked worker process per host core, each owning `envs_per_worker`
reference envs (the reference source from /root/reference or from the staged copy oracle/_ref, behind
oracle/shims), actions sent and (observation planes, position, velocity, reward, terminated, truncated)
returned over pipes, same-step auto-reset (gymnasium 0.28.1 semantics). This is how the reference is
consumed: pgtg/train.py:54 (SubprocVecEnv over PGTGEnv.step, environment.py:1092).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root() -> str | None:
    for cand in (os.environ.get("PGTG_REFERENCE"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "pgtg")) and os.path.exists(os.path.join(cand, "pgtg", "environment.py")):
            return cand
    return None


def _worker(conn, root, kwargs, n_envs, seed0, max_episode_steps):
    os.environ["PGTG_REFERENCE"] = root
    from oracle import ref_runner

    ref_runner.REF = root
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        environment = ref_runner.import_reference()
        envs = [environment.PGTGEnv(**kwargs) for _ in range(n_envs)]
        keys = None
        obs = []
        for i, e in enumerate(envs):
            o, _ = e.reset(seed=seed0 + i)
            obs.append(o)
        keys = list(obs[0]["map"].keys())
        elapsed = [0] * n_envs

        def pack(obs_list):
            return (np.stack([np.stack([np.asarray(o["map"][k], np.int8) for k in keys]) for o in obs_list]),
                    np.stack([np.asarray(o["position"], np.int32) for o in obs_list]),
                    np.stack([np.asarray(o["velocity"], np.int32) for o in obs_list]))

        conn.send(("ready", pack(obs)))
        while True:
            msg = conn.recv()
            if msg is None:
                break
            actions = msg
            rew = np.zeros(n_envs)
            term = np.zeros(n_envs, bool)
            trunc = np.zeros(n_envs, bool)
            resets = 0
            for i, e in enumerate(envs):
                o, r, te, tr, _ = e.step(int(actions[i]))
                elapsed[i] += 1
                tr = bool(tr) or bool(max_episode_steps and elapsed[i] >= max_episode_steps)
                if te or tr:  # same-step auto-reset: the returned observation is the new episode's first one
                    o, _ = e.reset()
                    elapsed[i] = 0
                    resets += 1
                obs[i] = o
                rew[i], term[i], trunc[i] = r, te, tr
            conn.send((pack(obs), rew, term, trunc, resets))
    conn.close()


class ReferencePool:
    def __init__(self, kwargs: dict, workers: int | None = None, envs_per_worker: int = 4, seed: int = 0, max_episode_steps: int | None = None):
        root = reference_root()
        if root is None:
            raise RuntimeError("reference tree not found (neither /root/reference nor the staged oracle/_ref)")
        self.root = root
        self.workers = workers or (os.cpu_count() or 1)
        self.envs_per_worker = envs_per_worker
        self.num_envs = self.workers * envs_per_worker
        ctx = mp.get_context("fork")
        self._conns, self._procs = [], []
        for w in range(self.workers):
            a, b = ctx.Pipe()
            pr = ctx.Process(target=_worker, args=(b, root, kwargs, envs_per_worker, seed + w * envs_per_worker, max_episode_steps), daemon=True)
            pr.start()
            b.close()
            self._conns.append(a)
            self._procs.append(pr)
        for c in self._conns:
            tag, _ = c.recv()
            assert tag == "ready"

    def step(self, actions: np.ndarray):
        k = self.envs_per_worker
        for w, c in enumerate(self._conns):
            c.send(np.asarray(actions[w * k:(w + 1) * k]))
        out = [c.recv() for c in self._conns]
        return out

    def run(self, seconds: float, action_seed: int = 0) -> dict:
        """Uniform random policy for about `seconds` of wall clock (after one warm-up step) -> throughput."""
        rng = np.random.default_rng(action_seed)
        self.step(rng.integers(0, 9, self.num_envs))
        t0, steps, resets = time.perf_counter(), 0, 0
        while time.perf_counter() - t0 < seconds:
            out = self.step(rng.integers(0, 9, self.num_envs))
            steps += 1
            resets += sum(o[4] for o in out)
        dt = time.perf_counter() - t0
        return dict(value=self.num_envs * steps / dt, env_steps=self.num_envs * steps, resets=resets, seconds=dt, workers=self.workers,
                    envs_per_worker=self.envs_per_worker)

    def close(self):
        for c in self._conns:
            try:
                c.send(None)
            except Exception:
                pass
        for p in self._procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()  # exactly the process this pool started

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
