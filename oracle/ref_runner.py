"""Run the UNMODIFIED reference `PGTGEnv` (from /root/reference) and record golden traces.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: /root/reference does not exist on
the GPU box, so nothing under tests/ -m gpu, smoke() or bench.py imports this at run time. The
traces it writes are committed under tests/golden/ together with tests/golden/make_golden.py.

How: four stand-in modules in oracle/shims/ (gymnasium, pygame,
graphic, graph) are put ahead of /root/reference/pgtg on sys.path; gymnasium.RNG_FACTORY wraps the
env's parent generator so that the five children spawned at environment.py:593-599 are recording
proxies. Every draw is logged as a *semantic* value: the double for random(), the index for
choice()/integers() (nothing for a 1-element population: numpy consumes no bits there).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# the reference tree: the read-only checkout in the build container, else the copy staged by oracle/stage_ref.py
REF = os.environ.get("PGTG_REFERENCE") or ("/root/reference" if os.path.isdir("/root/reference/pgtg") else os.path.join(_HERE, "_ref"))


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF, "pgtg"))


def import_reference():
    """-> the reference `environment` module (unmodified source, run behind the shims)."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected in the build container only)")
    for p in (REF, os.path.join(REF, "pgtg"), os.path.join(_HERE, "shims")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    # `parser` must be the reference's pgtg/parser.py, not a stdlib leftover
    for name in ("parser", "map", "graph", "graphic", "constants", "map_generator", "map_tiles_data"):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith((REF, _HERE)):
            del sys.modules[name]
    import environment  # noqa: E402

    return environment


class Tape:
    def __init__(self):
        self.values: list[float] = []
        self.tags: list[int] = []

    def log(self, stream: int, kind: int, value: float):
        self.values.append(float(value))
        self.tags.append(stream * 8 + kind)


class RecordingGenerator:
    """Forwards to the real numpy Generator and logs semantic draws (stream id fixed per child)."""

    def __init__(self, gen, stream: int, tape: Tape):
        self._g, self._s, self._t = gen, stream, tape

    def random(self):
        v = self._g.random()
        self._t.log(self._s, 0, v)
        return v

    def integers(self, low, high=None):
        if high is None:
            low, high = 0, low
        v = int(self._g.integers(low, high))
        if high - low > 1:
            self._t.log(self._s, 1, v - low)
        return v

    def choice(self, a, size=None, replace=True, p=None):
        items = None
        if isinstance(a, (int, np.integer)):
            n = int(a)
        else:
            items = list(a)
            n = len(items)
        idx = self._g.choice(n, size=size, replace=replace, p=p)  # same bit consumption as choice(a)
        if size is None:
            if n > 1 or p is not None:
                self._t.log(self._s, 1, int(idx))
            return int(idx) if items is None else items[int(idx)]
        idx = np.asarray(idx)
        for v in idx.ravel():
            self._t.log(self._s, 1, int(v))
        return idx if items is None else [items[int(v)] for v in idx.ravel()]


class RecordingParent:
    def __init__(self, gen, tape: Tape):
        self._g, self._t = gen, tape

    def spawn(self, n):
        return [RecordingGenerator(g, i, self._t) for i, g in enumerate(self._g.spawn(n))]

    def __getattr__(self, name):
        return getattr(self._g, name)


def _obs_arrays(obs, keys, P):
    m = np.zeros((len(keys), P, P), np.int8)
    for i, k in enumerate(keys):
        m[i] = np.asarray(obs["map"][k], dtype=np.int8)
    nsd = int(obs.get("next_subgoal_direction", -1))
    return m, np.asarray(obs["position"], np.int32), np.asarray(obs["velocity"], np.int32), nsd


def record_trace(kwargs: dict, num_envs: int, ticks: int, seed: int, action_seed: int = 0,
                 max_episode_steps: int | None = None, actions: np.ndarray | None = None,
                 policy: str = "random", epsilon: float = 0.3) -> dict:
    """Roll `num_envs` independent reference envs for `ticks` ticks with same-step auto-reset
    (gymnasium 0.28.1 vector semantics) and return every output as arrays [ticks(+1), N, ...]."""
    environment = import_reference()
    import gymnasium
    from pgtg_b200.config import PROFILE_NAMES, make_config
    from pgtg_b200._names import ROUTE_NAMES

    hc = make_config(num_envs=num_envs, max_episode_steps=max_episode_steps, **kwargs)
    keys, P, Cn = hc.observation_keys, hc.window, len(hc.observation_keys)
    MC = max(1, hc.pod.max_cars)
    T = hc.pod.map_w * hc.pod.map_h
    arng = np.random.default_rng(action_seed)
    if actions is None:
        actions = arng.integers(0, 9, size=(ticks, num_envs)).astype(np.int32)
    actions = np.array(actions, np.int32)

    def seek_action(env):
        """Episode-lengthening policy for richer traces: steer at walking speed toward the nearest
        remaining subgoal / final-goal square (uses env internals; the actions are recorded)."""
        best, tgt = None, None
        for x in range(env.map.width):
            for y in range(env.map.height):
                if env.map.feature_at(x, y, "subgoal") or env.map.feature_at(x, y, "final goal"):
                    d = abs(x - env.position[0]) + abs(y - env.position[1])
                    if best is None or d < best:
                        best, tgt = d, (x, y)
        want = (int(np.sign(tgt[0] - env.position[0])), int(np.sign(tgt[1] - env.position[1])))
        ax = int(np.clip(want[0] - env.velocity[0], -1, 1))
        ay = int(np.clip(want[1] - env.velocity[1], -1, 1))
        return (ax + 1) * 3 + (ay + 1)
    N = num_envs
    out = dict(
        actions=actions,  # filled in as played when policy != random
        obs_map=np.zeros((ticks + 1, N, Cn, P, P), np.int8),
        obs_position=np.zeros((ticks + 1, N, 2), np.int32),
        obs_velocity=np.zeros((ticks + 1, N, 2), np.int32),
        obs_nsd=np.full((ticks + 1, N), -1, np.int32),
        reward=np.zeros((ticks, N), np.float64),
        cost=np.zeros((ticks, N), np.float64),
        terminated=np.zeros((ticks, N), np.uint8),
        truncated=np.zeros((ticks, N), np.uint8),
        step_state=np.zeros((ticks, N, 4), np.int32),
        step_flags=np.zeros((ticks, N), np.uint8),
        final_obs_map=np.zeros((ticks, N, Cn, P, P), np.int8),
        final_obs_position=np.zeros((ticks, N, 2), np.int32),
        final_obs_velocity=np.zeros((ticks, N, 2), np.int32),
        final_obs_nsd=np.full((ticks, N), -1, np.int32),
        agent=np.zeros((ticks + 1, N, 4), np.int32),
        num_cars=np.zeros((ticks + 1, N), np.int32),
        cars=np.zeros((ticks + 1, N, MC, 7), np.int32),
        tiles=np.zeros((ticks + 1, N, T), np.uint16),
        plan=np.zeros((ticks + 1, N, 8), np.int32),
        agent_direction=np.zeros((ticks + 1, N), np.int32),
    )
    from pgtg_b200.config import AGENT_DIRECTIONS
    from pgtg_b200._names import CARDINALS, MASK_NAMES, OBSTACLE_NAMES

    def dump_state(env, t, i):
        out["agent"][t, i] = [env.position[0], env.position[1], env.velocity[0], env.velocity[1]]
        out["num_cars"][t, i] = len(env.cars)
        for k, c in enumerate(env.cars):
            out["cars"][t, i, k] = [c.id, c.position.x, c.position.y, ROUTE_NAMES.index(c.route),
                                    PROFILE_NAMES.index(c.driver_profile.value), c.patience_counter, c.last_action_delay]
        mp = env.map_plan
        dirs = env.map.tile_coordinates_to_subgoal_directions
        for y in range(mp.height):
            for x in range(mp.width):
                tl = mp.tiles[y][x]
                ex = tl["exits"]
                v = int(ex[0]) | int(ex[1]) << 1 | int(ex[2]) << 2 | int(ex[3]) << 3
                if tl.get("obstacle_type") is not None:
                    v |= (1 + OBSTACLE_NAMES.index(tl["obstacle_type"])) << 4
                    v |= MASK_NAMES.index(tl["obstacle_mask"]) << 7
                if (x, y) in dirs:
                    v |= (1 + CARDINALS.index(dirs[(x, y)])) << 11
                out["tiles"][t, i, y * mp.width + x] = v
        out["plan"][t, i] = [mp.start[0], mp.start[1], CARDINALS.index(mp.start[2]), mp.goal[0], mp.goal[1],
                             CARDINALS.index(mp.goal[2]), env.map.num_subgoals, 0]
        out["agent_direction"][t, i] = AGENT_DIRECTIONS.index(env.get_agent_direction_string())

    ref_kwargs = dict(kwargs)
    tapes = []
    for i in range(N):
        tape = Tape()
        tapes.append(tape)
        gymnasium.RNG_FACTORY = lambda g, tape=tape: RecordingParent(g, tape)
        env = environment.PGTGEnv(**ref_kwargs)
        obs, _ = env.reset(seed=seed + i)
        m, p, v, d = _obs_arrays(obs, keys, P)
        out["obs_map"][0, i], out["obs_position"][0, i], out["obs_velocity"][0, i], out["obs_nsd"][0, i] = m, p, v, d
        dump_state(env, 0, i)
        elapsed = 0
        for t in range(ticks):
            if policy == "seek" and arng.random() >= epsilon:
                actions[t, i] = seek_action(env)
            obs, rew, term, trunc, info = env.step(int(actions[t, i]))
            elapsed += 1
            trunc = bool(max_episode_steps) and elapsed >= max_episode_steps
            out["reward"][t, i] = rew
            out["cost"][t, i] = info.get("cost", 0) if kwargs.get("separate_reward_cost") else 0
            out["terminated"][t, i] = term
            out["truncated"][t, i] = trunc
            out["step_state"][t, i] = [info["x"], info["y"], info["x_velocity"], info["y_velocity"]]
            out["step_flags"][t, i] = (1 if info["flat_tire"] else 0) | (2 if info["traffic_rules"]["braking_applied"] else 0)
            if term or trunc:
                m, p, v, d = _obs_arrays(obs, keys, P)
                out["final_obs_map"][t, i], out["final_obs_position"][t, i], out["final_obs_velocity"][t, i], out["final_obs_nsd"][t, i] = m, p, v, d
                obs, _ = env.reset()
                elapsed = 0
            m, p, v, d = _obs_arrays(obs, keys, P)
            out["obs_map"][t + 1, i], out["obs_position"][t + 1, i], out["obs_velocity"][t + 1, i], out["obs_nsd"][t + 1, i] = m, p, v, d
            dump_state(env, t + 1, i)
    gymnasium.RNG_FACTORY = None
    offsets = np.zeros(N + 1, np.int64)
    for i, tp in enumerate(tapes):
        offsets[i + 1] = offsets[i] + len(tp.values)
    out["tape_values"] = np.concatenate([np.asarray(tp.values, np.float64) for tp in tapes]) if offsets[-1] else np.zeros(0)
    out["tape_tags"] = np.concatenate([np.asarray(tp.tags, np.uint8) for tp in tapes]) if offsets[-1] else np.zeros(0, np.uint8)
    out["tape_offsets"] = offsets
    meta = dict(kwargs=kwargs, num_envs=N, ticks=ticks, seed=seed, max_episode_steps=max_episode_steps,
                observation_keys=keys, numpy=np.__version__)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


# ---- digests of a full-size reference run (see oracle/digest.py) -----------------------------------
def _digest_chunk(args):
    kwargs, start, n, ticks, seed, actions, max_episode_steps, keep_tape = args
    import warnings

    from oracle import digest

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        tr = record_trace(kwargs, num_envs=n, ticks=ticks, seed=seed + start, max_episode_steps=max_episode_steps, actions=actions)
    Cn, P = tr["obs_map"].shape[2], tr["obs_map"].shape[3]
    T, MC = tr["tiles"].shape[2], tr["cars"].shape[2]
    h = np.zeros((n, ticks + 1, 16), np.uint8)

    def obs(t):
        return dict(map=tr["obs_map"][t], position=tr["obs_position"][t], velocity=tr["obs_velocity"][t], nsd=tr["obs_nsd"][t])

    def state(t):
        return dict(agent=tr["agent"][t], num_cars=tr["num_cars"][t], cars=tr["cars"][t], tiles=tr["tiles"][t], plan=tr["plan"][t],
                    agent_direction=tr["agent_direction"][t])

    h[:, 0] = digest.row_hashes(digest.rows(n, Cn, P, T, MC, obs=obs(0), state=state(0)))
    for t in range(ticks):
        step = {k: tr[k][t] for k in ("reward", "cost", "terminated", "truncated", "step_state", "step_flags")}
        final = dict(map=tr["final_obs_map"][t], position=tr["final_obs_position"][t], velocity=tr["final_obs_velocity"][t], nsd=tr["final_obs_nsd"][t])
        h[:, t + 1] = digest.row_hashes(digest.rows(n, Cn, P, T, MC, step=step, final=final, obs=obs(t + 1), state=state(t + 1)))
    tape = (tr["tape_values"], tr["tape_tags"], np.diff(tr["tape_offsets"])) if keep_tape else None
    counts = dict(done=int(tr["terminated"].sum() + tr["truncated"].sum()), reward_pos=int((tr["reward"] > 0).sum()),
                  brake=int(((tr["step_flags"] & 2) > 0).sum()), draws=int(tr["tape_offsets"][-1]), cars_max=int(tr["num_cars"].max()))
    return start, h, tape, counts


def record_digests(kwargs: dict, num_envs: int, ticks: int, seed: int, actions: np.ndarray, max_episode_steps: int | None = None,
                   workers: int = 8, chunk: int = 16, keep_tape: bool = True) -> dict:
    """Run `num_envs` reference envs (seed + i) for `ticks` ticks of `actions` [ticks, num_envs] with same-step
    auto-reset, in `workers` forked processes, and keep only the digests (+ optionally the draw tape)."""
    import multiprocessing as mp

    jobs = [(kwargs, s, min(chunk, num_envs - s), ticks, seed, np.ascontiguousarray(actions[:, s:s + chunk]), max_episode_steps, keep_tape)
            for s in range(0, num_envs, chunk)]
    h = np.zeros((num_envs, ticks + 1, 16), np.uint8)
    tapes, totals = {}, {}
    with mp.get_context("fork").Pool(workers) as pool:
        for start, hc, tape, counts in pool.imap_unordered(_digest_chunk, jobs):
            h[start:start + hc.shape[0]] = hc
            tapes[start] = tape
            for k, v in counts.items():
                totals[k] = max(totals.get(k, 0), v) if k == "cars_max" else totals.get(k, 0) + v
    from oracle import digest

    tick_digest, env_digest = digest.fold(h)
    out = dict(tick_digest=tick_digest, env_digest=env_digest, totals=totals)
    if keep_tape:
        order = sorted(tapes)
        out["tape_values"] = np.concatenate([tapes[s][0] for s in order])
        out["tape_tags"] = np.concatenate([tapes[s][1] for s in order])
        out["tape_offsets"] = np.concatenate([[0], np.cumsum(np.concatenate([tapes[s][2] for s in order]))]).astype(np.int64)
    return out
