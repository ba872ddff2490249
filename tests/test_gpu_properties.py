"""Size-independent properties at the BASELINE.json sizes (where lock-step comparison with the
oracle would take too long): determinism, structural invariants of every observation, reward
support, statistics bookkeeping, invariance to the CTA decomposition."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rollout(n, ticks, seed, device="cuda:0", **kw):
    import torch

    from pgtg_b200 import PGTGVectorEnv

    env = PGTGVectorEnv(n, device=device, seed=seed, **kw)
    env.reset()
    g = torch.Generator(device=device)
    g.manual_seed(7)
    digest = torch.zeros((), dtype=torch.int64, device=device)
    w = torch.arange(1, 730, device=device, dtype=torch.int64)
    rsum = torch.zeros((), dtype=torch.float64, device=device)
    done = 0
    for t in range(ticks):
        a = torch.randint(0, 9, (n,), device=device, dtype=torch.int32, generator=g)
        obs, rew, term, trunc, info = env.step(a)
        m = env._t["obs_map"]
        flat = m.reshape(n, -1).to(torch.int64)
        if flat.shape[1] == 729:
            digest = digest * 1000003 + (flat * w).sum()
        else:
            digest = digest * 1000003 + flat.sum()
        rsum += rew.sum()
        done += int((term | trunc).sum())
        if t % 8 == 0:
            assert int(m.min()) == 0 and int(m.max()) == 1
            assert bool(((obs["position"] >= 0) & (obs["position"] < 9)).all())
            # a fresh or running env always sees walls, and never more goal squares than one 3-square line
            walls = obs["map"]["walls"].reshape(n, -1).sum(1)
            goals = obs["map"]["goals"].reshape(n, -1).sum(1)
            assert int(walls.min()) >= 30 and int(goals.max()) <= 6
            assert bool((obs["velocity"][term | trunc] == 0).all())
            support = torch.unique(rew)
            assert all(-100 <= v <= 100 for v in support.tolist())  # a tick can collect subgoal rewards and still crash
    stats = env.episode_stats()
    assert stats["episodes"] == done
    err = env.get_state()["error"]
    assert not err.any()
    env.close()
    return int(digest.item()), float(rsum.item()), done


def test_default_65536_envs_deterministic():
    a = _rollout(65536, 24, seed=11)
    b = _rollout(65536, 24, seed=11)
    c = _rollout(65536, 24, seed=12)
    assert a == b and a != c and a[2] > 400000


def test_config3_65536_envs():
    """BASELINE config 3 at full size: traffic 0.05, obstacles 0.2."""
    a = _rollout(65536, 16, seed=5, traffic_density=0.05, random_map_obstacle_probability=0.2)
    b = _rollout(65536, 16, seed=5, traffic_density=0.05, random_map_obstacle_probability=0.2)
    assert a == b


def test_config4_large_maps_dense_traffic_32k():
    """BASELINE config 4 settings (8x8 tiles, 80 % connections, traffic 0.2, obstacles 0.5) at 32768 envs."""
    a = _rollout(32768, 6, seed=8, random_map_width=8, random_map_height=8, random_map_percentage_of_connections=0.8,
                 traffic_density=0.2, random_map_obstacle_probability=0.5)
    assert a[2] > 0


def test_results_do_not_depend_on_sharding():
    """Global env ids: two half-size handles with env_id_base reproduce the full-size run (the
    single-GPU analogue of the multi-GPU shards of SURVEY.md 8e)."""
    import torch

    from pgtg_b200 import PGTGVectorEnv

    kw = dict(traffic_density=0.05, random_map_obstacle_probability=0.2, seed=21)
    full = PGTGVectorEnv(1000, device="cuda:0", **kw)
    lo = PGTGVectorEnv(600, device="cuda:0", env_id_base=0, **kw)
    hi = PGTGVectorEnv(400, device="cuda:0", env_id_base=600, **kw)
    for e in (full, lo, hi):
        e.reset()
    rng = np.random.default_rng(1)
    for t in range(12):
        act = torch.from_numpy(rng.integers(0, 9, 1000).astype(np.int32)).cuda()
        of, rf, *_ = full.step(act)
        ol, rl, *_ = lo.step(act[:600].contiguous())
        oh, rh, *_ = hi.step(act[600:].contiguous())
        assert torch.equal(full._t["obs_map"], torch.cat([lo._t["obs_map"], hi._t["obs_map"]]))
        assert torch.equal(rf, torch.cat([rl, rh]))
    s = np.array([lo.raw.stats(), hi.raw.stats()]).sum(0)
    assert np.allclose(s, full.raw.stats())
    for e in (full, lo, hi):
        e.close()


@pytest.mark.parametrize("kw", [dict(), dict(random_map_obstacle_probability=0.4), dict(random_map_width=3, random_map_height=3),
                                dict(use_sliding_observation_window=True, sliding_observation_window_size=5, use_next_subgoal_direction=True,
                                     random_map_obstacle_probability=0.3),
                                dict(use_next_subgoal_direction=True), dict(use_sliding_observation_window=True, final_observation=True)],
                         ids=["default", "obstacles", "3x3", "sliding5+nsd", "nsd", "sliding9+final"])
def test_specialised_kernels_match_general(kw, monkeypatch):
    """The lean tick (plain, and its SLIDE variant for the sliding window / next_subgoal_direction) and the tabled map
    generation are compile-time specialisations of the general kernels (same source, cold code not emitted): switching
    them off must not change one bit."""
    import torch

    from pgtg_b200 import PGTGVectorEnv

    n, ticks = 20000, 40  # ragged last CTA; enough ticks for several generations of maps per env
    g = torch.Generator(device="cuda:0")
    g.manual_seed(11)
    actions = [torch.randint(0, 9, (n,), device="cuda:0", dtype=torch.int32, generator=g) for _ in range(ticks)]

    def run():
        env = PGTGVectorEnv(n, device="cuda:0", seed=5, **kw)
        env.reset()
        out = []
        for a in actions:
            obs, rew, term, trunc, info = env.step(a)
            row = [env._t["obs_map"].clone(), obs["position"].clone(), obs["velocity"].clone(), rew.clone(), term.clone(), trunc.clone()]
            if "next_subgoal_direction" in obs:
                row.append(obs["next_subgoal_direction"].clone())
            if kw.get("final_observation"):
                fin = info["_final_observation"]
                row.append(fin.clone())
                row.append(env._t["final_obs_map"][fin].clone())
            out.append(tuple(row))
        return out, env.episode_stats()

    fast, fast_stats = run()
    monkeypatch.setenv("PGTG_NO_LEAN", "1")
    monkeypatch.setenv("PGTG_NO_TABLED", "1")
    slow, slow_stats = run()
    for t, (a, b) in enumerate(zip(fast, slow)):
        for x, y in zip(a, b):
            assert torch.equal(x, y), f"tick {t}"
    # (sums of returns are accumulated with atomics: the order, hence the last bits, may differ)
    assert fast_stats.keys() == slow_stats.keys()
    for k in fast_stats:
        assert fast_stats[k] == pytest.approx(slow_stats[k], rel=1e-12), k
    # third leg: lean tick, tabled map generation through the staged kernel instead of the register-resident one
    monkeypatch.delenv("PGTG_NO_LEAN")
    monkeypatch.delenv("PGTG_NO_TABLED")
    monkeypatch.setenv("PGTG_NO_MAP_IN_REGISTERS", "1")
    staged, _ = run()
    for t, (a, b) in enumerate(zip(fast, staged)):
        for x, y in zip(a, b):
            assert torch.equal(x, y), f"tick {t} (staged map generation)"


@pytest.mark.gpu
def test_lean_final_observation_matches_general(monkeypatch):
    """final_observation=True on the plain configuration runs the lean FINAL instantiation: terminal observations (rows of the
    finished envs), the observation after the reset and every other output equal the general tick's, bit for bit."""
    import torch

    from pgtg_b200 import PGTGVectorEnv

    n, ticks = 20000, 30
    g = torch.Generator(device="cuda:0")
    g.manual_seed(3)
    actions = [torch.randint(0, 9, (n,), device="cuda:0", dtype=torch.int32, generator=g) for _ in range(ticks)]

    def run():
        env = PGTGVectorEnv(n, device="cuda:0", seed=8, final_observation=True, random_map_obstacle_probability=0.5, max_episode_steps=6)
        info = env.raw.kernel_info()
        env.reset()
        out = []
        for a in actions:
            obs, rew, term, trunc, inf = env.step(a)
            done = term | trunc
            fo = inf["final_observation"]
            out.append((env._t["obs_map"].clone(), rew.clone(), done.clone(), env._t["final_obs_map"][done].clone(), fo["position"][done].clone(), fo["velocity"][done].clone()))
        env.close()
        return out, info

    fast, info_fast = run()
    monkeypatch.setenv("PGTG_NO_LEAN", "1")
    slow, info_slow = run()
    assert "lean+final" in info_fast and "general" in info_slow
    finished = 0
    for t, (a, b) in enumerate(zip(fast, slow)):
        for x, y in zip(a, b):
            assert torch.equal(x, y), f"tick {t}"
        finished += int(a[2].sum())
    assert finished > 100000


@pytest.mark.gpu
def test_inline_map_generation_matches_the_pipeline(monkeypatch):
    """PGTG_INLINE_MAPGEN: the lean tick rebuilds consumed ring slots itself instead of queueing requests for the
    map-generation kernel (faster below ~100 k envs, where the step is launch-bound; slower at the headline size --
    DESIGN.md section 7). Same maps, same everything."""
    import torch

    from pgtg_b200 import PGTGVectorEnv

    n, ticks = 30000, 30
    g = torch.Generator(device="cuda:0")
    g.manual_seed(3)
    actions = [torch.randint(0, 9, (n,), device="cuda:0", dtype=torch.int32, generator=g) for _ in range(ticks)]

    def run():
        env = PGTGVectorEnv(n, device="cuda:0", seed=9)
        env.reset()
        out = []
        for a in actions:
            obs, rew, term, trunc, _ = env.step(a)
            out.append((env._t["obs_map"].clone(), obs["position"].clone(), rew.clone(), term.clone()))
        launches = env.launch_count()
        st = env.episode_stats()
        env.close()
        return out, st, launches

    a, sa, la = run()
    monkeypatch.setenv("PGTG_INLINE_MAPGEN", "1")
    b, sb, lb = run()
    assert lb < la  # no map-generation launches after the reset
    for t, (x, y) in enumerate(zip(a, b)):
        for u, v in zip(x, y):
            assert torch.equal(u, v), f"tick {t}"
    assert sa["episodes"] == sb["episodes"] > 0
