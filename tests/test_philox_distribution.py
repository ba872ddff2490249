"""Philox mode is distribution-exact, not draw-exact: its draw specification (per-car blocks, Feistel placement, tile-major index
spaces, edge draws over untried grid edges by rejection -- DESIGN.md section 5) must leave every distribution of the
reference untouched. Checked here on the oracle by comparing Philox-mode statistics with the numpy-exact mode (which is the
reference draw for draw) over many independent envs: tile-type frequencies per tile, obstacle frequencies, initial car
counts and car-position occupancy, spawner choice after many ticks."""
import warnings

import numpy as np

from oracle.oracle import OracleVectorEnv


def _envs(n, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = OracleVectorEnv(num_envs=n, threads=8, seed=12345, rng_mode="philox", **kw)
        b = OracleVectorEnv(num_envs=n, threads=8, rng_mode="numpy", **kw)
    a.reset()
    b.reset(seeds=777 + np.arange(n, dtype=np.int64))
    return a, b


def _close(p, q, n, what, sigmas=5.0):
    """two empirical frequency vectors over n independent trials each"""
    p, q = np.asarray(p, float), np.asarray(q, float)
    se = np.sqrt((p * (1 - p) + q * (1 - q)) / n) + 1e-9
    z = np.abs(p - q) / se
    assert z.max() < sigmas, f"{what}: largest deviation {z.max():.1f} sigma at {int(z.argmax())} ({p.ravel()[z.argmax()]:.4f} vs {q.ravel()[z.argmax()]:.4f})"


def test_map_and_traffic_distributions_match_the_reference_exact_mode():
    n = 30000
    kw = dict(random_map_obstacle_probability=0.4, traffic_density=0.05, random_map_percentage_of_connections=0.6)
    a, b = _envs(n, **kw)
    sa, sb = a.get_state(), b.get_state()
    T = sa["tiles"].shape[1]
    # frequency of every exits pattern at every tile; obstacle type frequencies; subgoal-path membership
    fa = np.stack([(sa["tiles"] & 15) == k for k in range(16)]).mean(axis=1)
    fb = np.stack([(sb["tiles"] & 15) == k for k in range(16)]).mean(axis=1)
    _close(fa, fb, n, "tile types")
    _close(np.stack([((sa["tiles"] >> 4) & 7) == k for k in range(5)]).mean(axis=1), np.stack([((sb["tiles"] >> 4) & 7) == k for k in range(5)]).mean(axis=1), n, "obstacle types")
    _close((((sa["tiles"] >> 11) & 7) > 0).mean(axis=0), (((sb["tiles"] >> 11) & 7) > 0).mean(axis=0), n, "path tiles")
    # initial traffic: number of cars, which tile a car stands in, profile and route mix
    ca, cb = sa["num_cars"], sb["num_cars"]
    assert abs(ca.mean() - cb.mean()) < 5 * np.sqrt((ca.var() + cb.var()) / n)
    def car_hist(s, col, bins):
        m = np.arange(s["cars"].shape[1])[None, :] < s["num_cars"][:, None]
        return np.bincount(s["cars"][:, :, col][m], minlength=bins)[:bins] / m.sum(), int(m.sum())
    (pa, na), (pb, nb) = car_hist(sa, 4, 5), car_hist(sb, 4, 5)
    _close(pa, pb, min(na, nb), "driver profiles")
    (ra, _), (rb, _) = car_hist(sa, 3, 20), car_hist(sb, 3, 20)
    _close(ra, rb, min(na, nb), "routes")
    def cars_per_tile(s):
        """[env, tile] counts: the envs are the independent trials (the cars of one env share its map)"""
        m = np.arange(s["cars"].shape[1])[None, :] < s["num_cars"][:, None]
        t = (s["cars"][:, :, 2] // 9) * 4 + s["cars"][:, :, 1] // 9
        out = np.zeros((m.shape[0], T))
        for k in range(T):
            out[:, k] = ((t == k) & m).sum(axis=1)
        return out
    ta, tb = cars_per_tile(sa), cars_per_tile(sb)
    z = np.abs(ta.mean(axis=0) - tb.mean(axis=0)) / np.sqrt((ta.var(axis=0) + tb.var(axis=0)) / n + 1e-12)
    assert z.max() < 5.0, f"cars per tile: largest deviation {z.max():.1f} sigma at tile {int(z.argmax())}"
    a.close(); b.close()


def test_traffic_dynamics_statistics_match():
    """after 40 ticks of idling agents: patience, delay and respawn statistics of the two modes agree"""
    n = 6000
    a, b = _envs(n, traffic_density=0.15, ignore_traffic_collisions=True, random_map_percentage_of_connections=0.8)
    act = np.full(n, 4, np.int32)
    for _ in range(40):
        a.step(act); b.step(act)
    sa, sb = a.get_state(), b.get_state()
    def stat(s):
        m = np.arange(s["cars"].shape[1])[None, :] < s["num_cars"][:, None]
        c = s["cars"]
        return np.array([c[:, :, 5][m].mean(), (c[:, :, 5][m] == 0).mean(), (c[:, :, 6][m] > 0).mean(), (c[:, :, 0][m] >= s["num_cars"].repeat(c.shape[1]).reshape(c.shape[:2])[m]).mean()]), int(m.sum())
    (xa, na), (xb, nb) = stat(sa), stat(sb)
    # mean patience, share of cars that just moved, share in a reaction delay, share of respawned cars
    assert np.all(np.abs(xa - xb) < np.array([0.12, 0.01, 0.01, 0.01])), (xa, xb)
    a.close(); b.close()


def test_edge_removal_pair_statistics_match():
    """The edge draws of the Philox map stream are by rejection over fixed-width chunks of its words (pgtg_logic.cuh
    generate_map); the removal process must keep the reference's joint law of the surviving edges: every pair
    P(edge i kept and edge j kept) on the default 4x4 grid, and the number of kept edges."""
    n = 40000
    a, b = _envs(n)
    ta, tb = a.get_state()["tiles"], b.get_state()["tiles"]
    def edges(t):
        ex = (t & 15).astype(np.uint8)
        east = ((ex >> 1) & 1)[:, [i for i in range(16) if i % 4 != 3]]   # start / goal border exits are west / east of column 0 / 3
        south = ((ex >> 2) & 1)[:, :12]
        return np.concatenate([east, south], axis=1).astype(np.float64)
    ea, eb = edges(ta), edges(tb)
    assert ea.shape[1] == 24
    _close(ea.T @ ea / n, eb.T @ eb / n, n, "edge pairs")
    ka, kb = ea.sum(axis=1), eb.sum(axis=1)
    _close(np.bincount(ka.astype(int), minlength=25) / n, np.bincount(kb.astype(int), minlength=25) / n, n, "kept edges")
    # add_connections_to_borders: a uniformly random 7-subset of the 14 border slots (rejection draws in Philox mode)
    def borders(t):
        ex = (t & 15).astype(np.uint8)
        cols = [(k, 0) for k in range(4)] + [(k, 1) for k in (7, 11, 15)] + [(k, 2) for k in range(12, 16)] + [(k, 3) for k in (0, 4, 8)]
        return np.stack([(ex[:, k] >> d) & 1 for k, d in cols], axis=1).astype(np.float64)
    ba, bb = borders(ta), borders(tb)
    assert (ba.sum(axis=1) == 7).all() and (bb.sum(axis=1) == 7).all()
    _close(ba.T @ ba / n, bb.T @ bb / n, n, "border slot pairs")
    a.close(); b.close()
