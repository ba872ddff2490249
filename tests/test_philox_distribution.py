"""Philox mode is distribution-exact, not draw-exact: its draw specification (per-car blocks, Feistel placement, tile-major index
spaces, edge draws over untried grid edges with shared words -- DESIGN.md section 5) must leave every distribution of the
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
    def tile_of_car(s):
        m = np.arange(s["cars"].shape[1])[None, :] < s["num_cars"][:, None]
        t = (s["cars"][:, :, 2] // 9) * 4 + s["cars"][:, :, 1] // 9
        return np.bincount(t[m], minlength=T)[:T] / m.sum()
    _close(tile_of_car(sa), tile_of_car(sb), min(na, nb), "car tiles")
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
