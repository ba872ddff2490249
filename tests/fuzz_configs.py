"""Deterministic pseudo-random PGTGEnv constructor arguments for the fuzz suites."""
from __future__ import annotations

import numpy as np

FEATURE_POOL = ["walls", "goals", "ice", "broken road", "sand", "traffic", "traffic_light", "traffic_light_green",
                "start", "used subgoal", "car_spawner", "subgoal", "final goal", "wall", "bogus"]


def random_kwargs(i: int) -> dict:
    r = np.random.default_rng(10_000 + i)
    W, H = int(r.integers(1, 7)), int(r.integers(1, 7))
    kw = dict(random_map_width=W, random_map_height=H,
              random_map_percentage_of_connections=float(r.choice([0.0, 0.25, 0.5, 0.75, 0.85, 1.0])))
    if r.random() < 0.7:
        kw["random_map_obstacle_probability"] = float(r.choice([0.2, 0.5, 1.0]))
        for k in ("ice", "broken_road", "sand", "traffic_light"):
            kw[f"random_map_{k}_probability_weight"] = float(r.choice([0, 1, 1, 3]))
        if sum(kw[f"random_map_{k}_probability_weight"] for k in ("ice", "broken_road", "sand", "traffic_light")) == 0:
            kw["random_map_ice_probability_weight"] = 1.0
    mode = r.integers(0, 4)
    if mode == 1:
        kw["random_map_start_position"] = "random"
        kw["random_map_goal_position"] = "random"
        if W + H > 3 and r.random() < 0.5:
            kw["random_map_minimum_distance_between_start_and_goal"] = int(r.integers(1, W + H - 1))
    elif mode == 2:
        kw["random_map_start_position"] = (0, int(r.integers(0, H)))
        kw["random_map_goal_position"] = (W - 1, int(r.integers(0, H)))
    elif mode == 3 and (W > 1 or H > 1):
        kw["random_map_start_position"] = (int(r.integers(0, W)), 0, "north")
        kw["random_map_goal_position"] = (int(r.integers(0, W)), H - 1, "south")
    if r.random() < 0.6:
        kw["traffic_density"] = float(r.choice([0.02, 0.05, 0.2, 0.5, 1.0]))
        kw["ignore_traffic_collisions"] = bool(r.random() < 0.5)
        kw["traffic_light_phases_duration"] = tuple(int(v) for v in r.integers(1, 6, 3))
        pct = r.random(5)
        pct[r.integers(0, 5)] = 0.0
        for name, v in zip(("conservative", "normal", "aggressive", "elderly", "reckless"), pct):
            kw[f"{name}_driver_percentage"] = float(v)
    if r.random() < 0.4:
        kw["use_sliding_observation_window"] = True
        kw["sliding_observation_window_size"] = int(r.integers(0, 9))
    kw["use_next_subgoal_direction"] = bool(r.random() < 0.5)
    if r.random() < 0.5:
        n = int(r.integers(1, 9))
        kw["features_to_include_in_observation"] = [str(f) for f in r.choice(FEATURE_POOL, size=n, replace=False)]
    if r.random() < 0.5:
        kw.update(standing_still_penalty=int(r.integers(0, 5)), already_visited_position_penalty=int(r.integers(0, 5)),
                  final_goal_bonus=int(r.integers(0, 50)), sum_subgoals_reward=int(r.integers(1, 300)), crash_penalty=int(r.integers(0, 200)),
                  traffic_light_violation_penalty=int(r.integers(0, 60)))
    kw.update(ice_probability=float(r.choice([0.1, 0.5, 1.0])), street_damage_probability=float(r.choice([0.1, 0.5])),
              sand_probability=float(r.choice([0.2, 0.7])))
    kw["separate_reward_cost"] = bool(r.random() < 0.3)
    return kw
