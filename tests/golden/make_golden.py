#!/usr/bin/env python3
"""Generate the golden traces under tests/golden/ by running the UNMODIFIED reference PGTGEnv
(/root/reference, behind the stand-ins in oracle/shims) -- build container only.

    python tests/golden/make_golden.py [name ...]

Each trace_<name>.npz holds, for N reference envs seeded seed+i and T ticks with same-step
auto-reset: the actions played, every recorded np_random draw (values, tags, per-env offsets),
and every output per tick (observation planes, position, velocity, next_subgoal_direction, reward,
cost, terminated, truncated, step info, terminal observations, agent state, car lists, the map
plan with its subgoal directions, and the rule engine's agent direction). The oracle
(tests/test_oracle_golden.py), the host emulation of the kernels (tests/test_emu_golden.py) and
the CUDA kernels (tests/test_gpu_golden.py) must reproduce all of it bit for bit.
"""
import json
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# Fixed map plans written for these tests (MapPlan dicts, map_generator.py:10-40 format).
def _tile(n, e, s, w, otype=None, omask=None):
    t = {"exits": [n, e, s, w]}
    if otype:
        t["obstacle_type"], t["obstacle_mask"] = otype, omask
    return t


MAPS = {
    # one straight tile west -> east
    "straight_1x1": dict(width=1, height=1, start=[0, 0, "west"], goal=[0, 0, "east"], map=[[_tile(0, 1, 0, 1)]]),
    # one crossing
    "crossing_1x1": dict(width=1, height=1, start=[0, 0, "west"], goal=[0, 0, "east"], map=[[_tile(1, 1, 1, 1)]]),
    # corridor of four tiles with sand, ice and a traffic light on the way
    "corridor_4x1": dict(width=4, height=1, start=[0, 0, "west"], goal=[3, 0, "east"], map=[[
        _tile(0, 1, 0, 1), _tile(0, 1, 0, 1, "sand", "chess_field"), _tile(1, 1, 0, 1, "traffic_light", "traffic_light_east_and_west"),
        _tile(0, 1, 0, 1, "ice", "left_half")]]),
    # 5x3 serpentine that uses every path direction, plus dead ends (config 2 "fixed example map")
    "serpentine_5x3": dict(width=5, height=3, start=[0, 2, "west"], goal=[4, 0, "east"], map=[
        [_tile(0, 1, 1, 0), _tile(0, 1, 0, 1), _tile(0, 0, 1, 1), _tile(0, 1, 1, 0), _tile(0, 1, 0, 1)],
        [_tile(1, 0, 1, 0), _tile(0, 0, 0, 0), _tile(1, 1, 0, 0), _tile(1, 0, 0, 1), _tile(0, 0, 1, 0)],
        [_tile(1, 1, 0, 1), _tile(0, 0, 0, 1), _tile(0, 1, 0, 0), _tile(0, 1, 0, 1), _tile(1, 0, 0, 1)]]),
}

# name -> (PGTGEnv kwargs, recorder options)
TRACES = {
    "default": (dict(), dict(n=8, ticks=60)),
    "config2_fixed_map": (dict(map_plan="serpentine_5x3"), dict(n=16, ticks=64)),
    "obstacles": (dict(random_map_obstacle_probability=0.6), dict(policy="seek", n=6)),
    "traffic": (dict(traffic_density=0.05, random_map_obstacle_probability=0.3), dict(policy="seek", n=6)),
    "traffic_dense": (dict(traffic_density=0.3, random_map_obstacle_probability=0.8, random_map_percentage_of_connections=0.9), dict(policy="seek")),
    "traffic_lights_nocollide": (dict(traffic_density=0.2, random_map_obstacle_probability=0.8, ignore_traffic_collisions=True,
                                      random_map_percentage_of_connections=1.0, random_map_traffic_light_probability_weight=5), dict(policy="seek", epsilon=0.1)),
    "sliding_nsd": (dict(use_sliding_observation_window=True, sliding_observation_window_size=3, traffic_density=0.1,
                         use_next_subgoal_direction=True, random_map_obstacle_probability=0.5), dict(policy="seek")),
    "sliding_wide": (dict(use_sliding_observation_window=True, sliding_observation_window_size=6, random_map_width=2, random_map_height=3,
                          traffic_density=0.05), dict(policy="seek", ticks=40)),
    "next_subgoal_direction": (dict(use_next_subgoal_direction=True), dict(policy="seek", epsilon=0.5)),
    "random_start_goal": (dict(random_map_start_position="random", random_map_goal_position="random", random_map_width=5,
                               random_map_height=3, random_map_minimum_distance_between_start_and_goal=3), dict(n=8)),
    "two_tuple_start_goal": (dict(random_map_start_position=(0, 0), random_map_goal_position=(-1, -1), random_map_width=3,
                                  random_map_height=5), dict(policy="seek")),
    "penalties": (dict(standing_still_penalty=3, already_visited_position_penalty=7, final_goal_bonus=11, sum_subgoals_reward=90,
                       crash_penalty=55, random_map_obstacle_probability=1.0, sand_probability=0.6, ice_probability=0.5,
                       street_damage_probability=0.3), dict(policy="seek", epsilon=0.4)),
    "separate_reward_cost": (dict(separate_reward_cost=True, traffic_density=0.1, random_map_obstacle_probability=0.7,
                                  random_map_traffic_light_probability_weight=4, traffic_light_phases_duration=(2, 1, 4)), dict(policy="seek")),
    "fixed_corridor": (dict(map_plan="corridor_4x1", traffic_density=0.1), dict(policy="seek", epsilon=0.2)),
    "fixed_crossing_full_traffic": (dict(map_plan="crossing_1x1", traffic_density=1.0, ignore_traffic_collisions=True), dict(ticks=60)),
    "fixed_straight": (dict(map_plan="straight_1x1"), dict(policy="seek", epsilon=0.2, ticks=40)),
    "fixed_serpentine_traffic": (dict(map_plan="serpentine_5x3", traffic_density=0.3, ignore_traffic_collisions=True), dict(policy="seek")),
    "big_8x8": (dict(random_map_width=8, random_map_height=8, random_map_percentage_of_connections=0.8, traffic_density=0.2,
                     random_map_obstacle_probability=0.5), dict(policy="seek", n=2, ticks=30)),
    "time_limit": (dict(traffic_density=0.05), dict(policy="seek", max_episode_steps=7)),
    "literal_features": (dict(features_to_include_in_observation=["walls", "goals", "traffic", "traffic_light", "start", "used subgoal",
                                                                  "car_spawner", "subgoal", "final goal", "wall", "ice", "bogus"],
                              random_map_obstacle_probability=0.9, random_map_traffic_light_probability_weight=6, traffic_density=0.1), dict(policy="seek")),
    # 1-wide / 1-high maps: t + W == t + 1, found by tools/fuzz_reference.py
    "narrow_1x4": (dict(random_map_width=1, random_map_height=4, random_map_percentage_of_connections=0.5,
                        random_map_start_position=(0, 0, "north"), random_map_goal_position=(0, 3, "south"), use_next_subgoal_direction=True,
                        random_map_obstacle_probability=0.5, traffic_density=0.2), dict(policy="seek", ticks=60)),
    "narrow_5x1_random_ends": (dict(random_map_width=5, random_map_height=1, random_map_start_position="random", random_map_goal_position="random",
                                    traffic_density=0.1, use_sliding_observation_window=True, sliding_observation_window_size=2),
                               dict(policy="seek", ticks=60)),
    "driver_mix": (dict(traffic_density=0.25, conservative_driver_percentage=0.0, normal_driver_percentage=0.1, aggressive_driver_percentage=0.5,
                        elderly_driver_percentage=0.1, reckless_driver_percentage=0.3, random_map_percentage_of_connections=0.7,
                        random_map_obstacle_probability=0.6, random_map_traffic_light_probability_weight=3, ignore_traffic_collisions=True),
                   dict(policy="seek", epsilon=0.15)),
}


def main(argv):
    from oracle import ref_runner

    names = argv or list(TRACES)
    for name in names:
        kw, opt = TRACES[name]
        kw, opt = dict(kw), dict(opt)
        n, ticks = opt.pop("n", 4), opt.pop("ticks", 80)
        mes = opt.pop("max_episode_steps", None)
        ref_kw = dict(kw)
        tmp_json = None
        if "map_plan" in kw:  # the reference loads fixed maps from a JSON file
            tmp_json = os.path.join("/tmp", f"pgtg_golden_{kw['map_plan']}.json")
            with open(tmp_json, "w") as f:
                json.dump(MAPS[kw["map_plan"]], f)
        t0 = time.time()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            tr = ref_runner.record_trace(ref_kw if tmp_json is None else {**{k: v for k, v in kw.items() if k != "map_plan"}, "map_path": tmp_json},
                                         num_envs=n, ticks=ticks, seed=1000 + len(name), max_episode_steps=mes, **opt)
        meta = json.loads(bytes(tr["meta"]).decode())
        meta["kwargs"] = kw  # portable form: the map by name, not by temp path
        meta["maps"] = {kw["map_plan"]: MAPS[kw["map_plan"]]} if "map_plan" in kw else {}
        tr["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        path = os.path.join(HERE, f"trace_{name}.npz")
        np.savez_compressed(path, **tr)
        print(f"{name:32s} {time.time() - t0:5.1f}s draws={tr['tape_offsets'][-1]:7d} done={int(tr['terminated'].sum()):4d} "
              f"trunc={int(tr['truncated'].sum()):3d} reward>0={int((tr['reward'] > 0).sum()):3d} "
              f"brake={int(((tr['step_flags'] & 2) > 0).sum()):3d} cars<={int(tr['num_cars'].max()):3d} {os.path.getsize(path) // 1024} KiB")


if __name__ == "__main__":
    main(sys.argv[1:])
