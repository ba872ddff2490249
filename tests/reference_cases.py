"""Known-answer cases restated from the reference's own unit tests (tests/test_environment.py,
tests/test_map.py of Inuri04/pgtg), written against the array-level env surface so the same cases
run on the CPU oracle, on the host emulation of the kernels and on the CUDA kernels.

`make(**kwargs)` must return an env adapter (OracleVectorEnv or NativeAdapter) with num_envs = 1."""
from __future__ import annotations

import numpy as np


def _tile(n, e, s, w):
    return {"exits": [n, e, s, w]}


# tests/test_data/1x1_map.json and 4x1_map.json of the reference, as plan dicts
MAP_1X1 = dict(width=1, height=1, start=[0, 0, "west"], goal=[0, 0, "east"], map=[[_tile(0, 1, 0, 1)]])
MAP_4X1 = dict(width=4, height=1, start=[0, 0, "west"], goal=[3, 0, "east"], map=[[_tile(0, 1, 0, 1)] * 4])
# tests/test_data/map_with_all_deadends.json
MAP_DEADENDS = dict(width=5, height=5, start=[0, 4, "west"], goal=[4, 0, "east"], map=[
    [_tile(0, 0, 0, 0), _tile(0, 0, 0, 0), _tile(0, 1, 1, 0), _tile(0, 1, 0, 1), _tile(0, 1, 1, 1)],
    [_tile(0, 0, 0, 0), _tile(0, 0, 0, 0), _tile(1, 0, 1, 0), _tile(0, 0, 0, 0), _tile(1, 0, 0, 0)],
    [_tile(0, 0, 0, 0), _tile(0, 1, 0, 0), _tile(1, 1, 1, 1), _tile(0, 0, 0, 1), _tile(0, 0, 0, 0)],
    [_tile(0, 0, 1, 0), _tile(0, 0, 0, 0), _tile(1, 0, 1, 0), _tile(0, 0, 0, 0), _tile(0, 0, 0, 0)],
    [_tile(1, 1, 0, 1), _tile(0, 1, 0, 1), _tile(1, 0, 0, 1), _tile(0, 0, 0, 0), _tile(0, 0, 0, 0)]])

ZERO = dict(final_goal_bonus=0, crash_penalty=0, standing_still_penalty=0, already_visited_position_penalty=0, sum_subgoals_reward=0)


def _step(env, a):
    env.step(np.array([a], np.int32))
    return float(env.reward[0])


def subgoal_reward(make, sum_subgoals_reward):
    """TestReward.test_subgoal_reward (:870-892): every tile of the 4x1 corridor pays sum / 4."""
    env = make(map_plan=MAP_4X1, **{**ZERO, "sum_subgoals_reward": sum_subgoals_reward})
    env.reset()
    for n in range(4):
        if n == 0:
            _step(env, 7)
            for _ in range(6):
                _step(env, 4)
        else:
            for _ in range(8):
                _step(env, 4)
        assert _step(env, 4) == sum_subgoals_reward / 4
    env.close()


def final_goal_bonus(make, bonus):
    """TestReward.test_final_goal_bonus_reward (:894-912)."""
    env = make(map_plan=MAP_1X1, **{**ZERO, "final_goal_bonus": bonus})
    env.reset()
    _step(env, 7)
    for _ in range(6):
        _step(env, 4)
    assert _step(env, 4) == bonus
    assert env.terminated[0] == 1
    env.close()


def crash_penalty(make, penalty):
    """TestReward.test_crash_penalty (:914-934): position forced to (0, 4), drive into the wall."""
    env = make(map_plan=MAP_1X1, **{**ZERO, "crash_penalty": penalty})
    env.reset()
    env.set_state(agent=np.array([[0, 4, 0, 0]], np.int32))
    _step(env, 5)
    assert _step(env, 4) == -1 * penalty
    assert env.terminated[0] == 1
    env.close()


def standing_still_penalty(make, penalty):
    """TestReward.test_standing_still_penalty_reward (:936-975)."""
    env = make(map_plan=MAP_1X1, **{**ZERO, "standing_still_penalty": penalty})
    env.reset()
    for _ in range(3):
        assert _step(env, 4) == -1 * penalty
        assert env.obs_velocity[0].tolist() == [0, 0]
    assert _step(env, 7) == 0
    for _ in range(3):
        assert _step(env, 4) == 0
        assert env.obs_velocity[0].tolist() != [0, 0]
    assert _step(env, 1) == 0
    for _ in range(3):
        assert _step(env, 4) == -1 * penalty
        assert env.obs_velocity[0].tolist() == [0, 0]
    env.close()


def already_visited_penalty(make, penalty):
    """TestReward.test_already_visited_position_penalty_reward (:977-1083), same action script."""
    env = make(map_plan=MAP_1X1, **{**ZERO, "already_visited_position_penalty": penalty})
    env.reset()
    env.set_state(agent=np.array([[0, 3, 0, 0]], np.int32))
    P = -1 * penalty
    script = [(7, 0), (4, 0), (4, 0), (4, 0), (1, P), (4, 0), (4, 0), (4, 0), (1, P), (7, P), (7, P), (4, 0), (2, 0), (4, 0),
              (6, 0), (4, 0), (0, 0), (2, 0), (1, 0), (4, 0), (7, 0), (6, P), (8, P), (7, P)]
    for i, (a, want) in enumerate(script):
        got = _step(env, a)
        assert got == want, (i, a, got, want)
    env.close()


def initial_traffic(make, density, count):
    """TestTraffic.TestInitialPlacement (:645-703): int(18 * density) cars on distinct lane squares."""
    env = make(map_plan=MAP_1X1, traffic_density=density)
    env.reset()
    st = env.get_state()
    assert st["num_cars"][0] == count
    cars = st["cars"][0, :count]
    pos = [(int(c[1]), int(c[2])) for c in cars]
    assert len(set(pos)) == count
    assert [int(c[0]) for c in cars] == list(range(count))
    lanes = {(x, 3) for x in range(9)} | {(x, 5) for x in range(9)}
    assert set(pos) <= lanes
    if density == 1:
        assert set(pos) == lanes
    env.close()


def respawning_keeps_count(make, density, count):
    """TestTraffic.TestRespawning (:718-757): cars that leave the map are replaced one for one."""
    env = make(map_plan=MAP_1X1, traffic_density=density, ignore_traffic_collisions=True)
    env.reset()
    for _ in range(20):
        _step(env, 4)
        assert env.get_state()["num_cars"][0] == count
    env.close()


def car_spawners_on_deadend_map(make):
    """tests/test_map.py:41-59: the six car spawners of map_with_all_deadends, read back through the
    literal "car_spawner" observation plane in sliding-window mode."""
    want = {(0, 36 + 5), (36 + 8, 3), (36 + 5, 9 + 5), (9 + 3, 18 + 5), (3, 27 + 3), (27 + 5, 18 + 3)}
    k = 15
    env = make(map_plan=MAP_DEADENDS, features_to_include_in_observation=["car_spawner"],
               use_sliding_observation_window=True, sliding_observation_window_size=k)
    env.reset()
    found = set()
    # sweep the window over the 45x45 map by teleporting the agent
    for cx in (15, 30):
        for cy in (15, 30):
            env.set_state(agent=np.array([[cx, cy, 0, 0]], np.int32))
            env.observe()
            plane = env.obs_map[0, 0]
            for ix, iy in zip(*np.nonzero(plane)):
                found.add((cx - k + int(ix), cy - k + int(iy)))
    assert found == want
    env.close()


ALL = [
    ("subgoal_reward_100", lambda mk: subgoal_reward(mk, 100)),
    ("subgoal_reward_444", lambda mk: subgoal_reward(mk, 444)),
    ("subgoal_reward_4", lambda mk: subgoal_reward(mk, 4)),
    ("subgoal_reward_0", lambda mk: subgoal_reward(mk, 0)),
    ("final_goal_bonus_100", lambda mk: final_goal_bonus(mk, 100)),
    ("final_goal_bonus_1", lambda mk: final_goal_bonus(mk, 1)),
    ("crash_penalty_100", lambda mk: crash_penalty(mk, 100)),
    ("crash_penalty_1", lambda mk: crash_penalty(mk, 1)),
    ("standing_still_10", lambda mk: standing_still_penalty(mk, 10)),
    ("already_visited_10", lambda mk: already_visited_penalty(mk, 10)),
    ("already_visited_1000", lambda mk: already_visited_penalty(mk, 1000)),
    ("initial_traffic_full", lambda mk: initial_traffic(mk, 1, 18)),
    ("initial_traffic_half", lambda mk: initial_traffic(mk, 0.5, 9)),
    ("respawn_full", lambda mk: respawning_keeps_count(mk, 1, 18)),
    ("respawn_half", lambda mk: respawning_keeps_count(mk, 0.5, 9)),
    ("car_spawners_deadends", car_spawners_on_deadend_map),
]
