"""Lock-step comparison of two vector envs in Philox mode (oracle vs emulation on the CPU, oracle vs
CUDA on the GPU). Both implement the same draw specification (include/pgtg_b200.h), so every array
must be bit-equal at every tick."""
from __future__ import annotations

import numpy as np

ARRAYS = ("obs_map", "obs_position", "obs_velocity", "obs_nsd", "reward", "cost", "terminated", "truncated",
          "step_state", "step_flags")


def _eq(name, a, b, t):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"tick {t}: {name} differs at {bad[:5].tolist()} ({len(bad)} cells): got {a[tuple(bad[0])]} want {b[tuple(bad[0])]}")


def compare(env, ora, ticks, action_seed=0, final_obs=True, state_every=0, check_obs_every=1, stay=0.0):
    """env: implementation under test; ora: the oracle. Returns the number of finished episodes."""
    N = ora.N
    rng = np.random.default_rng(action_seed)
    env.reset()
    ora.reset()
    for name in ARRAYS[:4]:
        _eq(name, getattr(env, name), getattr(ora, name), -1)
    episodes = 0
    for t in range(ticks):
        a = rng.integers(0, 9, N).astype(np.int32)
        if stay > 0:  # mostly "no acceleration": the agent idles on the start line, so episodes (and their traffic) live long
            a = np.where(rng.random(N) < stay, 4, a).astype(np.int32)
        env.step(a)
        ora.step(a)
        for name in ARRAYS:
            if name == "obs_map" and check_obs_every > 1 and t % check_obs_every:
                continue
            _eq(name, getattr(env, name), getattr(ora, name), t)
        done = (ora.terminated | ora.truncated).astype(bool)
        episodes += int(done.sum())
        if final_obs and done.any():
            _eq("final_obs_map", env.final_obs_map[done], ora.final_obs_map[done], t)
            _eq("final_obs_position", env.final_obs_position[done], ora.final_obs_position[done], t)
            _eq("final_obs_velocity", env.final_obs_velocity[done], ora.final_obs_velocity[done], t)
            _eq("final_obs_nsd", env.final_obs_nsd[done], ora.final_obs_nsd[done], t)
        if state_every and (t + 1) % state_every == 0:
            compare_state(env, ora, t)
    compare_state(env, ora, ticks)
    return episodes


def compare_state(env, ora, t):
    a, b = env.get_state(), ora.get_state()
    assert not a["error"].any(), f"tick {t}: error flags {a['error'][a['error'] != 0][:4]}"
    assert not b["error"].any()
    for k in ("agent", "flat_tire", "light_counter", "elapsed", "num_cars", "tiles", "plan", "used"):
        _eq("state." + k, a[k], b[k], t)
    mc = min(a["cars"].shape[1], b["cars"].shape[1])
    _eq("state.cars", a["cars"][:, :mc], b["cars"][:, :mc], t)


# configurations exercised in Philox mode: name -> (kwargs, num_envs, ticks)
CONFIGS = {
    "default": (dict(), 300, 40),
    "config3_traffic_obstacles": (dict(traffic_density=0.05, random_map_obstacle_probability=0.2), 200, 40),
    "dense_lights": (dict(traffic_density=0.3, random_map_obstacle_probability=0.9, random_map_traffic_light_probability_weight=4,
                          random_map_percentage_of_connections=0.9, ignore_traffic_collisions=True, traffic_light_phases_duration=(3, 2, 4)), 40, 60),
    "sliding_nsd": (dict(use_sliding_observation_window=True, sliding_observation_window_size=5, use_next_subgoal_direction=True,
                         traffic_density=0.1, random_map_obstacle_probability=0.5), 60, 40),
    "random_start_goal": (dict(random_map_start_position="random", random_map_goal_position="random", random_map_width=6,
                               random_map_height=2, random_map_percentage_of_connections=0.3), 150, 30),
    "penalties_timelimit": (dict(standing_still_penalty=2, already_visited_position_penalty=5, final_goal_bonus=9, max_episode_steps=6,
                                 random_map_obstacle_probability=1.0, sand_probability=0.7, ice_probability=0.6, street_damage_probability=0.4,
                                 separate_reward_cost=True), 150, 40),
    "config4_big": (dict(random_map_width=8, random_map_height=8, random_map_percentage_of_connections=0.8, traffic_density=0.2,
                         random_map_obstacle_probability=0.5), 6, 25),
    "wide_16x16": (dict(random_map_width=16, random_map_height=16, random_map_percentage_of_connections=0.6, traffic_density=0.01,
                        random_map_obstacle_probability=0.3, use_next_subgoal_direction=True), 4, 12),
    "tiny_1x1": (dict(random_map_width=1, random_map_height=1, traffic_density=0.5, ignore_traffic_collisions=True), 130, 30),
    # 8-tile grids: the emulation build has the connectivity / path tables up to 13 edges, so these run the tabled edge
    # removal and the register-resident map assembly the headline configuration uses on the device (map_in_registers)
    "tabled_2x4_registers": (dict(random_map_width=2, random_map_height=4), 300, 40),
    "tabled_2x4_registers_obstacles": (dict(random_map_width=2, random_map_height=4, random_map_obstacle_probability=0.6,
                                            random_map_traffic_light_probability_weight=3, traffic_density=0.1), 200, 40),
    "tabled_4x2_registers_sparse": (dict(random_map_width=4, random_map_height=2, random_map_percentage_of_connections=0.2,
                                         max_episode_steps=12), 300, 40),
    "tabled_3x3_shared": (dict(random_map_width=3, random_map_height=3, random_map_obstacle_probability=0.3), 200, 40),
    # 16-tile strips and slabs: other edge / border-slot counts through the register-resident generator on the device
    # (1x16: 15 edges, 32 border slots, no grid faces; 8x2: 22 edges, 18 slots)
    "strip_1x16": (dict(random_map_width=1, random_map_height=16, random_map_obstacle_probability=0.2), 200, 40),
    "strip_16x1_traffic": (dict(random_map_width=16, random_map_height=1, traffic_density=0.1), 120, 40),
    "slab_8x2": (dict(random_map_width=8, random_map_height=2, random_map_percentage_of_connections=0.3), 200, 40),
    # car-free configurations outside the lean tick's promise (they run the traffic tick's parallel observation phases)
    "carfree_sliding_nsd": (dict(use_sliding_observation_window=True, sliding_observation_window_size=5, use_next_subgoal_direction=True,
                                 random_map_obstacle_probability=0.5), 200, 40),
    "carfree_nsd_penalties": (dict(use_next_subgoal_direction=True, standing_still_penalty=1, already_visited_position_penalty=2,
                                   random_map_width=3, random_map_height=5, max_episode_steps=9), 200, 40),
}

# long-lived episodes (the agent mostly idles) for the traffic dynamics: blocking chains, patience, push-through,
# lights, despawn / respawn bursts, occupancy counters; name -> (kwargs, num_envs, ticks, stay probability)
_T = lambda n, e, s, w: {"exits": [n, e, s, w]}  # noqa: E731
CROSSING = dict(width=1, height=1, start=[0, 0, "west"], goal=[0, 0, "east"], map=[[_T(1, 1, 1, 1)]])
RING_3X3 = dict(width=3, height=3, start=[0, 2, "west"], goal=[2, 0, "east"], map=[
    [_T(0, 1, 1, 0), _T(1, 1, 0, 1), _T(0, 1, 1, 1)],
    [_T(1, 1, 1, 0), _T(1, 1, 1, 1), _T(1, 0, 1, 1)],
    [_T(1, 1, 0, 1), _T(0, 1, 1, 1), _T(1, 0, 0, 1)]])
TRAFFIC_CONFIGS = {
    "crossing_full": (dict(map_plan=CROSSING, traffic_density=1.0, ignore_traffic_collisions=True), 70, 120, 0.97),
    "crossing_half_collide": (dict(map_plan=CROSSING, traffic_density=0.5), 70, 60, 0.9),
    "ring_dense_lights": (dict(map_plan=RING_3X3, traffic_density=0.6, ignore_traffic_collisions=True, traffic_light_phases_duration=(4, 2, 5)), 40, 150, 0.97),
    "train_py": (dict(random_map_width=4, random_map_height=4, random_map_obstacle_probability=0.2, random_map_percentage_of_connections=0.8,
                      traffic_density=0.2, conservative_driver_percentage=0.15, normal_driver_percentage=0.50, aggressive_driver_percentage=0.20,
                      elderly_driver_percentage=0.10, reckless_driver_percentage=0.05, sliding_observation_window_size=5,
                      use_sliding_observation_window=True, use_next_subgoal_direction=True, final_goal_bonus=200, standing_still_penalty=1,
                      max_episode_steps=100, ignore_traffic_collisions=True), 50, 130, 0.95),
    "aggressive_push": (dict(traffic_density=0.5, conservative_driver_percentage=0, normal_driver_percentage=0, aggressive_driver_percentage=0.6,
                             elderly_driver_percentage=0, reckless_driver_percentage=0.4, random_map_percentage_of_connections=0.9,
                             ignore_traffic_collisions=True, random_map_obstacle_probability=0.7, random_map_traffic_light_probability_weight=8,
                             traffic_light_phases_duration=(2, 1, 6)), 40, 120, 0.97),
    "big_8x8_idle": (dict(random_map_width=8, random_map_height=8, random_map_percentage_of_connections=0.8, traffic_density=0.2,
                          random_map_obstacle_probability=0.5, ignore_traffic_collisions=True), 5, 60, 0.97),
    "rules_without_traffic": (dict(traffic_density=0.1, traffic_rules=[dict(name="always", tile_type="0101", velocity_range=[0.0, 50.0], min_traffic=0,
                                                                            min_matching_traffic=0, maneuvers=[])]), 100, 30, 0.5),
}
