"""Host-side mirror of PGTGEnv.__init__ (pgtg_b200/config.py): argument checks, derived numbers,
observation layout, direction table."""
import math
import warnings

import numpy as np
import pytest

from pgtg_b200 import config as cfg


def test_defaults_match_reference_constructor():
    hc = cfg.make_config()
    c = hc.pod
    assert (c.map_w, c.map_h) == (4, 4)
    assert c.edges_to_keep == round(48 * 0.5) and c.border_connections == round(14 * 0.5)
    assert (c.start_x, c.start_y, c.start_dir) == (0, 3, 3) and (c.goal_x, c.goal_y, c.goal_dir) == (3, 0, 1)
    assert hc.observation_keys == cfg.DEFAULT_FEATURES and hc.window == 9
    assert (c.sum_subgoals_reward, c.crash_penalty, c.traffic_light_violation_penalty) == (100.0, 100.0, 50.0)
    assert (c.light_green, c.light_yellow, c.light_red) == (10, 3, 10)
    assert c.num_rules == 2 and c.rules[0].tile_type == 15 and c.rules[1].tile_type == 7


def test_bankers_rounding_of_edge_counts():
    # Python round() is banker's rounding (map_generator.py:242, 362-364)
    hc = cfg.make_config(random_map_width=3, random_map_height=2, random_map_percentage_of_connections=0.25)
    assert hc.pod.edges_to_keep == round(14 * 0.25) == 4  # 3.5 -> 4
    hc = cfg.make_config(random_map_width=2, random_map_height=2, random_map_percentage_of_connections=0.25)
    assert hc.pod.border_connections == round(6 * 0.25) == 2  # 1.5 -> 2


def test_derived_driver_thresholds_are_host_doubles():
    c = cfg.make_config().pod
    assert list(c.drv_patience_threshold) == [0.9 * 10, 0.7 * 10, 0.3 * 10, 0.95 * 10, 0.1 * 10]
    assert list(c.drv_push_probability) == [1.0 - 0.9, 1.0 - 0.7, 1.0 - 0.3, 1.0 - 0.95, 1.0 - 0.1]
    assert c.drv_push_probability[0] == 0.09999999999999998 and c.drv_push_probability[1] == 0.30000000000000004


def test_profile_cdf_is_numpy_cumsum_normalised():
    c = cfg.make_config(conservative_driver_percentage=1, normal_driver_percentage=2, aggressive_driver_percentage=3,
                        elderly_driver_percentage=4, reckless_driver_percentage=0).pod
    p = np.array([1, 2, 3, 4, 0], float) / 10
    cdf = p.cumsum()
    cdf /= cdf[-1]
    assert list(c.profile_cdf) == cdf.tolist()
    c = cfg.make_config(conservative_driver_percentage=0, normal_driver_percentage=0, aggressive_driver_percentage=0,
                        elderly_driver_percentage=0, reckless_driver_percentage=0).pod
    assert list(c.profile_cdf) == [0.0, 1.0, 1.0, 1.0, 1.0]  # everything NORMAL (environment.py:506-508)


@pytest.mark.parametrize("kw,msg", [
    (dict(random_map_start_position=(1, 1)), "start_position must specify a tile on the map border."),
    (dict(random_map_goal_position=(2, 2, "east")), "goal_position must specify a tile on the map border."),
    (dict(random_map_start_position=(0, 0, "south")), "The direction in start_position is not a map border."),
    (dict(random_map_start_position=(0, 0, "west"), random_map_goal_position=(0, 0, "west")), "can't be the same tile and direction"),
    (dict(random_map_minimum_distance_between_start_and_goal=2), "can only be used if start_position and goal_position are 'random'"),
    (dict(random_map_start_position="random", random_map_goal_position="random", random_map_minimum_distance_between_start_and_goal=7),
     "can't be larger than width + height - 2"),
])
def test_generate_map_argument_errors(kw, msg):
    """The ValueErrors of generate_map (map_generator.py:92-154), raised at construction."""
    import re

    with pytest.raises(ValueError, match=re.escape(msg)):
        cfg.make_config(**kw)


def test_unobservable_obstacle_warnings():
    """environment.py:366-412."""
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        cfg.make_config(random_map_obstacle_probability=0.5, features_to_include_in_observation=["walls", "goals"], traffic_density=0.1)
    texts = " ".join(str(x.message) for x in w)
    for word in ("ice obstacle", "broken road obstacle", "sand obstacle", "green traffic lights", "yellow traffic lights", "red traffic lights", "Traffic is generated"):
        assert word in texts
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        cfg.make_config()
    assert not w


def test_observation_layout_quirks():
    # the default traffic_light_* names are literal features no square carries -> zero planes (A.3-1)
    lay = dict(cfg.observation_layout(cfg.DEFAULT_FEATURES))
    assert lay["traffic_light_green"] == lay["traffic_light_red"] == cfg.CH_ZERO
    assert lay["walls"] == cfg.CH_WALLS and lay["goals"] == cfg.CH_GOALS and lay["traffic"] == cfg.CH_TRAFFIC
    # a literal "traffic_light" entry expands to the three phase planes (environment.py:1411-1439)
    lay = cfg.observation_layout(["walls", "traffic_light"])
    assert [k for k, _ in lay] == ["walls", "traffic_light_green", "traffic_light_yellow", "traffic_light_red"]
    assert [v for _, v in lay][1:] == [cfg.CH_LIGHT_GREEN, cfg.CH_LIGHT_YELLOW, cfg.CH_LIGHT_RED]
    # ... unless the generic loop overwrites them afterwards (:1441-1445)
    lay = dict(cfg.observation_layout(["traffic_light", "traffic_light_green"]))
    assert lay["traffic_light_green"] == cfg.CH_ZERO and lay["traffic_light_red"] == cfg.CH_LIGHT_RED
    with pytest.raises(NotImplementedError):
        cfg.observation_layout(["car_lane all up"])


def test_direction_lut_matches_exact_sector_geometry():
    """The octant of atan2(dy, dx) never sits on a sector boundary for integer offsets (tan(pi/8) is
    irrational), so the float classification must equal the exact integer one."""
    R = 40
    lut = cfg.direction_lut(R)
    for dy in range(-R, R + 1):
        for dx in range(-R, R + 1):
            if dx == 0 and dy == 0:
                continue
            ax, ay = abs(dx), abs(dy)
            # |angle to the x axis| < pi/8  <=>  ay < ax * tan(pi/8)  <=>  (ay + ax)^2 < 2 ax^2
            near_x = (ay + ax) ** 2 < 2 * ax * ax
            near_y = (ax + ay) ** 2 < 2 * ay * ay
            if near_x:
                want = 2 if dx > 0 else 6
            elif near_y:
                want = 4 if dy > 0 else 0
            else:
                want = {(1, 1): 3, (-1, 1): 5, (-1, -1): 7, (1, -1): 1}[(int(math.copysign(1, dx)), int(math.copysign(1, dy)))]
            assert lut[dy + R, dx + R] & 7 == want, (dx, dy)


def test_fixed_map_plan_packing():
    plan = dict(width=2, height=1, start=[0, 0, "west"], goal=[1, 0, "east"],
                map=[[{"exits": [0, 1, 0, 1]}, {"exits": [1, 1, 0, 1], "obstacle_type": "sand", "obstacle_mask": "left_half"}]])
    hc = cfg.make_config(map_plan=plan)
    tiles = hc.map_plan.packed_tiles()
    assert (tiles[0].exits, tiles[0].obstacle_type) == (0b1010, 0)
    assert (tiles[1].exits, tiles[1].obstacle_type, tiles[1].obstacle_mask) == (0b1011, 3, 6)
    with pytest.raises(KeyError):  # the stale example maps have no start/goal (MapPlan.from_dict)
        cfg.MapPlan.from_dict({"width": 1, "height": 1, "map": [[{"exits": [0, 1, 0, 1]}]]})


def test_rule_flattening():
    r = cfg.rule_to_pod(cfg.DEFAULT_RULES[1])
    w = np.array([[r.weight[a][i] for i in range(20)] for a in range(6)])
    from pgtg_b200._names import ROUTE_NAMES

    assert w[cfg.AGENT_DIRECTIONS.index("south_to_north"), ROUTE_NAMES.index("west_to_east")] == 1
    assert w[cfg.AGENT_DIRECTIONS.index("west_to_east"), ROUTE_NAMES.index("south_to_north")] == 1
    assert w.sum() == 3
