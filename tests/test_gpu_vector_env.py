"""`PGTGVectorEnv`: the Gymnasium-facing surface over the CUDA kernels (spaces, reset/step contract,
auto-reset and final_observation, host-buffer steps, state snapshot / set_to_state, statistics)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _env(n=64, **kw):
    from pgtg_b200 import PGTGVectorEnv

    return PGTGVectorEnv(n, device="cuda:0", **kw)


def test_spaces_follow_the_reference():
    env = _env(8, use_next_subgoal_direction=True)
    assert env.single_action_space.n == 9
    sp = env.single_observation_space
    assert list(sp["map"].keys()) == ["walls", "goals", "ice", "broken road", "sand", "traffic", "traffic_light_green",
                                       "traffic_light_yellow", "traffic_light_red"]
    assert tuple(sp["map"]["walls"].shape) == (9, 9)
    assert sp["next_subgoal_direction"].n == 9 and sp["next_subgoal_direction"].start == -1
    env2 = _env(8, use_sliding_observation_window=True, sliding_observation_window_size=5)
    assert tuple(env2.single_observation_space["map"]["walls"].shape) == (11, 11)
    env.close(); env2.close()


def test_reset_and_step_contract():
    import torch

    env = _env(256, traffic_density=0.05, seed=3)
    obs, info = env.reset()
    assert obs["position"].shape == (256, 2) and obs["position"].dtype == torch.int32 and obs["position"].is_cuda
    assert obs["map"]["walls"].shape == (256, 9, 9) and obs["map"]["walls"].dtype == torch.int8
    assert (obs["velocity"] == 0).all()
    a = torch.randint(0, 9, (256,), device="cuda:0")  # int64 actions are accepted as they are
    obs, rew, term, trunc, info = env.step(a)
    assert rew.dtype == torch.float64 and term.dtype == torch.bool and trunc.dtype == torch.bool
    assert set(info) >= {"x", "y", "x_velocity", "y_velocity", "flat_tire", "braking_applied"}
    obs2, *_ = env.step(a.to(torch.int32).cpu().numpy())  # host actions are copied in
    assert obs2["map"]["walls"].max() == 1
    with pytest.raises(ValueError):
        env.step(np.zeros(3, np.int32))
    env.close()


def test_step_before_reset_raises():
    env = _env(4)
    with pytest.raises(RuntimeError):
        env.step(np.zeros(4, np.int32))
    env.close()


def test_same_seed_same_trajectory_and_seed_offsets():
    """tests/test_environment.py:48-126 (seeding determinism), vector form: env i is seeded seed+i."""
    import torch

    acts = torch.randint(0, 9, (20, 128), device="cuda:0", dtype=torch.int32)
    outs = []
    for _ in range(2):
        env = _env(128, traffic_density=0.1, random_map_obstacle_probability=0.5)
        env.reset(seed=1234)
        rs = []
        for t in range(20):
            obs, rew, term, trunc, _ = env.step(acts[t])
            rs.append((obs["map"]["walls"].clone(), obs["map"]["traffic"].clone(), rew.clone(), term.clone()))
        outs.append(rs)
        env.close()
    for a, b in zip(*outs):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    # env 1 of seed s equals env 0 of seed s+1
    e0, e1 = _env(2), _env(2)
    o0, _ = e0.reset(seed=50)
    o1, _ = e1.reset(seed=51)
    assert torch.equal(o0["map"]["walls"][1], o1["map"]["walls"][0]) and torch.equal(o0["position"][1], o1["position"][0])
    e0.close(); e1.close()


def test_step_host_equals_device_step():
    import torch

    kw = dict(traffic_density=0.05, random_map_obstacle_probability=0.3, seed=9)
    a, b = _env(300, **kw), _env(300, **kw)
    a.reset(); b.reset()
    rng = np.random.default_rng(0)
    for t in range(10):
        act = rng.integers(0, 9, 300).astype(np.int32)
        obs, rew, term, trunc, _ = a.step(torch.from_numpy(act).cuda())
        out = b.step_host(act)
        torch.cuda.synchronize()
        assert np.array_equal(out["obs_map"], a._t["obs_map"].cpu().numpy())
        assert np.array_equal(out["reward"], rew.cpu().numpy())
        assert np.array_equal(out["terminated"].astype(bool), term.cpu().numpy())
        assert np.array_equal(out["obs_position"], obs["position"].cpu().numpy())
    a.close(); b.close()


def test_final_observation_and_time_limit():
    import torch

    env = _env(512, max_episode_steps=3, final_observation=True, seed=2)
    env.reset()
    seen_trunc = False
    for t in range(9):
        obs, rew, term, trunc, info = env.step(torch.full((512,), 4, device="cuda:0", dtype=torch.int32))  # stand still: never crashes
        done = term | trunc
        assert torch.equal(info["_final_observation"], done)
        if (t + 1) % 3 == 0:
            assert trunc.all() and not term.any()
            seen_trunc = True
            # the terminal observation is the standing position, the returned one belongs to a new map
            assert (info["final_observation"]["velocity"] == 0).all()
        else:
            assert not done.any()
    assert seen_trunc
    stats = env.episode_stats()
    assert stats["episodes"] == 3 * 512 and stats["truncations"] == 3 * 512 and stats["mean_length"] == 3.0
    env.close()


def test_set_to_state_and_get_state():
    """PGTGEnv.set_to_state (environment.py:1301-1342): position, velocity, flat_tire, cars."""
    env = _env(4, map_plan=dict(width=1, height=1, start=[0, 0, "west"], goal=[0, 0, "east"], map=[[{"exits": [0, 1, 0, 1]}]]),
               traffic_density=0.2, ignore_traffic_collisions=True)
    env.reset()
    st = env.get_state()
    mc = st["cars"].shape[1]
    agent = np.array([[4, 4, 1, 0]] * 4, np.int32)
    cars = np.zeros((4, mc, 7), np.int32)
    cars[:, 0] = [7, 2, 3, 3, 1, 0, 0]  # id 7 at (2, 3) on route east_to_west, profile normal
    obs, _ = env.set_to_state(agent=agent, flat_tire=np.ones(4, np.uint8), num_cars=np.ones(4, np.int32), cars=cars)
    assert obs["position"].cpu().numpy().tolist() == [[4, 4]] * 4 and obs["velocity"].cpu().numpy().tolist() == [[1, 0]] * 4
    assert obs["map"]["traffic"][:, 2, 3].all() and int(obs["map"]["traffic"].sum()) == 4
    st = env.get_state()
    assert st["num_cars"].tolist() == [1] * 4 and st["flat_tire"].tolist() == [1] * 4 and st["cars"][0, 0, :5].tolist() == [7, 2, 3, 3, 1]
    infos = env.get_info_dicts()
    assert infos[0]["cars"][0]["route"] == "east_to_west" and infos[0]["flat_tire"] is True
    env.close()


def test_traffic_rule_management():
    env = _env(8, traffic_density=0.1)
    assert env.remove_traffic_rule("t_intersection_brake") is True
    assert env.remove_traffic_rule("t_intersection_brake") is False
    with pytest.raises(ValueError):
        env.add_traffic_rule(dict(name="four_way_intersection_brake", tile_type="1111", velocity_range=[0.5, 10], min_traffic=1,
                                  min_matching_traffic=1, maneuvers=[]))
    env.add_traffic_rule(dict(name="always", tile_type="0101", velocity_range=[0.0, 99.0], min_traffic=0, min_matching_traffic=0, maneuvers=[]))
    env.reset()
    env.step(np.full(8, 7, np.int32))
    env.close()


def test_conformance_draws_constructor():
    """`conformance_draws=` switches the kernels to the recorded np_random tape."""
    import torch

    import parity

    tr = parity.load_trace([p for p in parity.golden_traces() if p.endswith("trace_default.npz")][0])
    kw = parity.trace_kwargs(tr)
    n = kw.pop("num_envs")
    kw.pop("max_episode_steps")
    env = _env(n, conformance_draws=(tr["tape_values"], tr["tape_tags"], tr["tape_offsets"]), **kw)
    obs, _ = env.reset()
    assert np.array_equal(obs["position"].cpu().numpy(), tr["obs_position"][0])
    for t in range(10):
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(tr["actions"][t]).cuda())
        assert np.array_equal(rew.cpu().numpy(), tr["reward"][t]) and np.array_equal(obs["map"]["walls"].cpu().numpy(), tr["obs_map"][t + 1][:, 0])
    env.close()


@pytest.mark.gpu
def test_clone_and_light_step_leave_the_env_unchanged():
    """PGTGEnv.light_step (environment.py:1283-1299): a copy steps, the original does not; a clone is an exact twin."""
    import torch
    from pgtg_b200 import PGTGVectorEnv

    n = 2048
    env = PGTGVectorEnv(n, seed=21, traffic_density=0.1, random_map_obstacle_probability=0.3)
    env.reset()
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    acts = [torch.randint(0, 9, (n,), device="cuda", dtype=torch.int32, generator=g) for _ in range(6)]
    for a in acts[:3]:
        env.step(a)
    before = env.get_state()
    obs_l, rew_l, term_l, _, _ = env.light_step(acts[3])
    obs_l = {"position": obs_l["position"].clone(), "walls": obs_l["map"]["walls"].clone(), "traffic": obs_l["map"]["traffic"].clone()}
    rew_l, term_l = rew_l.clone(), term_l.clone()
    after = env.get_state()
    for k in ("agent", "cars", "num_cars", "tiles", "elapsed"):
        assert (before[k] == after[k]).all(), k  # the original did not move
    twin = env.clone()
    obs, rew, term, _, _ = env.step(acts[3])       # now the real step: identical to what light_step predicted
    assert torch.equal(rew, rew_l) and torch.equal(term, term_l) and torch.equal(obs["position"], obs_l["position"])
    assert torch.equal(obs["map"]["walls"], obs_l["walls"]) and torch.equal(obs["map"]["traffic"], obs_l["traffic"])
    obs_t, rew_t, _, _, _ = twin.step(acts[3])
    assert torch.equal(rew, rew_t) and torch.equal(obs["map"]["traffic"], obs_t["map"]["traffic"])
    env.close(); twin.close()


@pytest.mark.gpu
def test_evaluate_matches_the_reference_evaluator_bookkeeping():
    import torch
    from pgtg_b200 import PGTGVectorEnv

    n = 4096
    env = PGTGVectorEnv(n, seed=5, random_map_obstacle_probability=0.3)
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    def policy(obs):  # mostly "no acceleration": many agents idle on the start line until the step cap
        a = torch.randint(0, 9, (n,), device="cuda", dtype=torch.int32, generator=g)
        return torch.where(torch.rand(n, device="cuda", generator=g) < 0.9, torch.full_like(a, 4), a)

    mean_ret, (terminated, truncated, over, negative) = env.evaluate(policy, number=5000, max_steps=9, GAMMA=0.95)
    assert terminated + over >= 5000 and truncated == 0 and over > 0 and 0 < negative <= terminated + over
    assert -100.0 <= mean_ret <= 100.0
    assert env.episode_stats()["episodes"] >= 5000
    env.close()


@pytest.mark.gpu
def test_scalar_seed_is_offset_by_the_shard_base():
    """reset(seed=S) on a shard seeds env i with S + env_id_base + i: two half shards == one full handle."""
    import torch
    from pgtg_b200 import PGTGVectorEnv

    kw = dict(rng_mode="numpy", traffic_density=0.05)
    full = PGTGVectorEnv(256, **kw)
    lo, hi = PGTGVectorEnv(128, env_id_base=0, **kw), PGTGVectorEnv(128, env_id_base=128, **kw)
    of, _ = full.reset(seed=77)
    ol, _ = lo.reset(seed=77)
    oh, _ = hi.reset(seed=77)
    a = torch.randint(0, 9, (256,), device="cuda", dtype=torch.int32)
    for _ in range(5):
        of, rf, _, _, _ = full.step(a)
        ol, rl, _, _, _ = lo.step(a[:128])
        oh, rh, _, _, _ = hi.step(a[128:])
        assert torch.equal(rf, torch.cat([rl, rh])) and torch.equal(of["map"]["walls"], torch.cat([ol["map"]["walls"], oh["map"]["walls"]]))
        assert torch.equal(of["map"]["traffic"], torch.cat([ol["map"]["traffic"], oh["map"]["traffic"]]))
    assert not torch.equal(ol["map"]["walls"], oh["map"]["walls"])
    full.close(); lo.close(); hi.close()


@pytest.mark.gpu
def test_info_is_a_mapping_and_info_dicts_carry_the_reference_fields():
    import torch
    from pgtg_b200 import PGTGVectorEnv

    env = PGTGVectorEnv(64, seed=1, traffic_density=0.2, final_observation=True)
    env.reset()
    _, _, _, _, info = env.step(torch.full((64,), 4, device="cuda", dtype=torch.int32))
    copied = dict(info)
    assert copied["flat_tire"] is not None and copied["braking_applied"].dtype == torch.bool and copied["_final_observation"] is not None
    d = env.get_info_dicts()[0]
    for key in ("x", "y", "x_velocity", "y_velocity", "flat_tire", "current_tile_type", "cars", "driver_profile_stats", "traffic_rules"):
        assert key in d
    assert len(d["current_tile_type"]) == 4 and set(d["current_tile_type"]) <= {"0", "1"}
    assert d["driver_profile_stats"]["total_cars"] == len(d["cars"]) == sum(d["driver_profile_stats"]["counts"].values())
    assert d["traffic_rules"]["active_rules"] == ["four_way_intersection_brake", "t_intersection_brake"]
    with pytest.raises(RuntimeError, match="outside 0..8"):
        env.step(torch.full((64,), 11, device="cuda", dtype=torch.int32))
        env.episode_stats()
    env.close()
