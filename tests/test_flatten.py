"""Flattened observation: the kernel's float32 [N, D] equals a literal restatement of
gymnasium 0.28.1 `flatten(Dict)` on the dict observation (sorted keys, MultiBinary as is, Discrete and
MultiDiscrete one-hot, Box as is)."""
import ctypes as C

import numpy as np
import pytest


def gymnasium_flatten(keys, obs_map, pos, vel, nsd, use_nsd):
    """flatten(Dict) of gymnasium 0.28.1 for one env: OrderedDict(sorted(...)) at every Dict level."""
    parts = []
    for k in sorted(keys):  # "map" sub-dict
        parts.append(np.asarray(obs_map[keys.index(k)]).flatten())
    if use_nsd:
        oh = np.zeros(9, np.int64); oh[nsd + 1] = 1  # Discrete(9, start=-1)
        parts.append(oh)
    ohp = np.zeros(18, np.int64); ohp[pos[0]] = 1; ohp[9 + pos[1]] = 1  # MultiDiscrete([9, 9])
    parts.append(ohp)
    parts.append(np.asarray(vel))
    return np.concatenate(parts).astype(np.float32)


def _check(env, use_nsd):
    ptr, dim = env.raw.flatten()
    n = env.N
    flat = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n, dim)).copy() if env.backend == "emu" else None
    if flat is None:
        import torch

        torch.cuda.synchronize()
        flat = torch.from_dlpack(env.raw.dlpack_capsule("obs_flat")).cpu().numpy()
    keys = env.hc.observation_keys
    m, pos, vel, nsd = env.obs_map, env.obs_position, env.obs_velocity, env.obs_nsd
    for i in range(n):
        want = gymnasium_flatten(keys, m[i], pos[i], vel[i], int(nsd[i]), use_nsd)
        assert flat.shape[1] == want.size
        assert np.array_equal(flat[i], want), i


@pytest.mark.parametrize("use_nsd", [False, True])
def test_flatten_emulated(use_nsd):
    from native_env import NativeAdapter

    env = NativeAdapter("emu", num_envs=40, seed=3, traffic_density=0.1, random_map_obstacle_probability=0.5, use_next_subgoal_direction=use_nsd)
    env.reset()
    rng = np.random.default_rng(0)
    for _ in range(6):
        env.step(rng.integers(0, 9, 40).astype(np.int32))
        _check(env, use_nsd)
    env.close()


@pytest.mark.gpu
def test_flatten_cuda_and_save_map(tmp_path):
    import json

    import torch

    from native_env import NativeAdapter
    from pgtg_b200 import PGTGVectorEnv

    env = NativeAdapter("cuda", num_envs=300, seed=3, traffic_density=0.1, random_map_obstacle_probability=0.5, use_next_subgoal_direction=True)
    env.reset()
    rng = np.random.default_rng(0)
    for _ in range(4):
        env.step(rng.integers(0, 9, 300).astype(np.int32))
        _check(env, True)
    env.close()
    venv = PGTGVectorEnv(8, device="cuda:0", random_map_obstacle_probability=1.0, seed=1)
    venv.reset()
    f = venv.flat_observation()
    assert f.shape == (8, 9 * 81 + 18 + 2) and f.dtype == torch.float32
    venv.save_map(str(tmp_path / "m"), env_index=3)
    plan = json.load(open(tmp_path / "m.json"))
    assert plan["width"] == 4 and plan["start"] == [0, 3, "west"] and plan["goal"] == [3, 0, "east"]
    assert any("obstacle_type" in t for row in plan["map"] for t in row)
    # the saved plan loads back as a fixed map and shows the same first observation planes
    again = PGTGVectorEnv(1, device="cuda:0", map_plan=plan)
    o2, _ = again.reset()
    assert torch.equal(o2["map"]["walls"][0], venv._observation()["map"]["walls"][3])
    venv.close(); again.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [0, 2, 5, 8])
def test_flatten_cuda_sliding_windows(k):
    """the flat index space of the flatten kernel (cell -> plane by a multiply-high) over window sizes 1, 5, 11 and 17
    (the position space stays MultiDiscrete([9, 9]), environment.py:428, so gymnasium itself cannot flatten windows beyond k = 8)"""
    from native_env import NativeAdapter

    n = 70
    env = NativeAdapter("cuda", num_envs=n, seed=5, traffic_density=0.1, random_map_obstacle_probability=0.5, use_next_subgoal_direction=True,
                        use_sliding_observation_window=True, sliding_observation_window_size=k)
    env.reset()
    rng = np.random.default_rng(1)
    for _ in range(3):
        env.step(rng.integers(0, 9, n).astype(np.int32))
        _check(env, True)
    env.close()
