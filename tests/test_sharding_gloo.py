"""Multi-GPU host logic on the CPU: two gloo ranks each own a contiguous shard with GLOBAL env ids;
results must equal the single-process run over the whole range (streams do not depend on the number
of ranks) and the episode statistics must all-reduce to the global ones."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pgtg_b200.distributed import all_reduce_stats, shard

KW = dict(traffic_density=0.05, random_map_obstacle_probability=0.3, seed=77)
TOTAL, TICKS = 96, 12


def _actions():
    return np.random.default_rng(3).integers(0, 9, (TICKS, TOTAL)).astype(np.int32)


def _run(base, n):
    from native_env import NativeAdapter

    env = NativeAdapter("emu", num_envs=n, env_id_base=base, **KW)
    env.reset()
    acts = _actions()[:, base:base + n]
    rewards, obs = [], []
    for t in range(TICKS):
        env.step(acts[t])
        rewards.append(env.reward.copy())
        obs.append(env.obs_map.copy())
    return np.stack(rewards), np.stack(obs), env.stats()


def _worker(rank, world, port, out):
    sys.path[:0] = [os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, n = shard(TOTAL)
    rewards, obs, stats = _run(base, n)
    total = all_reduce_stats(torch.from_numpy(stats.copy()))
    np.savez(os.path.join(out, f"rank{rank}.npz"), rewards=rewards, obs=obs, base=base, n=n, total=total.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_partition():
    assert shard(10, 0, 3) == (0, 3) and shard(10, 1, 3) == (3, 3) and shard(10, 2, 3) == (6, 4)
    assert sum(shard(16 * 1024 * 1024, r, 8)[1] for r in range(8)) == 16 * 1024 * 1024
    with pytest.raises(ValueError):
        shard(2, 0, 3)  # rank 0 of 3 would own no env


def test_two_ranks_equal_one_process(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    ref_rewards, ref_obs, ref_stats = _run(0, TOTAL)
    for r in range(2):
        z = np.load(tmp_path / f"rank{r}.npz")
        b, n = int(z["base"]), int(z["n"])
        assert np.array_equal(z["rewards"], ref_rewards[:, b:b + n])
        assert np.array_equal(z["obs"], ref_obs[:, b:b + n])
        assert np.allclose(z["total"], ref_stats)


def test_numa_binding_helper_is_safe_without_topology(tmp_path):
    """bind_to_gpu_numa_node: cpulist parsing, and no effect (None) where the GPU / its sysfs entry is not visible."""
    import os

    from pgtg_b200.distributed import _parse_cpulist, bind_to_gpu_numa_node

    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0, sysfs=str(tmp_path)) is None
    assert os.sched_getaffinity(0) == before
