"""Multi-GPU sharding: envs are independent units, so the env index range is split contiguously over
the ranks with NO per-step communication. Every rank keeps GLOBAL env ids
(`env_id_base`), so per-env random streams -- and therefore every result -- do not depend on the
number of GPUs. The only collective is the optional all-reduce (sum) of the 8 episode-statistics
doubles."""
from __future__ import annotations

import os


def shard(num_envs: int, rank: int | None = None, world_size: int | None = None) -> tuple[int, int]:
    """-> (env_id_base, local_num_envs) of this rank for `num_envs` global envs."""
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world_size = int(os.environ.get("WORLD_SIZE", "1")) if world_size is None else world_size
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base = num_envs * rank // world_size
    end = num_envs * (rank + 1) // world_size
    if end <= base:
        raise ValueError("more ranks than envs")
    return base, end - base


def all_reduce_stats(stats):
    """Sum the 8 episode-statistics doubles over all ranks (NCCL for CUDA tensors, gloo for CPU)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats)
    return stats


def make_sharded(total_envs: int, device=None, **kwargs):
    """A `PGTGVectorEnv` over this rank's shard of `total_envs` global envs (one process per GPU,
    launched with torchrun)."""
    import torch

    from .vector_env import PGTGVectorEnv

    base, n = shard(total_envs)
    if device is None:
        device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    return PGTGVectorEnv(n, device=device, env_id_base=base, **kwargs)
