"""Data-parallel host logic (new functionality: the reference is single-device, runners.py:193).

One process per GPU.  The global batch is split contiguously by rank; every rank scales its
per-sample terms by 1/global_batch, so the SUM all-reduce of the flat gradient buffer (NCCL over
NVLink, inside the native step) equals the single-device gradient of the global mean.  Adam then
runs redundantly on every rank on identical inputs, which keeps the replicas bit-identical without
a broadcast."""
from __future__ import annotations

import os
from typing import Tuple


def shard_bounds(global_batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of this rank's rows; the first `global_batch % world_size` ranks
    take one extra row."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(global_batch, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def env_world() -> Tuple[int, int, int]:
    """(world_size, rank, local_rank) from the torchrun environment (1, 0, 0 when absent)."""
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: str = "nccl"):
    """torch.distributed is used only as plumbing: rendezvous, the NCCL unique-id broadcast,
    barriers and max-over-ranks timing.  The gradient all-reduce itself is issued by the native
    library on its own NCCL communicator (gmvae_allreduce_grads)."""
    import torch
    import torch.distributed as dist
    world, rank, local = env_world()
    if world == 1 or dist.is_initialized():
        return world, rank, local
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    return world, rank, local
