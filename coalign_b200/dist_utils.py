"""Process-group plumbing for the data-parallel scene sharding (one process per GPU; no data-path collective).

Scenes are independent units (fusion is intra-scene, /root/reference/opencood/models/fuse_modules/fusion_in_one.py:123-133),
so ranks only ever exchange scalars: a barrier and a MAX all-reduce of elapsed time (bench) / result counters.
Works with nccl (GPU) and gloo (CPU tests)."""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def world_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: str, device_id=None):
    rank, _local, world = world_info()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device_id} if (device_id is not None and backend == "nccl") else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def shard_scenes(scene_ids: Sequence[int], rank: int, world: int) -> List[int]:
    """Round-robin scene assignment, the DistributedSampler pattern of train_ddp.py:46-55 without padding."""
    return [s for i, s in enumerate(scene_ids) if i % world == rank]


def max_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
