"""Data-parallel plumbing: one process per GPU, independent chunks / streams per rank, ONE weight broadcast at init and
no collective on the data path.  torch.distributed is used for plumbing only (NCCL on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Round-robin ownership: item i belongs to rank i % world (equal-length chunks need no bucketing)."""
    return range(rank, n_items, world)


def stream_owner(stream_id: int, world: int) -> int:
    """Streams are sticky: their rolling buffers are host state of one rank."""
    return stream_id % world


def broadcast_weights(weights: Optional[Dict[str, torch.Tensor]], device: torch.device, src: int = 0) -> Dict[str, torch.Tensor]:
    """Rank `src` holds the packed engine weights; every other rank allocates the same tensors and receives them."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert weights is not None
        return weights
    rank = dist.get_rank()
    meta = [(k, tuple(v.shape), str(v.dtype).replace("torch.", "")) for k, v in weights.items()] if rank == src else None
    box = [meta]
    dist.broadcast_object_list(box, src=src)
    meta = box[0]
    if rank != src:
        weights = {k: torch.empty(shape, dtype=getattr(torch, dt), device=device) for k, shape, dt in meta}
    for k, _, _ in meta:
        dist.broadcast(weights[k], src=src)
    return weights


def gather_results(local: list, dst: int = 0) -> Optional[List[list]]:
    """Small Python results (token ids, words) travel over host memory, not the GPU fabric."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(local, out, dst=dst)
    return out
