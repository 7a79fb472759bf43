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


_ALIGN = 256  # every tensor starts on a 256-byte boundary of the flat buffer (TMA / vector loads need 16)


def _layout(meta) -> Tuple[list, int]:
    offs, total = [], 0
    for _, shape, dt in meta:
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=getattr(torch, dt)).element_size()
        offs.append((total, nbytes))
        total += (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
    return offs, total


def flatten_weights(weights: Dict[str, torch.Tensor], device: torch.device):
    """-> (flat uint8 buffer on `device` holding every tensor back to back, meta list, dict of typed VIEWS into it)."""
    meta = [(k, tuple(v.shape), str(v.dtype).replace("torch.", "")) for k, v in weights.items()]
    offs, total = _layout(meta)
    flat = torch.empty(total, dtype=torch.uint8, device=device)
    for (k, _, _), (o, nb) in zip(meta, offs):
        flat[o:o + nb].copy_(weights[k].contiguous().view(-1).view(torch.uint8))
    return flat, meta, views_of(flat, meta)


def views_of(flat: torch.Tensor, meta) -> Dict[str, torch.Tensor]:
    offs, _ = _layout(meta)
    return {k: flat[o:o + nb].view(getattr(torch, dt)).view(shape) for (k, shape, dt), (o, nb) in zip(meta, offs)}


def broadcast_weights(weights: Optional[Dict[str, torch.Tensor]], device: torch.device, src: int = 0) -> Dict[str, torch.Tensor]:
    """Rank `src` holds the packed engine weights.  They travel as ONE flat buffer in ONE collective (a single ncclBroadcast of
    3.3 GB for large-v3 -- round 1 issued ~1000 per-tensor broadcasts); every rank, `src` included, then uses typed views into
    its copy of that buffer, so no second copy of the weights exists anywhere.  The only collective of the whole job."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert weights is not None
        return weights
    rank = dist.get_rank()
    flat, meta = None, None
    if rank == src:
        flat, meta, _ = flatten_weights(weights, device)
    box = [meta]
    dist.broadcast_object_list(box, src=src)  # names / shapes / dtypes: a few KB over the host
    meta = box[0]
    if rank != src:
        flat = torch.empty(_layout(meta)[1], dtype=torch.uint8, device=device)
    dist.broadcast(flat, src=src)
    return views_of(flat, meta)


def gather_results(local: list, dst: int = 0) -> Optional[List[list]]:
    """Small Python results (token ids, words) travel over host memory, not the GPU fabric."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(local, out, dst=dst)
    return out
