"""Multi-GPU sharding of the two hot paths: one process per GPU, no collective on the data path.

Frames (FFT) and channels (IIR) are independent objects in the reference (one complex_array per call,
fft.h:258-360; one casc_2o_iir object per channel, casc_2o_iir.h:8-20), so rank r simply owns a
contiguous block of them, with its share of the coefficient / history bank.  The only communication is
the OPTIONAL result gather below (NCCL over NVLink when the tensors live on GPUs, gloo in the CPU
tests); throughput numbers never include it.
"""
from __future__ import annotations

from typing import List, Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `total` units owned by `rank`; the first total % world ranks get one more."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError((rank, world))
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total: int, world: int) -> List[int]:
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def gather_shards(local, total: int, dst: int = 0, group=None):
    """Optional result gather: every rank passes its shard (first dimension = its units); rank `dst`
    gets the concatenation in rank order, the others None.  Uses torch.distributed (backend of the
    initialised process group: nccl for CUDA tensors, gloo for CPU tensors)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} units, expected {sizes[rank]}")
    # all_gather needs equal shapes: pad the short shards by one unit
    width = max(sizes)
    padded = local
    if local.shape[0] < width:
        pad = torch.zeros((width - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    if rank != dst:
        return None
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
