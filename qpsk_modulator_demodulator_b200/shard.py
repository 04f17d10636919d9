"""Multi-GPU partitioning of the hot path: independent channels / bursts / frames are
split into contiguous blocks, one block per rank, with no data-path collective.  The only exchange is
the gather of the per-channel BER counters at the end (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def channel_range(rank: int, world: int, channels: int) -> Tuple[int, int]:
    """Contiguous block [first, last) of `channels` owned by `rank`: channel c lives on rank floor(c*world/channels)."""
    if world <= 0 or not (0 <= rank < world) or channels < 0:
        raise ValueError("bad rank/world/channels")
    first = -(-rank * channels // world)          # ceil(rank*channels/world)
    last = -(-(rank + 1) * channels // world)
    return first, last


def owner_of(channel: int, world: int, channels: int) -> int:
    if not (0 <= channel < channels):
        raise ValueError("channel out of range")
    return channel * world // channels


def gather_counters(local_counters, dist=None, counts=None):
    """All-gather the per-channel {errors, bits} counters ([n_local, 2] integer tensor) of every rank
    into one [channels, 2] tensor in channel order.  `dist` is torch.distributed (initialised) or None
    for a single process.  `counts` = per-rank channel counts when the blocks are ragged."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_counters.clone()
    world = dist.get_world_size()
    n_local = local_counters.shape[0]
    if counts is None:
        sizes = torch.tensor([n_local], dtype=torch.int64, device=local_counters.device)
        all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(all_sizes, sizes)
        counts = [int(s.item()) for s in all_sizes]
    width = max(counts)
    padded = torch.zeros((width, 2), dtype=local_counters.dtype, device=local_counters.device)
    padded[:n_local] = local_counters
    parts = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:n] for p, n in zip(parts, counts)], dim=0)
