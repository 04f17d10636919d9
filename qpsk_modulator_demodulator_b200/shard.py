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


class CounterComm:
    """The same gather through the C ABI (qpsk_comm_* / qpsk_ber_gather, csrc/comm.cu): an NCCL communicator owned by
    libqpskcuda.so, what a C# host would use.  `exchange(id_bytes_or_None) -> id_bytes` ships rank 0's 128-byte id to the
    other ranks (torch.distributed broadcast in bench.py; any transport in a C# host)."""

    def __init__(self, world: int, rank: int, exchange):
        import ctypes as C

        import numpy as np

        from . import _native as N
        self._N, self._C, self._np = N, C, np
        self.world, self.rank = world, rank
        ident = np.zeros(128, np.uint8)
        if rank == 0:
            N.check(N.lib().qpsk_comm_unique_id(ident.ctypes.data, 128))
        ident = np.frombuffer(exchange(ident.tobytes() if rank == 0 else None), np.uint8).copy()
        self._h = C.c_void_p()
        N.check(N.lib().qpsk_comm_create(ident.ctypes.data, world, rank, C.byref(self._h)))

    def info(self) -> dict:
        C = self._C
        n, r, v = C.c_int(), C.c_int(), C.c_int()
        self._N.check(self._N.lib().qpsk_comm_info(self._h, C.byref(n), C.byref(r), C.byref(v)))
        return {"n_ranks": n.value, "rank": r.value, "nccl_version": v.value}

    def gather(self, d_counters: int, channels_local: int, channels_total: int):
        """d_counters: device pointer to uint32 {errors, bits}[channels_local] (complete).  Returns the [channels_total, 2]
        table in channel order (numpy, identical on every rank)."""
        np = self._np
        counts = [channel_range(r, self.world, channels_total) for r in range(self.world)]
        width = max(b - a for a, b in counts)
        out = np.zeros((self.world, max(width, 1), 2), np.uint32)
        self._N.check(self._N.lib().qpsk_ber_gather(self._h, d_counters, channels_local, max(width, 1), out.ctypes.data))
        return np.concatenate([out[r, : b - a] for r, (a, b) in enumerate(counts)], axis=0)

    def close(self):
        if self._h:
            self._N.lib().qpsk_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
