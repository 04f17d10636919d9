"""CPU: the N>1 host path — contiguous channel blocks per rank and the gather of the per-channel BER
counters (the only collective of the hot path) — exercised with world_size 2 and 3 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qpsk_modulator_demodulator_b200 import shard


def test_channel_ranges_partition_exactly():
    for channels in (0, 1, 7, 1024, 16384, 16385):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.channel_range(r, world, channels) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == channels
            for (a0, a1), (b0, b1) in zip(blocks[:-1], blocks[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
            for c in range(0, channels, max(1, channels // 50)):
                r = shard.owner_of(c, world, channels)
                assert blocks[r][0] <= c < blocks[r][1]
    assert shard.channel_range(3, 8, 16384) == (6144, 8192)          # 2048 per GPU
    with pytest.raises(ValueError):
        shard.channel_range(2, 2, 10)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, channels, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = shard.channel_range(rank, world, channels)
    # each rank "measures" its own block: errors = channel index, bits = 1000 + channel index
    idx = torch.arange(first, last, dtype=torch.int32)
    local = torch.stack([idx, idx + 1000], dim=1)
    allc = shard.gather_counters(local, dist)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                          # the bench's max-over-ranks timing
    np.save(os.path.join(out_dir, f"r{rank}.npy"), allc.numpy())
    assert t.item() == world
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,channels", [(2, 10), (2, 7), (3, 8)])
def test_gather_counters_gloo(tmp_path, world, channels):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, channels, str(tmp_path)), nprocs=world, join=True)
    want = np.stack([np.arange(channels), np.arange(channels) + 1000], axis=1)
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npy")
        assert got.shape == (channels, 2) and np.array_equal(got, want)


def test_gather_counters_single_process():
    local = torch.tensor([[1, 2], [3, 4]], dtype=torch.int32)
    assert torch.equal(shard.gather_counters(local, None), local)
