"""Host-side logic of the multi-GPU path on CPU: world_size 2 over gloo.  Each rank owns a contiguous
block of frames / channels (no data-path collective); the optional gather reassembles them in order.
The per-rank compute is stood in for by the CPU oracle -- the sharding, not the kernels, is under test."""
import os
import socket

import numpy as np
import pytest

from simpledsp_b200.shard import shard_range, shard_sizes


def test_c_abi_shard_range_matches_the_python_sharder():
    import ctypes as C

    from simpledsp_b200 import _capi as K

    for total in (0, 1, 7, 64, 65536, 262144):
        for world in (1, 2, 3, 8):
            for rank in range(world):
                first, count = C.c_size_t(), C.c_size_t()
                K.check(K.lib().sdsp_b200_shard_range(total, rank, world, C.byref(first), C.byref(count)))
                lo, hi = shard_range(total, rank, world)
                assert (first.value, first.value + count.value) == (lo, hi)
    assert K.lib().sdsp_b200_shard_range(10, 2, 2, C.byref(first), C.byref(count)) == K.lib().sdsp_b200_shard_range(10, -1, 2, C.byref(first), C.byref(count)) != 0


def test_shard_ranges_partition_exactly():
    for total in (0, 1, 7, 8, 65536, 16385):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(ranges[:-1], ranges[1:]))
            sizes = shard_sizes(total, world)
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, frames, n, out_dir):
    import torch
    import torch.distributed as dist

    from oracle import oracle as O
    from simpledsp_b200.shard import gather_shards, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)  # every rank can regenerate the whole synthetic batch
    x = rng.standard_normal((frames, n)) + 1j * rng.standard_normal((frames, n))
    lo, hi = shard_range(frames, rank, world)
    mine = O.fft(x[lo:hi], 4)  # stand-in for the per-GPU transform of this rank's frames
    got = gather_shards(torch.from_numpy(mine), frames, dst=0)
    # channels: ragged split of an IIR bank
    ch = 5
    sig = rng.standard_normal((ch, 256))
    clo, chi = shard_range(ch, rank, world)
    ft, f0 = np.where(np.arange(ch) % 2 == 0, 1, 2), np.geomspace(1e3, 2e4, ch)
    part = O.iir_bank_port(sig[clo:chi], ft[clo:chi], f0[clo:chi], 100e3) if chi > clo else np.zeros((0, 256))
    got2 = gather_shards(torch.from_numpy(part), ch, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "fft.npy"), got.numpy())
        np.save(os.path.join(out_dir, "iir.npy"), got2.numpy())
    else:
        assert got is None and got2 is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_gather(tmp_path):
    import torch.multiprocessing as mp

    from oracle import oracle as O

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    frames, n = 7, 64
    mp.spawn(_worker, args=(2, port, frames, n, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(42)
    x = rng.standard_normal((frames, n)) + 1j * rng.standard_normal((frames, n))
    assert np.array_equal(np.load(tmp_path / "fft.npy"), O.fft(x, 4))
    ch = 5
    sig = rng.standard_normal((ch, 256))
    ft, f0 = np.where(np.arange(ch) % 2 == 0, 1, 2), np.geomspace(1e3, 2e4, ch)
    assert np.array_equal(np.load(tmp_path / "iir.npy"), O.iir_bank_port(sig, ft, f0, 100e3))
