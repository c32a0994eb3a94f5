"""Two ranks on two GPUs over NCCL: each rank runs the CUDA path on its contiguous shard of frames / channels (no
collective on the data path), then the optional gather (NCCL over NVLink) reassembles the result on rank 0, which
checks it against the oracle.  Skipped on a single-GPU box (the CPU twin of this test is tests/test_multi_rank_cpu.py)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    import simpledsp_b200 as S
    from simpledsp_b200 import _capi as K
    from simpledsp_b200.shard import gather_shards, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(42)  # every rank regenerates the whole synthetic batch and keeps its block
    frames, n = 37, 4096
    x = (rng.standard_normal((frames, n)) + 1j * rng.standard_normal((frames, n))).astype(np.complex64)
    lo, hi = shard_range(frames, rank, world)
    mine = torch.from_numpy(x[lo:hi]).cuda(rank)
    S.FftPlan(n, 4, K.F32, K.FORWARD, rank)(mine)
    got = gather_shards(mine, frames, dst=0)
    # channels: ragged split of an IIR bank, each rank holds its share of the coefficients
    ch, ns, fs = 70, 5000, 100e3
    sig = rng.standard_normal((ch, ns)).astype(np.float32)
    ft, f0 = np.where(np.arange(ch) % 2 == 0, 1, 2), np.geomspace(1e3, 2e4, ch)
    clo, chi = shard_range(ch, rank, world)
    coef = [S.design(int(t), 4, float(f), fs) for t, f in zip(ft[clo:chi], f0[clo:chi])]
    bank = S.IirBank(4, chi - clo, K.F32, K.NUM_GENERIC, rank)
    bank.set_coeffs(np.array([c[0] for c in coef]), np.array([c[1] for c in coef]), np.array([c[2] for c in coef]))
    part = torch.from_numpy(sig[clo:chi]).cuda(rank)
    bank.process(part)
    got2 = gather_shards(part, ch, dst=0)
    torch.cuda.synchronize()
    if rank == 0:
        np.save(os.path.join(out_dir, "fft.npy"), got.cpu().numpy())
        np.save(os.path.join(out_dir, "iir.npy"), got2.cpu().numpy())
    else:
        assert got is None and got2 is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus_shard_compute_gather(tmp_path):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from oracle import oracle as O
    from tests.util import FFT_TOL, IIR_TOL, peak_rel, rel_l2

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(42)
    frames, n = 37, 4096
    x = (rng.standard_normal((frames, n)) + 1j * rng.standard_normal((frames, n))).astype(np.complex64)
    assert rel_l2(np.load(tmp_path / "fft.npy"), O.fft(x.astype(np.complex128), 4)) <= FFT_TOL["f32"]
    ch, ns = 70, 5000
    sig = rng.standard_normal((ch, ns)).astype(np.float32)
    ft, f0 = np.where(np.arange(ch) % 2 == 0, 1, 2), np.geomspace(1e3, 2e4, ch)
    assert peak_rel(np.load(tmp_path / "iir.npy"), O.iir_bank_port(sig.astype(np.float64), ft, f0, 100e3)) <= IIR_TOL["f32"]
