"""GPU parity of the batched FFT, through the C ABI.  Run with -m gpu on a B200."""
import numpy as np
import pytest

import simpledsp_b200 as S
from oracle import oracle as O
from simpledsp_b200 import _capi as K
from tests.util import FFT_TOL, f32_noise, ref_vectors, rel_l2

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps
PREC = {"f64": (K.F64, np.complex128), "f32": (K.F32, np.complex64)}


def oracle_fft(x, inverse=False):
    x = np.asarray(x, dtype=np.complex128)
    n = x.shape[-1]
    if n < 4:
        return np.fft.ifft(x) if inverse else np.fft.fft(x)
    radix = 4 if (n.bit_length() - 1) % 2 == 0 else 2
    return O.fft(x, radix, inverse)


# ------------------------------------------------------------------ integer path: bit exact
@pytest.mark.parametrize("n", [2, 4, 8, 16, 64, 256, 1024, 2048, 4096, 16384, 65536])
def test_digit_reversal_tables_bit_exact(n):
    for base in (2, 4):
        if base == 4 and (n.bit_length() - 1) % 2:
            continue
        if n >= 4:
            assert np.array_equal(S.digit_reverse_table(n, base, half_table=True), O.swap_lookup(n, base)), (n, base)
        if n >= base:
            assert np.array_equal(S.digit_reverse_table(n, base), O.digit_reverse(n, base)), (n, base)


@pytest.mark.parametrize("n,base", [(64, 2), (64, 4), (1024, 4), (2048, 2), (4096, 4)])
def test_digit_reverse_permute_matches_swap_sweep(n, base):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((5, n)) + 1j * rng.standard_normal((5, n))).astype(np.complex64)
    rev = O.digit_reverse(n, base)
    want = np.empty_like(x)
    want[:, rev] = x
    got = S.digit_reverse_permute(x.copy(), base)
    assert np.array_equal(got, want)
    assert np.array_equal(S.digit_reverse_permute(got.copy(), base), x)  # involution


# ------------------------------------------------------------------ parity against the oracle
@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("lg", range(1, 19))
def test_fft_matches_oracle_all_sizes(lg, prec):
    """Single-CTA kernels up to 2^14 (f32) / 2^13 (f64); above that the multi-pass path (n = n1 x 256
    through an L2-resident scratch), up to 2^18 (f32) / 2^17 (f64)."""
    n = 1 << lg
    if prec == "f64" and lg > 17:
        pytest.skip("f64 2^18 is beyond the multi-pass path")
    code, dt = PREC[prec]
    rng = np.random.default_rng(100 + lg)
    frames = 7 if n <= 4096 else 3
    x = f32_noise(rng, (frames, n)) + 1j * f32_noise(rng, (frames, n))
    for inv in (False, True):
        ref = oracle_fft(x, inv)
        for radix in (2, 4):
            if radix == 4 and lg % 2:
                continue
            got = S.FftPlan(n, radix, code, K.REVERSE if inv else K.FORWARD)(np.ascontiguousarray(x.astype(dt)))
            assert rel_l2(got, ref) <= FFT_TOL[prec], (n, prec, inv, radix)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_fft_matches_reference_vectors(prec):
    code, dt = PREC[prec]
    z = ref_vectors()
    for n in (4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
        x = z[f"fft_in_{n}"]
        for radix in (2, 4):
            for d, inv in (("fwd", False), ("inv", True)):
                key = f"fft_r{radix}_{d}_{n}"
                if key not in z:
                    continue
                got = (S.fft_radix4 if radix == 4 else S.fft_radix2)(np.ascontiguousarray(x.astype(dt)), inverse=inv)
                assert rel_l2(got, z[key]) <= FFT_TOL[prec], key


@pytest.mark.parametrize("fn", [S.fft_radix2, S.fft_radix4])
def test_reference_known_answer_tests(fn):
    # reference test/testFFT.cpp:17-68, 127-178 with the reference's own bound 4*N*eps
    N, n = 64, 7
    i = np.arange(N)
    s = np.cos(n * 2 * np.pi * i / N).astype(np.complex128)
    Sx = np.zeros(N, dtype=np.complex128)
    Sx[n] = Sx[N - n] = N / 2
    tol = 4 * N * EPS
    assert np.abs(fn(s.copy()) - Sx).max() < tol
    assert np.abs(fn(Sx.copy(), inverse=True) - s).max() < tol
    s2 = np.cos(n * 2 * np.pi * i / N + np.pi / 2).astype(np.complex128)
    S2 = np.zeros(N, dtype=np.complex128)
    S2[n], S2[N - n] = 1j * N / 2, -1j * N / 2
    assert np.abs(fn(s2.copy()) - S2).max() < tol
    # linearity, test/testFFT.cpp:70-125, 180-235
    N, fs, a1, a2 = 256, 8000.0, 1.5, 2.5
    i = np.arange(N)
    x1 = np.sin(2 * np.pi * 1000.0 / fs * i).astype(np.complex128)
    x2 = np.sin(2 * np.pi * 500.0 / fs * i).astype(np.complex128)
    lhs = fn(a1 * x1 + a2 * x2)
    rhs = a1 * fn(x1.copy()) + a2 * fn(x2.copy())
    assert np.abs(lhs - rhs).max() < 4 * N * EPS


@pytest.mark.parametrize("frames", [1, 2, 3, 15, 16, 17, 255, 257, 1000])
def test_ragged_frame_counts(frames):
    rng = np.random.default_rng(frames)
    for n in (16, 64, 1024):
        x = (f32_noise(rng, (frames, n)) + 1j * f32_noise(rng, (frames, n)))
        got = S.fft_radix2(np.ascontiguousarray(x.astype(np.complex64)))
        assert rel_l2(got, np.fft.fft(x)) <= FFT_TOL["f32"]
    assert S.fft_radix2(np.zeros((0, 64), dtype=np.complex64)).shape == (0, 64)  # empty batch is a no-op


# ------------------------------------------------------------------ device-resident data, full size
def test_device_pointer_path_and_full_size_properties():
    torch = pytest.importorskip("torch")
    n, frames = 4096, 65536  # BASELINE config 2
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn(frames, n, 2, device="cuda", generator=g, dtype=torch.float32)
    xc = torch.view_as_complex(x)
    fwd = S.FftPlan(n, 4, K.F32, K.FORWARD)
    inv = S.FftPlan(n, 4, K.F32, K.REVERSE)
    y = xc.clone()
    fwd(y)
    torch.cuda.synchronize()
    # sampled frames against the oracle
    idx = torch.linspace(0, frames - 1, 256).long()  # parity on >= 256 frames spread across the batch
    ref = oracle_fft(xc[idx].cpu().numpy())
    assert rel_l2(y[idx].cpu().numpy(), ref) <= FFT_TOL["f32"]
    # Parseval on every frame: sum |X|^2 = N sum |x|^2
    e_t = (xc.abs() ** 2).sum(dim=1, dtype=torch.float64)
    e_f = (y.abs() ** 2).sum(dim=1, dtype=torch.float64) / n
    assert float(((e_t - e_f).abs() / e_t).max()) < 1e-5
    # round trip on every frame
    inv(y)
    torch.cuda.synchronize()
    err = (y - xc).abs().pow(2).sum(dim=1).sqrt() / xc.abs().pow(2).sum(dim=1).sqrt()
    assert float(err.max()) <= 2 * FFT_TOL["f32"]


def test_device_f64_batch_linearity():
    torch = pytest.importorskip("torch")
    n, frames = 4096, 2048
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.view_as_complex(torch.randn(frames, n, 2, device="cuda", generator=g, dtype=torch.float64))
    b = torch.view_as_complex(torch.randn(frames, n, 2, device="cuda", generator=g, dtype=torch.float64))
    plan = S.FftPlan(n, 4, K.F64, K.FORWARD)
    lhs = plan((1.5 * a + 2.5 * b).contiguous())
    rhs = 1.5 * plan(a.clone()) + 2.5 * plan(b.clone())
    torch.cuda.synchronize()
    err = (lhs - rhs).abs().pow(2).sum(dim=1).sqrt() / rhs.abs().pow(2).sum(dim=1).sqrt()
    assert float(err.max()) <= FFT_TOL["f64"]
    ref = oracle_fft(a[:4].cpu().numpy())
    assert rel_l2(plan(a[:4].clone()).cpu().numpy(), ref) <= FFT_TOL["f64"]


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_large_frames_many_frames_and_unsupported_sizes(prec):
    """BASELINE config 5's transform size (65536) over more frames than one scratch slab holds."""
    torch = pytest.importorskip("torch")
    code, dt = PREC[prec]
    n, frames = 65536, 150 if prec == "f32" else 70  # more frames than the scratch ring holds (64 / 32)
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.view_as_complex(torch.randn(frames, n, 2, device="cuda", generator=g, dtype=torch.float32 if prec == "f32" else torch.float64))
    fwd, inv = S.FftPlan(n, 4, code, K.FORWARD), S.FftPlan(n, 4, code, K.REVERSE)
    assert "single pass over HBM" in fwd.describe() and fwd.launches(frames) == 1
    y = x.clone()
    fwd(y)
    torch.cuda.synchronize()
    idx = [0, 1, frames // 2, frames - 1]
    ref = oracle_fft(x[idx].cpu().numpy())
    assert rel_l2(y[idx].cpu().numpy(), ref) <= FFT_TOL[prec]
    inv(y)
    torch.cuda.synchronize()
    err = (y - x).abs().pow(2).sum(dim=1).sqrt() / x.abs().pow(2).sum(dim=1).sqrt()
    assert float(err.max()) <= 2 * FFT_TOL[prec]
    with pytest.raises(S.SdspError) as e:
        S.FftPlan(1 << 20, 2, code, K.FORWARD)
    assert e.value.status == K.ERR_UNSUPPORTED


def _check_conjugate_symmetry(y):
    """Spectra of real 65536-point frames from fft_real64k_kernel: bins with k mod 256 in 1 .. 127 are computed, their mirrors
    N - k are stored as the conjugate of the same registers (equal to the bit); the bins of rows 0 and 128 (k mod 256 in
    {0, 128}) mirror into their own row and are each computed: symmetric to rounding."""
    import torch

    n = y.shape[1]
    k = torch.arange(1, n, device=y.device)
    mirror = torch.conj(y[:, n - k]).resolve_conj()
    exact = (k % 256 != 0) & (k % 256 != 128)
    assert torch.equal(y[:, k[exact]], mirror[:, exact])
    rest = ~exact
    scale = y.abs().amax(dim=1, keepdim=True)
    assert float(((y[:, k[rest]] - mirror[:, rest]).abs() / scale).max()) <= 1e-5


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("n", [64, 1024, 4096, 16384, 32768, 65536, 131072])
def test_real_input_frames_match_oracle(n, prec):
    """sdsp_b200_fft_exec_real: real frames in, spectra out -- what the reference's callers do by filling only the
    real part of a complex_array (test/testFFT.cpp:24, :86), folded into the first load."""
    torch = pytest.importorskip("torch")
    code, dt = PREC[prec]
    frames = 37 if n <= 4096 else 9
    rng = np.random.default_rng(n)
    x = rng.standard_normal((frames, n)).astype(np.float32)
    ref = oracle_fft(x.astype(np.complex128))
    plan = S.FftPlan(n, 4 if (n.bit_length() - 1) % 2 == 0 else 2, code, K.FORWARD)
    xd = torch.from_numpy(x.astype(np.float32 if prec == "f32" else np.float64)).cuda()
    yd = plan.real(xd)
    torch.cuda.synchronize()
    assert rel_l2(yd.cpu().numpy(), ref) <= FFT_TOL[prec]
    assert torch.equal(xd.cpu(), torch.from_numpy(x.astype(np.float32 if prec == "f32" else np.float64)))  # input untouched
    # host buffers through the staging path give the same bits as the device call
    yh = plan.real(x.astype(np.float32 if prec == "f32" else np.float64))
    assert np.array_equal(yh, yd.cpu().numpy())
    # and the same bits as the complex entry point fed (x, 0) -- except forward fp32 frames of 65536 points, which take the
    # half-work kernel that exploits the conjugate symmetry of a real frame's spectrum (fft_real64k_kernel): same transform,
    # different rounding, and X[N - k] == conj(X[k]) to the bit
    z = torch.complex(xd, torch.zeros_like(xd)).contiguous()
    plan(z)
    torch.cuda.synchronize()
    if n == 65536 and prec == "f32":
        assert rel_l2(yd.cpu().numpy(), z.cpu().numpy()) <= FFT_TOL[prec]
        _check_conjugate_symmetry(yd)
    else:
        assert torch.equal(z, yd)


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072])
def test_half_spectrum_of_real_frames_matches_oracle(n, prec):
    """sdsp_b200_fft_exec_r2c: bins 0 .. n/2 of the spectrum of real frames.  Checked against the oracle's transform of (x, 0) --
    the reference's calling convention, test/testFFT.cpp:24, :86 -- at the FFT tolerance, device and host buffers (same bits),
    odd frame counts (partial groups), input untouched."""
    torch = pytest.importorskip("torch")
    code, dt = PREC[prec]
    # (sizes without a direct kernel -- f64 from 32768 points, f32 from 131072 -- go through the full spectrum: same checks)
    rdt = np.float32 if prec == "f32" else np.float64
    frames = 261 if n <= 256 else 37 if n <= 4096 else 9
    rng = np.random.default_rng(n + 1)
    x = rng.standard_normal((frames, n)).astype(np.float32).astype(rdt)
    ref = oracle_fft(x.astype(np.complex128))[:, : n // 2 + 1]
    plan = S.FftPlan(n, 2, code, K.FORWARD)
    xd = torch.from_numpy(x).cuda()
    yd = plan.half_spectrum(xd)
    torch.cuda.synchronize()
    assert tuple(yd.shape) == (frames, n // 2 + 1)
    got = yd.cpu().numpy()
    assert rel_l2(got, ref) <= FFT_TOL[prec]
    assert torch.equal(xd.cpu(), torch.from_numpy(x))
    # the purely real bins of a real signal
    scale = np.abs(ref).max(axis=1)
    assert np.all(np.abs(got[:, 0].imag) <= 1e-6 * scale) and np.all(np.abs(got[:, n // 2].imag) <= 1e-6 * scale)
    if n <= (65536 if prec == "f32" else 16384):  # the direct kernels form the bin n/2 as a real number
        assert np.all(got[:, n // 2].imag == 0)
    yh = plan.half_spectrum(x)
    assert np.array_equal(yh, got)
    # one frame, and a frame count that leaves the last group of a CTA partly empty
    y1 = plan.half_spectrum(xd[:1].contiguous())
    torch.cuda.synchronize()
    assert np.array_equal(y1.cpu().numpy(), got[:1])


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072])
def test_real_frames_back_from_half_spectra(n, prec):
    """sdsp_b200_fft_exec_c2r (reverse plans): bins 0 .. n/2 in, real frames out with reverse_fft's 1/n (fft.h:128-132).  Against the
    oracle's REVERSE transform of the full, mirror-completed spectrum; the round trip through exec_r2c; host buffers = same bits."""
    torch = pytest.importorskip("torch")
    code, dt = PREC[prec]
    frames = 261 if n <= 256 else 37 if n <= 4096 else 9
    rng = np.random.default_rng(n + 2)
    m = n // 2
    half = (rng.standard_normal((frames, m + 1)) + 1j * rng.standard_normal((frames, m + 1))).astype(np.complex64).astype(dt)
    half[:, 0] = half[:, 0].real  # what a real signal's spectrum looks like
    half[:, m] = half[:, m].real
    full = np.concatenate([half, np.conj(half[:, m - 1:0:-1])], axis=1).astype(np.complex128)
    ref = oracle_fft(full, inverse=True)
    assert np.abs(ref.imag).max() <= 1e-9 * np.abs(ref.real).max()
    inv = S.FftPlan(n, 2, code, K.REVERSE)
    hd = torch.from_numpy(half).cuda()
    xd = inv.real_from_half_spectrum(hd)
    torch.cuda.synchronize()
    got = xd.cpu().numpy()
    assert got.shape == (frames, n)
    assert rel_l2(got, ref.real) <= FFT_TOL[prec]
    assert torch.equal(hd.cpu(), torch.from_numpy(half))
    assert np.array_equal(inv.real_from_half_spectrum(half), got)
    # round trip
    fwd = S.FftPlan(n, 2, code, K.FORWARD)
    again = fwd.half_spectrum(xd)
    torch.cuda.synchronize()
    assert rel_l2(again.cpu().numpy(), half.astype(np.complex128)) <= 2 * FFT_TOL[prec]
    with pytest.raises(RuntimeError):
        fwd.real_from_half_spectrum(hd)  # needs a reverse plan


def test_half_spectrum_full_size_and_errors():
    """Config-2-sized batch of real frames (65536 x 4096, fp32) against torch.fft.rfft, and the argument checks."""
    torch = pytest.importorskip("torch")
    n, frames = 4096, 65536
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(frames, n, device="cuda", generator=g, dtype=torch.float32)
    plan = S.FftPlan(n, 4, K.F32, K.FORWARD)
    y = plan.half_spectrum(x)
    ref = torch.fft.rfft(x.double(), dim=1)
    err = (y.to(torch.complex128) - ref).abs().pow(2).sum(dim=1).sqrt() / ref.abs().pow(2).sum(dim=1).sqrt()
    assert float(err.max()) <= FFT_TOL["f32"]
    idx = [0, 1, 32767, 65535]
    assert rel_l2(y[idx].cpu().numpy(), oracle_fft(x[idx].cpu().numpy().astype(np.complex128))[:, : n // 2 + 1]) <= FFT_TOL["f32"]
    with pytest.raises(RuntimeError):
        S.FftPlan(n, 4, K.F32, K.REVERSE).half_spectrum(x[:2].contiguous())
    with pytest.raises(RuntimeError):
        S.FftPlan(2, 2, K.F32, K.FORWARD).half_spectrum(np.zeros((3, 2), dtype=np.float32))
    # an empty batch is a no-op; a real frame that does not start on a pair boundary is refused (the kernel reads it in pairs)
    plan.exec_r2c_ptr(x.data_ptr(), y.data_ptr(), 0, K.PTR_DEVICE, None)
    with pytest.raises(RuntimeError):
        plan.exec_r2c_ptr(x.data_ptr() + 4, y.data_ptr(), 1, K.PTR_DEVICE, torch.cuda.current_stream().cuda_stream)
    with pytest.raises(RuntimeError):
        plan.exec_r2c_ptr(x.data_ptr(), y.data_ptr(), 1, 7, None)
    torch.cuda.synchronize()


@pytest.mark.parametrize("frames", [1, 2, 31, 32, 33, 47, 48, 49, 63, 64, 65, 87, 88, 89, 95, 96, 97, 127, 128, 129, 175, 176, 177, 191, 192, 193, 200, 365, 401])
def test_fused_65536_kernel_at_the_edges_of_its_work_queue(frames):
    """The 65536-point kernel orders column and row tiles through a queue with a 48-frame lag and a 96-frame scratch ring
    (fp32; 32 / 64 in the variant without the data-mover warp; 96 / 192 in the real-input kernel): frame counts below, at and
    just past those boundaries, forward and reverse, complex and real input."""
    torch = pytest.importorskip("torch")
    n = 65536
    g = torch.Generator(device="cuda").manual_seed(frames)
    x = torch.view_as_complex(torch.randn(frames, n, 2, device="cuda", generator=g, dtype=torch.float32))
    fwd, inv = S.FftPlan(n, 4, K.F32, K.FORWARD), S.FftPlan(n, 4, K.F32, K.REVERSE)
    y = x.clone()
    fwd(y)
    torch.cuda.synchronize()
    idx = sorted({0, frames // 2, frames - 1, min(31, frames - 1), min(48, frames - 1), min(64, frames - 1), min(96, frames - 1)})
    ref = oracle_fft(x[idx].cpu().numpy())
    assert rel_l2(y[idx].cpu().numpy(), ref) <= FFT_TOL["f32"]
    # every frame: Parseval, then the round trip
    e_t = (x.abs() ** 2).sum(dim=1, dtype=torch.float64)
    e_f = (y.abs() ** 2).sum(dim=1, dtype=torch.float64) / n
    assert float(((e_t - e_f).abs() / e_t).max()) < 1e-5
    inv(y)
    torch.cuda.synchronize()
    err = (y - x).abs().pow(2).sum(dim=1).sqrt() / x.abs().pow(2).sum(dim=1).sqrt()
    assert float(err.max()) <= 2 * FFT_TOL["f32"]
    # real input, out of place, forward: the half-work kernel (its own queue: 8 + 9 tiles per frame) against the complex entry point
    # fed (x, 0), every frame; reverse plans keep the complex kernels: same bits
    xr = x.real.contiguous()
    z = torch.complex(xr, torch.zeros_like(xr)).contiguous()
    fwd(z)
    out = fwd.real(xr)
    torch.cuda.synchronize()
    err = (z - out).abs().pow(2).sum(dim=1).sqrt() / z.abs().pow(2).sum(dim=1).sqrt()
    assert float(err.max()) <= FFT_TOL["f32"]
    _check_conjugate_symmetry(out)
    # half spectra out: the same kernel, which only stops writing the upper half: same bits
    h = fwd.half_spectrum(xr)
    torch.cuda.synchronize()
    assert torch.equal(h, out[:, : n // 2 + 1])
    zi = torch.complex(xr, torch.zeros_like(xr)).contiguous()
    inv(zi)
    outi = inv.real(xr)
    torch.cuda.synchronize()
    assert torch.equal(zi, outi)


_OTHER_QUEUE_SNIPPET = r"""
import sys, numpy as np, torch
import simpledsp_b200 as S
from simpledsp_b200 import _capi as K
n, frames = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator(device="cuda").manual_seed(7)
x = torch.view_as_complex(torch.randn(frames, n, 2, device="cuda", generator=g, dtype=torch.float32))
plan = S.FftPlan(n, 4 if (n.bit_length() - 1) % 2 == 0 else 2, K.F32, K.FORWARD)
print("PLAN", plan.describe())
plan(x)
torch.cuda.synchronize()
np.save(sys.argv[3], x.cpu().numpy())
"""


@pytest.mark.parametrize("n", [32768, 65536])
def test_fused_queue_with_and_without_the_data_mover_warp_agree(n, tmp_path):
    """fp32 frames of 2^15 / 2^16 points run the work queue fed by a TMA data-mover warp (two tile slots that double as exchange
    buffers) by default; SDSP_B200_FFT_FUSED_TMA=1 selects the one-slot data-mover kernel, =0 the variant whose compute threads
    load their own tiles.  Same arithmetic in the same order: all three must give the same bits (each in its own process: the
    choice is read once per process)."""
    import os
    import subprocess
    import sys

    from tests.util import ROOT

    outs = {}
    for flag in ("2", "1", "0"):
        path = str(tmp_path / f"out_{flag}.npy")
        env = dict(os.environ, SDSP_B200_FFT_FUSED_TMA=flag, PYTHONPATH=ROOT)
        r = subprocess.run([sys.executable, "-c", _OTHER_QUEUE_SNIPPET, str(n), "97", path], env=env, capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert r.returncode == 0, r.stderr[-2000:]
        assert ("data-mover" in r.stdout) == (flag != "0"), r.stdout
        outs[flag] = np.load(path)
    assert np.array_equal(outs["1"], outs["0"]) and np.array_equal(outs["2"], outs["0"])
    g = np.random.default_rng(0).integers(0, 97, 3)
    # and against the oracle, for a few frames, using the same generator seed as the snippet
    torch = pytest.importorskip("torch")
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = torch.view_as_complex(torch.randn(97, n, 2, device="cuda", generator=gen, dtype=torch.float32)).cpu().numpy()
    assert rel_l2(outs["1"][g], oracle_fft(x[g])) <= FFT_TOL["f32"]


def test_shutdown_gives_the_l2_carve_out_back_and_the_library_stays_usable():
    """The large-frame kernels pin part of their scratch ring in L2 (a persisting access-policy window; the carve-out is
    device-wide state set at the first such launch).  sdsp_b200_shutdown resets it; the next launch sets it up again and
    gives the same bits."""
    torch = pytest.importorskip("torch")
    n, frames = 65536, 130
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(frames, n, device="cuda", generator=g, dtype=torch.float32)
    plan = S.FftPlan(n, 4, K.F32, K.FORWARD)
    a = plan.real(x)
    torch.cuda.synchronize()
    K.check(K.lib().sdsp_b200_shutdown())
    b = plan.real(x)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert rel_l2(a[:2].cpu().numpy(), oracle_fft(x[:2].cpu().numpy().astype(np.complex128))) <= FFT_TOL["f32"]


@pytest.mark.parametrize("n,frames", [(65536, 300), (4096, 6000), (1024, 3)])
def test_host_buffers_in_slabs_match_the_device_resident_call(n, frames):
    """Host-pointer calls stage the batch through device memory in 64 MB slabs on two alternating streams
    (simpledsp_b200/csrc/host_stage.h).  The transforms of consecutive slabs share the plan's scratch ring and counters
    for n >= 32768, so they are chained by events; the result must be the bits of one device-resident call.  (300 frames
    of 65536 points = 157 MB = three slabs; three frames of 1024 points take the pinned bounce buffer.)"""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(n + frames)
    x = (rng.standard_normal((frames, n)) + 1j * rng.standard_normal((frames, n))).astype(np.complex64)
    plan = S.FftPlan(n, 4, K.F32, K.FORWARD)
    d = torch.from_numpy(x).cuda()
    plan(d)
    torch.cuda.synchronize()
    want = d.cpu().numpy()
    for _ in range(2):  # twice: the second call reuses the plan's streams, events and staging memory
        got = x.copy()
        plan(got)
        assert np.array_equal(got, want)
    # real frames through host buffers against the device-resident real-input call (same kernel: same bits)
    real = np.ascontiguousarray(x.real)
    spec = plan.real(real)
    d2 = plan.real(torch.from_numpy(real).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(spec, d2.cpu().numpy())
