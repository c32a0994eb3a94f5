"""GPU parity of the cascaded-biquad IIR banks, through the C ABI.  Run with -m gpu on a B200."""
import numpy as np
import pytest

import simpledsp_b200 as S
from oracle import oracle as O
from simpledsp_b200 import _capi as K
from tests.util import IIR_FIXTURE_F32, IIR_GOLDEN_ABS, IIR_TOL, f32_noise, golden_impulses, peak_rel, ref_vectors, rel_l2

pytestmark = pytest.mark.gpu
PREC = {"f64": (K.F64, np.float64), "f32": (K.F32, np.float32)}
CLS = {0: S.casc_2o_iir, 1: S.casc_2o_iir_lp, 2: S.casc_2o_iir_hp, 3: S.casc_2o_iir_bp}


def _design(f, ftype, f0, fs, q, gain=1.0):
    if ftype == 1:
        f.set_lp_coeff(f0, fs, gain)
    elif ftype == 2:
        f.set_hp_coeff(f0, fs, gain)
    else:
        f.set_bp_coeff(f0, fs, q, gain)


@pytest.mark.parametrize("num_fixed", [False, True])
def test_golden_impulse_responses_fp32_bank_and_exact_checkpoint(num_fixed):
    """The nine golden fixtures through the fp32 KERNELS (a bank of nine channels, one per fixture; generic and
    fixed numerators): within IIR_FIXTURE_F32 of the reference's fp64 impulse responses -- LPimpulse (f0/fs = 0.005)
    included, which a direct-form fp32 recurrence misses by 1e-4.  Then reference test/testIIR.cpp:61-75 on
    the bank: 32-sample blocks reproduce the whole-buffer run bit for bit, also through a checkpoint
    (get_state + get_state_diff -> another bank -> set_state + set_state_diff) in the middle of the stream."""
    cases = list(golden_impulses())
    n = cases[0][5]
    assert all(c[5] == n for c in cases)
    for kind in ((1, 2, 3) if num_fixed else (0,)):
        sel = [c for c in cases if not num_fixed or c[1] == kind]
        coef = [S.design(c[1], 4, c[3], c[2], c[4]) for c in sel]
        want = np.array([c[6] for c in sel])

        def make():
            bank = S.IirBank(4, len(sel), K.F32, kind)
            bank.set_coeffs(np.array([c[0] for c in coef]), None if num_fixed else np.array([c[1] for c in coef]),
                            np.array([c[2] for c in coef]))
            return bank

        x = np.zeros((len(sel), n), dtype=np.float32)
        x[:, 0] = 1.0
        y = make().process(x.copy())
        for i, c in enumerate(sel):
            assert peak_rel(y[i], want[i]) <= IIR_FIXTURE_F32, (c[0], kind)
        bank, y2 = make(), x.copy()
        for lo in range(0, n, 32):
            if lo == 512:  # checkpoint / resume on a fresh bank
                mem, dif = bank.get_state(), bank.get_state_diff()
                bank = make()
                bank.set_state(mem)
                bank.set_state_diff(dif)
            part = np.ascontiguousarray(y2[:, lo:lo + 32])
            bank.process(part)
            y2[:, lo:lo + 32] = part
        assert np.array_equal(y, y2), kind
        # without the differences the resumed stream is still within rounding of the uninterrupted one
        bank, y3 = make(), x.copy()
        bank.process(y3[:, :512].copy())
        mem = bank.get_state()
        bank = make()
        bank.set_state(mem)
        tail = np.ascontiguousarray(y3[:, 512:])
        bank.process(tail)
        err = np.abs(tail - y[:, 512:]).max(axis=1) / np.abs(want).max(axis=1)
        assert err.max() <= IIR_FIXTURE_F32, kind


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_golden_impulse_responses_and_block_streaming(prec):
    # reference test/testIIR.cpp:32-77 (generic) and :223-430 (fixed numerators), through the single-object classes: fp64
    # ARITHMETIC for either sample type, as in the reference (casc_2o_iir.h:13-18, 45-71) -- float samples are rounded once
    code, dt = PREC[prec]
    for name, ftype, fs, f0, q, n, h in golden_impulses():
        for num in (0, ftype):
            f = CLS[num](4, code)
            _design(f, ftype, f0, fs, q)
            f2 = f.copy()
            x = np.zeros(n, dtype=dt)
            x[0] = 1.0
            y = f.process(x.copy())
            if prec == "f64":
                assert np.abs(y - h).max() < IIR_GOLDEN_ABS, (name, num)
            assert peak_rel(y, h) <= (IIR_TOL["f64"] if prec == "f64" else 1e-7), (name, num)
            y2 = x.copy()
            for i in range(0, n, 32):  # 32-sample blocks and the 8-sample tail: exact equality
                f2.process(y2[i:i + 32])
            assert np.array_equal(y, y2), (name, num)


def test_gain_and_preload():
    # reference test/testIIR.cpp:79-218
    fs, f0, q = 100e3, 10e3, 1.1
    for ftype in (1, 2, 3):
        x = np.zeros(1024)
        x[0] = 1.0
        f1, f2 = S.casc_2o_iir(4), S.casc_2o_iir(4)
        _design(f1, ftype, f0, fs, q, 1.0)
        _design(f2, ftype, f0, fs, q, 2.0)
        assert np.abs(2.0 * f1.process(x.copy()) - f2.process(x.copy())).max() < 1e-12
        f = S.casc_2o_iir(4)
        _design(f, ftype, f0, fs, q)
        f.preload_filter(10.0)
        y = f.process(np.full(1024, 10.0))
        assert np.abs(y - (10.0 if ftype == 1 else 0.0)).max() < 1e-12
        assert peak_rel(y, ref_vectors()[f"iir_preload_t{ftype}"]) <= IIR_TOL["f64"] or ftype != 1


def test_matches_reference_vectors_all_section_counts():
    z = ref_vectors()
    x = z["iir_in"][0]
    for key in z["iir_cases"]:
        key = str(key)
        _, m, kind, t, fq = key.split("_")
        sections, ftype, f0 = int(m[1:]), int(t[1:]), float(fq[1:])
        num = {"generic": 0, "lp": 1, "hp": 2, "bp": 3}[kind]
        for prec in ("f64", "f32"):
            code, dt = PREC[prec]
            f = CLS[num](sections, code)
            _design(f, ftype, f0, 100e3, 1.1, 1.0 if f0 > 1e3 else 0.75)
            y = x.astype(dt)
            f.process(y[:300])
            f.process(y[300:])
            # single-object path: fp64 arithmetic, float samples rounded once on the way out (half an ulp of the peak)
            assert peak_rel(y, z[key]) <= (IIR_TOL["f64"] if prec == "f64" else 1e-7), (key, prec)


def _bank_case(n_channels, n_samples, prec, sections=4, seed=0):
    code, dt = PREC[prec]
    rng = np.random.default_rng(seed)
    fs = 100e3
    ftype = np.where(np.arange(n_channels) % 2 == 0, 1, 2)
    f0 = np.geomspace(1e3, 20e3, n_channels)
    gains, bs, as_ = [], [], []
    for c in range(n_channels):
        g, b, a = S.design(int(ftype[c]), sections, float(f0[c]), fs)
        gains.append(g)
        bs.append(b)
        as_.append(a)
    x = f32_noise(rng, (n_channels, n_samples))
    ref = O.iir_bank_port(x, ftype, f0, fs, sections=sections)
    bank = S.IirBank(sections, n_channels, code)
    bank.set_coeffs(np.array(gains), np.array(bs), np.array(as_))
    return bank, x.astype(dt), ref


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("n_channels,n_samples", [(1, 1000), (31, 257), (32, 64), (33, 4096), (100, 1000), (257, 31), (64, 1)])
def test_bank_matches_oracle_ragged_shapes(n_channels, n_samples, prec):
    bank, x, ref = _bank_case(n_channels, n_samples, prec, seed=n_channels)
    y = bank.process(x.copy())
    assert peak_rel(y, ref) <= IIR_TOL[prec]
    # streaming: same data in three uneven calls on a fresh state gives the identical bits
    bank.reset_state()
    y2 = x.copy()
    cuts = [0, n_samples // 3, n_samples // 3 + 7 if n_samples > 30 else n_samples // 3, n_samples]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        if hi > lo:
            part = np.ascontiguousarray(y2[:, lo:hi])
            bank.process(part)
            y2[:, lo:hi] = part
    assert np.array_equal(y, y2)


def test_bank_state_roundtrip_and_device_pointer_path():
    torch = pytest.importorskip("torch")
    bank, x, ref = _bank_case(96, 2048, "f32", seed=3)
    xd = torch.from_numpy(x).cuda()
    bank.process(xd[:, :1024].contiguous())  # first half on the device
    st = bank.get_state()
    assert st.shape == (96, 5, 2)
    bank2, _, _ = _bank_case(96, 2048, "f32", seed=3)
    bank2.set_state(st)
    second = xd[:, 1024:].contiguous()
    bank2.process(second)
    torch.cuda.synchronize()
    assert peak_rel(second.cpu().numpy(), ref[:, 1024:]) <= IIR_TOL["f32"]
    # strided device view: channel_stride > n_samples
    bank.reset_state()
    wide = torch.zeros(96, 2048 + 64, device="cuda", dtype=torch.float32)
    wide[:, :2048] = torch.from_numpy(x).cuda()
    bank.process_ptr(wide.data_ptr(), 2048, 2048 + 64, K.PTR_DEVICE, K.IIR_AUTO, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert peak_rel(wide[:, :2048].cpu().numpy(), ref) <= IIR_TOL["f32"]
    assert float(wide[:, 2048:].abs().max()) == 0.0  # padding untouched


@pytest.mark.parametrize("sections", [2, 6, 8])
def test_bank_other_section_counts(sections):
    bank, x, ref = _bank_case(40, 512, "f64", sections=sections, seed=sections)
    assert peak_rel(bank.process(x.copy()), ref) <= IIR_TOL["f64"]


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_tma_and_generic_kernels_agree_bit_for_bit(prec):
    """An unaligned base pointer forces the generic kernel; an aligned one takes the TMA kernel.  Same
    arithmetic, same bits -- which is what lets SDSP_B200_IIR_AUTO pick by layout."""
    torch = pytest.importorskip("torch")
    code, dt = PREC[prec]
    tdt = torch.float32 if prec == "f32" else torch.float64
    n_channels, n = 70, 5000
    bank, x, ref = _bank_case(n_channels, n, prec, seed=11)
    stride = 5008
    wide = torch.zeros(n_channels * stride + 8, device="cuda", dtype=tdt)
    a = wide[: n_channels * stride].view(n_channels, stride)
    a[:, :n] = torch.from_numpy(x).cuda()
    bank.process_ptr(a.data_ptr(), n, stride, K.PTR_DEVICE, K.IIR_AUTO, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert "tma" in bank.describe(n, stride)
    y_tma = a[:, :n].cpu().numpy()
    st_tma = bank.get_state()
    bank.reset_state()
    wide2 = torch.zeros(n_channels * stride + 8, device="cuda", dtype=tdt)
    b = wide2[1: n_channels * stride + 1].view(n_channels, stride)  # base off by one element
    b[:, :n] = torch.from_numpy(x).cuda()
    bank.process_ptr(b.data_ptr(), n, stride, K.PTR_DEVICE, K.IIR_AUTO, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    y_gen = b[:, :n].cpu().numpy()
    assert np.array_equal(y_tma, y_gen)
    assert np.array_equal(st_tma, bank.get_state())
    assert peak_rel(y_tma, ref) <= IIR_TOL[prec]
    assert float(a[:, n:].abs().max()) == 0.0 and float(wide[n_channels * stride:].abs().max()) == 0.0  # nothing outside the ranges written


# ------------------------------------------------------------------ the time-parallel (scan) path
def _scan_reference(ftype, f0, fs, x, sections=4):
    f = O.Iir(sections)
    f.design(ftype, f0, fs, 1.1)
    return f, f.process(x)


TIME_PARALLEL = {"auto": K.IIR_SCAN, "lookback": K.IIR_SCAN_LOOKBACK, "split": K.IIR_SCAN_SPLIT}


@pytest.mark.parametrize("how", ["split", "lookback"])
@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("case", [(1, 10e3, 100e3), (2, 10e3, 100e3), (3, 2000.0, 39e3), (1, 200.0, 39e3)])
def test_scan_single_long_channel_matches_sequential_reference(case, prec, how):
    """BASELINE config 4 in miniature: one channel, time axis split over the whole GPU -- by the time-split
    kernel (segments as rows of the lane-per-row kernel + natural-response correction) and by the look-back scan."""
    ftype, f0, fs = case
    code, dt = PREC[prec]
    tp = TIME_PARALLEL[how]
    chunk = 128 if prec == "f64" else 256
    n = 32 * chunk * 150 + 333  # whole tiles / segments + a ragged tail (sequential kernel)
    rng = np.random.default_rng(int(f0) + ftype)
    x = f32_noise(rng, n)
    f, ref = _scan_reference(ftype, f0, fs, x)
    g, b, a = S.design(ftype, 4, f0, fs, 1.1)
    bank = S.IirBank(4, 1, code)
    bank.set_coeffs([g], [b], [a])
    assert ("time-split" in bank.describe(n, n, tp)) == (how == "split")
    y = bank.process(x.astype(dt), path=tp)
    tol = IIR_TOL[prec]
    assert peak_rel(y, ref) <= tol
    # the history left behind continues the stream exactly where the scan stopped
    nxt = f32_noise(rng, 1000)
    want = f.process(nxt)
    got = bank.process(nxt.astype(dt), path=K.IIR_SEQUENTIAL)
    assert peak_rel(got, want) <= tol
    # and a second scan call on the same bank (non-zero incoming history) too
    more = f32_noise(rng, 32 * chunk * 3)
    want2 = f.process(more)
    got2 = bank.process(more.astype(dt), path=K.IIR_SCAN if how == "split" else tp)  # (short: auto may pick either)
    assert peak_rel(got2, want2) <= tol


def test_scan_general_carry_path_for_long_memory_filters():
    """f0/fs = 2e-4: the tile-to-tile propagation matrix does not vanish within the look-back window, so
    tiles wait for their predecessor's inclusive state.  Correctness only (this path is slow by design).
    fp64 error is ~1e-9 of peak here because such a filter amplifies any rounding by ~1e6."""
    n = 32 * 128 * 24
    rng = np.random.default_rng(20)
    x = f32_noise(rng, n)
    f, ref = _scan_reference(1, 20.0, 100e3, x)
    g, b, a = S.design(1, 4, 20.0, 100e3)
    bank = S.IirBank(4, 1, K.F64)
    bank.set_coeffs([g], [b], [a])
    assert "look-back" in bank.describe(n, n, K.IIR_SCAN)  # filter memory (~3e5 samples) exceeds any segment of this call
    with pytest.raises(RuntimeError):
        bank.process(x.copy(), path=K.IIR_SCAN_SPLIT)
    y = bank.process(x.copy(), path=K.IIR_SCAN)
    assert peak_rel(y, ref) <= 1e-8


@pytest.mark.parametrize("how", ["auto", "lookback"])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_scan_few_channels_each_split_along_time(prec, how):
    torch = pytest.importorskip("torch")
    code, dt = PREC[prec]
    tp = TIME_PARALLEL[how]
    chunk = 128 if prec == "f64" else 256
    n_channels, n = 6, 32 * chunk * 100 + 100
    stride = 32 * chunk * 101
    bank, x, ref = _bank_case(n_channels, n, prec, seed=77)
    tdt = torch.float32 if prec == "f32" else torch.float64
    wide = torch.zeros(n_channels, stride, device="cuda", dtype=tdt)
    wide[:, :n] = torch.from_numpy(x).cuda()
    if how == "auto":
        assert "time-split" in bank.describe(n, stride, tp)
    bank.process_ptr(wide.data_ptr(), n, stride, K.PTR_DEVICE, tp, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert peak_rel(wide[:, :n].cpu().numpy(), ref) <= IIR_TOL[prec]
    assert float(wide[:, n:].abs().max()) == 0.0
    # the bank history continues the stream: next block through the sequential kernel
    st = bank.get_state()
    assert np.isfinite(st).all()


def test_config4_scale_single_channel_time_split_against_the_sequential_kernel():
    """BASELINE config 4 at 1/16 scale (one fp64 channel of 2^26 samples): the time-split path against the
    sequential kernel on the same stream (the oracle pins the sequential kernel; a 2^26-sample stream through the
    scalar CPU oracle would dominate the suite), plus the oracle itself on windows of the stream re-run from the
    history the GPU left at their start."""
    torch = pytest.importorskip("torch")
    n = 1 << 26
    g, b, a = S.design(1, 4, 10e3, 100e3)
    gen = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(1, n, device="cuda", dtype=torch.float64, generator=gen)
    seq = S.IirBank(4, 1, K.F64)
    seq.set_coeffs([g], [b], [a])
    split = S.IirBank(4, 1, K.F64)
    split.set_coeffs([g], [b], [a])
    assert "time-split" in split.describe(n, n, K.IIR_SCAN)
    y_seq, y_split = x.clone(), x.clone()
    seq.process_ptr(y_seq.data_ptr(), n, n, K.PTR_DEVICE, K.IIR_SEQUENTIAL, torch.cuda.current_stream().cuda_stream)
    split.process_ptr(y_split.data_ptr(), n, n, K.PTR_DEVICE, K.IIR_SCAN, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    peak = float(y_seq.abs().max())
    assert float((y_seq - y_split).abs().max()) / peak <= IIR_TOL["f64"]
    assert np.abs(seq.get_state() - split.get_state()).max() / peak <= IIR_TOL["f64"]
    # oracle on two windows: the head of the stream, and a window deep inside it warmed up over 4096 samples
    f = O.Iir(4)
    f.design(1, 10e3, 100e3, 1.1)
    head = f.process(x[0, :65536].cpu().numpy())
    assert peak_rel(y_split[0, :65536].cpu().numpy(), head) <= IIR_TOL["f64"]
    lo = (n // 2) + 12345
    f2 = O.Iir(4)
    f2.design(1, 10e3, 100e3, 1.1)
    mid = f2.process(x[0, lo - 4096: lo + 65536].cpu().numpy())[4096:]  # 4096 samples >> the filter's memory (~450)
    assert peak_rel(y_split[0, lo: lo + 65536].cpu().numpy(), mid) <= IIR_TOL["f64"]


def test_config3_full_size_sampled_parity_both_paths():
    """BASELINE config 3 at full size -- 16384 channels x 2^20 fp32 samples, planar, in place (64 GiB): even channels
    low-pass, odd channels high-pass, f0 log-spaced 1-20 kHz at fs = 100 kHz.  Parity against the oracle on channels
    spread across the bank: the head of 64 channels, 8 whole channels, and the history left behind (the stream is
    continued on both sides).  First the bit-exact-streaming path, then the time-split path over the same buffer again
    (its input is the first pass's output)."""
    torch = pytest.importorskip("torch")
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    ch, n, fs, head = 16384, 1 << 20, 100e3, 1 << 16
    if free < (ch * n * 4) * 1.05:
        pytest.skip("needs 68 GB of free device memory")
    ftype = np.where(np.arange(ch) % 2 == 0, 1, 2)
    f0 = np.geomspace(1e3, 20e3, ch)
    coef = [S.design(int(t), 4, float(f), fs) for t, f in zip(ftype, f0)]
    bank = S.IirBank(4, ch, K.F32)
    bank.set_coeffs(np.array([c[0] for c in coef]), np.array([c[1] for c in coef]), np.array([c[2] for c in coef]))
    gen = torch.Generator(device="cuda").manual_seed(1234)
    data = torch.empty(ch, n, device="cuda", dtype=torch.float32)
    for lo in range(0, ch, 256):
        data[lo:lo + 256].normal_(generator=gen)
    idx = np.linspace(0, ch - 1, 64).astype(int)
    whole = idx[::8]
    d_idx, d_whole = torch.from_numpy(idx).cuda(), torch.from_numpy(whole).cuda()
    x_head = data[d_idx, :head].cpu().numpy().astype(np.float64)
    x_whole = data[d_whole].cpu().numpy().astype(np.float64)
    stream = torch.cuda.current_stream().cuda_stream
    filters = None
    for path, want_plan in ((K.IIR_AUTO, "sequential/tma"), (K.IIR_SCAN, "time-split")):
        assert want_plan in bank.describe(n, n, path)
        bank.reset_state()
        bank.process_ptr(data.data_ptr(), n, n, K.PTR_DEVICE, path, stream)
        torch.cuda.synchronize()
        got_head = data[d_idx, :head].cpu().numpy()
        assert peak_rel(got_head, O.iir_bank_port(x_head, ftype[idx], f0[idx], fs)) <= IIR_TOL["f32"], path
        filters = []
        got_whole = data[d_whole].cpu().numpy()
        for j, c in enumerate(whole):
            f = O.Iir(4)
            f.design(int(ftype[c]), float(f0[c]), fs, 1.1)
            assert peak_rel(got_whole[j], f.process(x_whole[j])) <= IIR_TOL["f32"], (path, c)
            filters.append(f)  # keeps the history of this pass
        x_head, x_whole = got_head.astype(np.float64), got_whole.astype(np.float64)  # what the next pass filters
    # the history the time-split pass left in the bank continues the stream like the oracle's
    tail = f32_noise(np.random.default_rng(8), (len(whole), 4096))
    block = torch.zeros(ch, 4096, device="cuda", dtype=torch.float32)
    block[d_whole] = torch.from_numpy(tail.astype(np.float32)).cuda()
    bank.process_ptr(block.data_ptr(), 4096, 4096, K.PTR_DEVICE, K.IIR_AUTO, stream)
    torch.cuda.synchronize()
    got_tail = block[d_whole].cpu().numpy()
    for j in range(len(whole)):
        assert peak_rel(got_tail[j], filters[j].process(tail[j])) <= IIR_TOL["f32"], whole[j]


def test_config4_full_size_windows_against_the_oracle():
    """BASELINE config 4 at full size: one fp64 channel of 2^30 samples (8 GiB) through the time-split path; windows
    of the output against the oracle run from 8192 samples before each window (18 x the filter's memory of ~450
    samples, so the oracle's zero start has decayed below 1e-70 of its state by the time the window begins)."""
    torch = pytest.importorskip("torch")
    torch.cuda.empty_cache()
    n, win, warm = 1 << 30, 1 << 16, 8192
    g, b, a = S.design(1, 4, 10e3, 100e3)
    gen = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.empty(1, n, device="cuda", dtype=torch.float64)
    for lo in range(0, n, 1 << 27):
        x[:, lo:lo + (1 << 27)].normal_(generator=gen)
    starts = [0, (n // 3) + 777, n - win]
    x_win = [x[0, max(0, s0 - warm): s0 + win].cpu().numpy() for s0 in starts]
    bank = S.IirBank(4, 1, K.F64)
    bank.set_coeffs([g], [b], [a])
    assert "time-split" in bank.describe(n, n, K.IIR_SCAN)
    bank.process_ptr(x.data_ptr(), n, n, K.PTR_DEVICE, K.IIR_SCAN, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(x[0, :: 4099]).all())
    for s0, xin in zip(starts, x_win):
        f = O.Iir(4)
        f.design(1, 10e3, 100e3, 1.1)
        ref = f.process(xin)[-win:]
        assert peak_rel(x[0, s0: s0 + win].cpu().numpy(), ref) <= IIR_TOL["f64"], s0


def test_banks_and_plans_on_two_devices_in_one_process():
    """One process, two GPUs: kernel attributes (dynamic shared memory) are per device, handles are bound to the device
    they were created on.  Skipped on a single-GPU box."""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(3)
    g, b, a = S.design(1, 4, 10e3, 100e3)
    x = f32_noise(rng, (40, 6000))
    f = [O.Iir(4) for _ in range(40)]
    ref = np.empty_like(x)
    for c in range(40):
        f[c].design(1, 10e3, 100e3, 1.1)
        ref[c] = f[c].process(x[c])
    z = (rng.standard_normal((8, 4096)) + 1j * rng.standard_normal((8, 4096))).astype(np.complex64)
    zref = np.fft.fft(z.astype(np.complex128))
    for dev in (1, 0, 1):
        with torch.cuda.device(dev):
            bank = S.IirBank(4, 40, K.F32, K.NUM_GENERIC, dev)
            bank.set_coeffs(np.full(40, g), np.tile(b, (40, 1, 1)), np.tile(a, (40, 1, 1)))
            d = torch.from_numpy(x.astype(np.float32)).cuda(dev)
            bank.process(d)
            torch.cuda.synchronize(dev)
            assert peak_rel(d.cpu().numpy(), ref) <= IIR_TOL["f32"], dev
            long = torch.from_numpy(np.tile(x[:1].astype(np.float32), (1, 40))).cuda(dev)  # 1 channel x 240000: time-split
            one = S.IirBank(4, 1, K.F32, K.NUM_GENERIC, dev)
            one.set_coeffs([g], [b], [a])
            one.process_ptr(long.data_ptr(), long.shape[1], long.shape[1], K.PTR_DEVICE, K.IIR_SCAN, torch.cuda.current_stream(dev).cuda_stream)
            torch.cuda.synchronize(dev)
            assert bool(torch.isfinite(long).all())
            plan = S.FftPlan(4096, 4, K.F32, K.FORWARD, dev)
            zd = torch.from_numpy(z).cuda(dev)
            plan(zd)
            torch.cuda.synchronize(dev)
            assert rel_l2(zd.cpu().numpy(), zref) <= 1e-5, dev


@pytest.mark.parametrize("seed", range(10))
def test_time_split_random_shapes_against_the_sequential_kernel(seed):
    """Random banks (1-40 channels), lengths and channel pitches through the time-parallel path against the sequential
    kernel on the same device data: exercises the planner's segment choice, the leftover rounds, the ragged tail and the
    3-D tensor map with pitches that are not multiples of anything convenient.  Memory outside the ranges stays untouched."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(1000 + seed)
    prec = "f32" if seed % 2 else "f64"
    code, dt = PREC[prec]
    tdt = torch.float32 if prec == "f32" else torch.float64
    per16 = 4 if prec == "f32" else 2
    ch = int(rng.integers(1, 41))
    n = int(rng.integers(60_000, 700_000))
    pitch = (n + int(rng.integers(0, 5000)) + per16 - 1) // per16 * per16  # 16-byte multiple, otherwise arbitrary
    fs = 100e3
    ftype = rng.integers(1, 3, size=ch)
    f0 = np.exp(rng.uniform(np.log(2e3), np.log(3e4), size=ch))
    coef = [S.design(int(t), 4, float(f), fs) for t, f in zip(ftype, f0)]
    banks = []
    for _ in range(2):
        b = S.IirBank(4, ch, code)
        b.set_coeffs(np.array([c[0] for c in coef]), np.array([c[1] for c in coef]), np.array([c[2] for c in coef]))
        banks.append(b)
    gen = torch.Generator(device="cuda").manual_seed(seed)
    base = torch.zeros(ch * pitch + 64, device="cuda", dtype=tdt)
    view = base[: ch * pitch].view(ch, pitch)
    view[:, :n] = torch.randn(ch, n, device="cuda", generator=gen, dtype=tdt)
    other = base.clone()
    stream = torch.cuda.current_stream().cuda_stream
    plan = banks[0].describe(n, pitch, K.IIR_SCAN)
    banks[0].process_ptr(base.data_ptr(), n, pitch, K.PTR_DEVICE, K.IIR_SCAN, stream)
    banks[1].process_ptr(other.data_ptr(), n, pitch, K.PTR_DEVICE, K.IIR_SEQUENTIAL, stream)
    torch.cuda.synchronize()
    a, b_ = base[: ch * pitch].view(ch, pitch), other[: ch * pitch].view(ch, pitch)
    peak = b_[:, :n].abs().amax(dim=1)
    err = ((a[:, :n] - b_[:, :n]).abs().amax(dim=1) / peak).max()
    assert float(err) <= (1e-10 if prec == "f64" else 2e-5), plan
    assert float(a[:, n:].abs().max() if pitch > n else 0.0) == 0.0 and float(base[ch * pitch:].abs().max()) == 0.0, plan
    st_a, st_b = banks[0].get_state(), banks[1].get_state()
    assert np.abs(st_a - st_b).max() <= (1e-10 if prec == "f64" else 2e-5) * max(1.0, float(peak.max())), plan


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_host_buffers_cut_along_time_match_the_device_resident_call(prec):
    """A host-pointer call stages the bank through device memory in chunks of time across all channels (64 MB, two
    streams, simpledsp_b200/csrc/host_stage.h); the history carries from chunk to chunk as from call to call, so the
    sequential path must give the bits of ONE device-resident call -- and the same history afterwards.  300 channels x
    150000 samples = 180 MB (fp32) / 360 MB (fp64): three / six chunks, ragged last one, host row pitch != chunk pitch."""
    torch = pytest.importorskip("torch")
    code, dt = PREC[prec]
    ch, n, pitch = 300, 150_000, 150_016
    bank, x, ref = _bank_case(ch, n, prec, seed=5)
    d = torch.from_numpy(x).cuda()
    bank.process(d, path=K.IIR_SEQUENTIAL)
    torch.cuda.synchronize()
    want, st = d.cpu().numpy(), (bank.get_state(), bank.get_state_diff())
    assert peak_rel(want, ref) <= IIR_TOL[prec]
    host = np.zeros((ch, pitch), dtype=dt)
    host[:, :n] = x
    bank.reset_state()
    bank.process_ptr(host.ctypes.data, n, pitch, K.PTR_HOST, K.IIR_SEQUENTIAL, None)
    assert np.array_equal(host[:, :n], want)
    assert not host[:, n:].any()
    assert np.array_equal(bank.get_state(), st[0]) and np.array_equal(bank.get_state_diff(), st[1])
    # the time-parallel request through host buffers: every chunk is split again; same stream within the tolerance
    bank.reset_state()
    host[:, :n] = x
    bank.process_ptr(host.ctypes.data, n, pitch, K.PTR_HOST, K.IIR_SCAN, None)
    assert peak_rel(host[:, :n], ref) <= IIR_TOL[prec]


def test_auto_goes_time_parallel_only_for_small_banks_on_long_calls():
    """SDSP_B200_IIR_AUTO (include/sdsp_b200.h): sequential kernels -- bit-identical re-blocking -- unless the bank has
    at most 2 x SMs warps of channels AND the call is at least 65536 samples long; then the time-split path."""
    torch = pytest.importorskip("torch")
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    small, n_long = 64, 1 << 19
    bank, x, ref = _bank_case(small, n_long, "f32", seed=21)
    assert "time-split" in bank.describe(n_long) and "auto" in bank.describe(n_long)
    assert "sequential" in bank.describe(65535)
    assert "sequential" in bank.describe(n_long, path=K.IIR_SEQUENTIAL)
    y = bank.process(torch.from_numpy(x).cuda()).cpu().numpy()
    assert peak_rel(y, ref) <= IIR_TOL["f32"]
    # the same stream in 32-sample blocks (reference test/testIIR.cpp:61-75) stays on the sequential kernels and is
    # bit-identical to one SEQUENTIAL call -- and within the tolerance of what AUTO gave for the long call
    bank.reset_state()
    whole = bank.process(torch.from_numpy(x[:, :4096].copy()).cuda(), path=K.IIR_SEQUENTIAL).cpu().numpy()
    bank.reset_state()
    blocks = x[:, :4096].copy()
    for lo in range(0, 4096, 32):
        part = np.ascontiguousarray(blocks[:, lo:lo + 32])
        bank.process(part)
        blocks[:, lo:lo + 32] = part
    assert np.array_equal(whole, blocks)
    assert peak_rel(y[:, :4096], whole) <= IIR_TOL["f32"]
    big = S.IirBank(4, 32 * (2 * sms + 1), K.F32)  # one warp of channels past the threshold
    assert "sequential" in big.describe(n_long)


@pytest.mark.parametrize("sections,ftype", [(12, 1), (16, 2), (10, 3)])
def test_single_object_with_more_sections_than_a_kernel_holds(sections, ftype):
    """The reference's template takes any even m_t (casc_2o_iir.h:23-26).  The kernels hold up to eight sections, so the
    single-object path chains groups of eight (sdsp_b200_iir_process_once); against the oracle's filter of the same order,
    whole buffer and 32-sample blocks (bit-identical, reference test/testIIR.cpp:61-75), float samples included."""
    rng = np.random.default_rng(sections)
    fs, f0, q = 100e3, 12e3, 1.3
    x = f32_noise(rng, 3000)
    f = O.Iir(sections)
    f.design(ftype, f0, fs, q)
    ref = f.process(x)
    for prec in ("f64", "f32"):
        code, dt = PREC[prec]
        g = S.casc_2o_iir(sections, code)
        _design(g, ftype, f0, fs, q)
        g2 = g.copy()
        y = g.process(x.astype(dt))
        assert peak_rel(y, ref) <= (IIR_TOL["f64"] if prec == "f64" else 1e-7), (sections, prec)
        y2 = x.astype(dt)
        for lo in range(0, x.size, 32):
            g2.process(y2[lo:lo + 32])
        assert np.array_equal(y, y2), (sections, prec)
        assert np.array_equal(g.mem, g2.mem)
