"""C-ABI library checks that need no GPU: it loads, exports every declared symbol, validates arguments,
designs filters exactly like the oracle, and the host emulation of the kernels' per-thread code agrees
with the oracle (index arithmetic / rounding of the CUDA path, checked where no device exists)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import simpledsp_b200 as S
from oracle import oracle as O
from simpledsp_b200 import _capi as K
from tests.util import FFT_TOL, IIR_GOLDEN_ABS, IIR_TOL, ROOT, golden_impulses, peak_rel, ref_vectors, rel_l2, IIR_FIXTURE_F32

dp = C.POINTER(C.c_double)


def _has_gpu():
    n = C.c_int()
    return K.lib().sdsp_b200_device_count(C.byref(n)) == K.OK and n.value > 0


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sdsp_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(sdsp_b200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    lib = C.CDLL(K.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/sdsp_b200.h but not exported"
    assert declared == set(K.SIGNATURES), declared ^ set(K.SIGNATURES)
    assert K.lib().sdsp_b200_version() == 200


def test_no_device_is_a_loud_error_not_a_fallback():
    if _has_gpu():
        pytest.skip("a GPU is present")
    x = np.zeros(64, dtype=np.complex64)
    with pytest.raises(S.SdspError) as e:
        S.fft_radix2(x)
    assert e.value.status == K.ERR_NO_DEVICE
    with pytest.raises(S.SdspError):
        S.IirBank(4, 8)
    with pytest.raises(S.SdspError):
        S.digit_reverse_table(64, 2)
    with pytest.raises(S.SdspError):
        S.FftPlan(64, 2, K.F32, K.FORWARD).half_spectrum(np.zeros((2, 64), dtype=np.float32))


def test_argument_validation():
    L = K.lib()
    h = C.c_void_p()
    assert L.sdsp_b200_fft_plan_create(C.byref(h), 100, 2, K.F32, K.FORWARD, 0) == K.ERR_INVALID_ARG
    assert b"power of 2" in L.sdsp_b200_last_error()       # fft.h:261
    assert L.sdsp_b200_fft_plan_create(C.byref(h), 32, 4, K.F32, K.FORWARD, 0) == K.ERR_INVALID_ARG
    assert b"power of 4" in L.sdsp_b200_last_error()       # fft.h:304
    assert L.sdsp_b200_fft_plan_create(C.byref(h), 64, 3, K.F32, K.FORWARD, 0) == K.ERR_INVALID_ARG
    assert L.sdsp_b200_iir_bank_create(C.byref(h), 9, 4, K.F32, 0, 0) == K.ERR_UNSUPPORTED
    assert L.sdsp_b200_iir_bank_create(C.byref(h), 4, 0, K.F32, 0, 0) == K.ERR_INVALID_ARG
    assert L.sdsp_b200_digit_reverse_table(48, 2, 0, None, 0) == K.ERR_INVALID_ARG
    with pytest.raises(ValueError):
        S.casc_2o_iir(3)                                    # casc_2o_iir.h:25 "M must be even!"


@pytest.mark.parametrize("sections", [2, 4, 6, 8])
def test_designers_match_oracle_bit_for_bit(sections):
    for ftype in (1, 2, 3):
        for f0, fs, q, gain in ((200.0, 39e3, 1.4, 1.0), (2000.0, 39e3, 0.8, 2.0), (15000.0, 39e3, 2.0, 0.5), (10e3, 100e3, 1.1, 1.0)):
            g, b, a = S.design(ftype, sections, f0, fs, q, gain)
            f = O.Iir(sections)
            f.design(ftype, f0, fs, q, gain)
            og, ob, oa = f.coefficients()
            assert g == og and np.array_equal(b, ob) and np.array_equal(a, oa)


def test_preload_state_matches_oracle():
    for ftype in (1, 2, 3):
        f = S.casc_2o_iir(4)
        {1: f.set_lp_coeff, 2: f.set_hp_coeff}.get(ftype, lambda a, b: f.set_bp_coeff(a, b, 1.1))(10e3, 100e3)
        f.preload_filter(10.0)
        o = O.Iir(4)
        o.design(ftype, 10e3, 100e3, 1.1)
        o.preload_filter(10.0)
        want = np.array([[o._s.mem[r][0], o._s.mem[r][1]] for r in range(5)])
        assert np.array_equal(f.mem, want)


@pytest.mark.parametrize("lg", range(1, 15))
def test_emulated_fft_kernel_code_matches_oracle(lg):
    n = 1 << lg
    rng = np.random.default_rng(lg)
    x = rng.standard_normal((2, n)).astype(np.float32) + 1j * rng.standard_normal((2, n)).astype(np.float32)
    x = x.astype(np.complex128)
    for inv in (False, True):
        ref = O.fft(x, 2, inv) if n >= 4 else (np.fft.ifft(x) if inv else np.fft.fft(x))
        for name, prec, dt in (("f64", K.F64, np.complex128), ("f32", K.F32, np.complex64)):
            if prec == K.F64 and lg > 13:
                continue
            a = np.ascontiguousarray(x.astype(dt))
            K.check(K.lib().sdsp_b200_debug_emulate_fft(n, prec, int(inv), a.ctypes.data, 2))
            assert rel_l2(a, ref) < FFT_TOL[name] * (1e-3 if name == "f64" else 0.1), (n, name, inv)


@pytest.mark.parametrize("lg", range(2, 16))
def test_emulated_half_spectrum_kernel_code_matches_numpy(lg):
    """sdsp_b200_debug_emulate_r2c runs the per-thread code of fft_r2c_kernel / fft_c2r_kernel on the host (the M = n/2 point frame code,
    the separation step r2c_bin / c2r_bin, the factors W_n^t x 64th root as the kernels form them): half spectra of real frames against
    numpy's rfft -- which agrees with the reference's transform of (x, 0) to 3e-16 (tests/test_oracle.py) -- and the way back against
    irfft, both precisions, at the FFT tolerances."""
    n = 1 << lg
    frames = 3
    rng = np.random.default_rng(lg)
    for prec, rdt, cdt in ((K.F64, np.float64, np.complex128), (K.F32, np.float32, np.complex64)):
        x = rng.standard_normal((frames, n)).astype(np.float32).astype(rdt)
        half = np.zeros((frames, n // 2 + 1), dtype=cdt)
        K.check(K.lib().sdsp_b200_debug_emulate_r2c(n, prec, 0, x.ctypes.data, half.ctypes.data, frames))
        ref = np.fft.rfft(x.astype(np.float64), axis=1)
        assert rel_l2(half, ref) <= FFT_TOL["f64" if prec == K.F64 else "f32"], (n, prec)
        assert np.all(half[:, n // 2].imag == 0)
        back = np.zeros((frames, n), dtype=rdt)
        K.check(K.lib().sdsp_b200_debug_emulate_r2c(n, prec, 1, half.ctypes.data, back.ctypes.data, frames))
        assert np.linalg.norm(back - x) / np.linalg.norm(x) <= 2 * FFT_TOL["f64" if prec == K.F64 else "f32"], (n, prec)
        # an arbitrary half spectrum (bins 0 and n/2 real) back to a real frame
        h2 = (rng.standard_normal((frames, n // 2 + 1)) + 1j * rng.standard_normal((frames, n // 2 + 1))).astype(cdt)
        h2[:, 0] = h2[:, 0].real
        h2[:, -1] = h2[:, -1].real
        K.check(K.lib().sdsp_b200_debug_emulate_r2c(n, prec, 1, h2.ctypes.data, back.ctypes.data, frames))
        ref2 = np.fft.irfft(h2.astype(np.complex128), n=n, axis=1)
        assert np.linalg.norm(back - ref2) / np.linalg.norm(ref2) <= FFT_TOL["f64" if prec == K.F64 else "f32"], (n, prec)
    assert K.lib().sdsp_b200_debug_emulate_r2c(48, K.F32, 0, x.ctypes.data, half.ctypes.data, 1) != 0


def test_emulated_iir_kernel_code_matches_golden():
    for name, ftype, fs, f0, q, n, h in golden_impulses():
        g, b, a = S.design(ftype, 4, f0, fs, q)
        for num in (0, ftype):
            for pname, prec, dt in (("f64", K.F64, np.float64), ("f32", K.F32, np.float32)):
                x = np.zeros(n, dtype=dt)
                x[0] = 1
                mem = np.zeros((5, 2))
                K.check(K.lib().sdsp_b200_debug_emulate_iir(4, num, prec, g, b.ctypes.data_as(dp), a.ctypes.data_as(dp),
                                                            mem.ctypes.data_as(dp), x.ctypes.data, n))
                if pname == "f64":
                    assert np.abs(x - h).max() < IIR_GOLDEN_ABS
                assert peak_rel(x, h) < (IIR_TOL[pname] if pname == "f64" else IIR_FIXTURE_F32), (name, num, pname)


@pytest.mark.parametrize("sections", [2, 4, 6, 8])
def test_skewed_tiles_are_bit_identical_to_the_plain_loop(sections):
    """The TMA kernel filters whole tiles with the sections software-skewed and ragged tails sample by
    sample (iir_core.cuh).  Emulated on the host: a stream cut into 7-sample calls (plain loop only) must
    reproduce the whole-buffer run (skewed tiles) bit for bit -- reference test/testIIR.cpp:61-75."""
    L = K.lib()
    rng = np.random.default_rng(sections)
    for num in (0, 1, 2, 3):
        g, b, a = S.design(num if num else 1, sections, 3000.0, 100e3, 1.1)
        for prec, dt in ((K.F64, np.float64), (K.F32, np.float32)):
            x = rng.standard_normal(1000).astype(dt)
            # the state of a stream = the reference's m_mem + (fp32, difference form) one running difference per section
            whole, mem, dif = x.copy(), np.zeros((sections + 1, 2)), np.zeros(sections)
            K.check(L.sdsp_b200_debug_emulate_iir_diff(sections, num, prec, g, b.ctypes.data_as(dp), a.ctypes.data_as(dp),
                                                       mem.ctypes.data_as(dp), dif.ctypes.data_as(dp), whole.ctypes.data, whole.size))
            parts, mem2, dif2 = x.copy(), np.zeros((sections + 1, 2)), np.zeros(sections)
            for i in range(0, 1000, 7):
                blk = parts[i:i + 7]
                K.check(L.sdsp_b200_debug_emulate_iir_diff(sections, num, prec, g, b.ctypes.data_as(dp), a.ctypes.data_as(dp),
                                                           mem2.ctypes.data_as(dp), dif2.ctypes.data_as(dp), blk.ctypes.data, blk.size))
            assert np.array_equal(whole, parts) and np.array_equal(mem, mem2) and np.array_equal(dif, dif2)
            if prec == K.F64:
                assert not dif.any()  # fp64 runs the direct form: nothing beyond m_mem


@pytest.mark.parametrize("case", [(1, 10e3, 100e3), (2, 10e3, 100e3), (3, 2000.0, 39e3), (1, 200.0, 39e3)])
def test_emulated_scan_algorithm_matches_oracle(case):
    """The chunked state-space scan (iir_scan_core.cuh) played on the host: zero-state chunks, Kogge-Stone
    carry over 32 lanes, look-back over tiles, natural-response correction -- against the sequential
    reference on the same input, both carry paths, ragged length (the tail takes the sequential loop)."""
    ftype, f0, fs = case
    L = K.lib()
    rng = np.random.default_rng(int(f0))
    g, b, a = S.design(ftype, 4, f0, fs, 1.1)
    for chunk in (16, 128):
        n = 32 * chunk * 3 + 77
        x = rng.standard_normal(n).astype(np.float32).astype(np.float64)
        f = O.Iir(4)
        f.design(ftype, f0, fs, 1.1)
        ref = f.process(x)
        for force_general in (0, 1):
            for pname, prec, dt in (("f64", K.F64, np.float64), ("f32", K.F32, np.float32)):
                y = x.astype(dt)
                mem = np.zeros((5, 2))
                K.check(L.sdsp_b200_debug_emulate_iir_scan(4, 0, prec, g, b.ctypes.data_as(dp), a.ctypes.data_as(dp),
                                                           mem.ctypes.data_as(dp), y.ctypes.data, n, chunk, force_general))
                assert peak_rel(y, ref) <= IIR_TOL[pname], (case, chunk, force_general, pname)
                # the history handed back continues the stream: next block through the sequential emulator
                nxt = rng.standard_normal(50).astype(dt)
                want = f.copy().process(nxt.astype(np.float64))
                K.check(L.sdsp_b200_debug_emulate_iir(4, 0, prec, g, b.ctypes.data_as(dp), a.ctypes.data_as(dp),
                                                      mem.ctypes.data_as(dp), nxt.ctypes.data, nxt.size))
                assert peak_rel(nxt, want) <= IIR_TOL[pname]


@pytest.mark.parametrize("case", [(1, 10e3, 100e3), (2, 10e3, 100e3), (3, 2000.0, 39e3), (1, 200.0, 39e3), (1, 1e3, 100e3)])
def test_time_split_algorithm_on_the_host(case):
    """The time-split path (iir_segment.cu) played on the host with the kernels' own per-sample code
    (sdsp_b200_debug_emulate_iir): segments run from (true input history, zero section history), each segment's
    section history handed to the next, natural response added over the first K samples, K from
    sdsp_b200_debug_iir_decay_length.  Checks (1) K against a direct simulation of the natural response, (2) the
    recombined stream against the sequential run, (3) that what is dropped beyond K is below the threshold."""
    ftype, f0, fs = case
    L = K.lib()
    g, b, a = S.design(ftype, 4, f0, fs, 1.1)
    rng = np.random.default_rng(int(f0) + 17)
    for pname, prec, dt, thr in (("f64", K.F64, np.float64, 2.0 ** -62), ("f32", K.F32, np.float32, 2.0 ** -32)):
        k = C.c_ulonglong()
        K.check(L.sdsp_b200_debug_iir_decay_length(4, 0, prec, b.ctypes.data_as(dp), a.ctypes.data_as(dp), C.byref(k)))
        k = int(k.value)
        assert 0 < k < 200000
        # (1) natural response from unit section histories, in double: negligible from sample k on, not yet much earlier
        worst_at_k, worst_before = 0.0, 0.0
        for r in range(1, 5):
            for slot in range(2):
                mem = np.zeros((5, 2))
                mem[r, slot] = 1.0
                z = np.zeros(k + 8)
                K.check(L.sdsp_b200_debug_emulate_iir(4, 0, K.F64, g, b.ctypes.data_as(dp), a.ctypes.data_as(dp),
                                                      mem.ctypes.data_as(dp), z.ctypes.data, z.size))
                worst_at_k = max(worst_at_k, np.abs(z[k:]).max(), np.abs(mem[1:]).max())
                worst_before = max(worst_before, np.abs(z[k // 2: k // 2 + 8]).max())
        assert worst_at_k <= thr * 1.0001
        assert worst_before > thr  # K is not wildly pessimistic: halfway there the response is still above the threshold
        # (2) segments + correction against the sequential run
        seg = max(4 * k, 1024)
        n = 5 * seg
        x = rng.standard_normal(n).astype(np.float32).astype(dt)
        whole = x.copy()
        mem_w = np.zeros((5, 2))
        K.check(L.sdsp_b200_debug_emulate_iir(4, 0, prec, g, b.ctypes.data_as(dp), a.ctypes.data_as(dp), mem_w.ctypes.data_as(dp),
                                              whole.ctypes.data, n))
        y = x.copy()
        carry, carry_d = np.zeros((4, 2)), np.zeros(4)
        gain_t = dt(g)
        for s_ in range(5):
            part = np.ascontiguousarray(y[s_ * seg:(s_ + 1) * seg])
            mem, dif = np.zeros((5, 2)), np.zeros(4)
            if s_:
                mem[0, 0] = float(dt(x[s_ * seg - 1]) * gain_t)  # the product iir_step() forms for row 0
                mem[0, 1] = float(dt(x[s_ * seg - 2]) * gain_t)
            K.check(L.sdsp_b200_debug_emulate_iir_diff(4, 0, prec, g, b.ctypes.data_as(dp), a.ctypes.data_as(dp), mem.ctypes.data_as(dp),
                                                       dif.ctypes.data_as(dp), part.ctypes.data, seg))
            if s_:
                corr = np.zeros(k, dtype=dt)
                cm, cd = np.zeros((5, 2)), carry_d.copy()
                cm[1:] = carry
                K.check(L.sdsp_b200_debug_emulate_iir_diff(4, 0, prec, g, b.ctypes.data_as(dp), a.ctypes.data_as(dp), cm.ctypes.data_as(dp),
                                                           cd.ctypes.data_as(dp), corr.ctypes.data, k))
                part[:k] += corr
            # what this segment ended with (zero-state run): section history and, fp32, the running differences
            carry, carry_d = mem[1:].copy(), dif.copy()
            y[s_ * seg:(s_ + 1) * seg] = part
        peak = np.abs(whole).max()
        # fp64: the two computations agree to rounding.  fp32: they round differently (y0 + correction vs one recurrence); with the
        # difference form both sit within a few 1e-6 of the truth whatever the cutoff, a tenth of the IIR tolerance bounds the gap
        tol = 1e-12 if pname == "f64" else 0.1 * IIR_TOL["f32"]
        assert np.abs(y - whole).max() / peak <= tol, (case, pname)


@pytest.mark.parametrize("half", [0, 1])
def test_real_input_queue_order_makes_every_wait_point_at_a_smaller_ticket(half):
    """The real-input 65536-point kernel (fft_real64k_kernel: 8 packed column tiles + 9 row tiles per frame) hands its items out in one
    static order too.  The same argument on its own decode function: every tile exactly once, a frame's row tiles after all of its
    column tiles, a ring slot's new column tiles after all row tiles of its previous tenant -- for batches below, at and past the
    lag and the ring."""
    L = K.lib()
    geom, item = (C.c_int * 4)(), (C.c_int * 3)()
    K.check(L.sdsp_b200_debug_fft_real_queue_item(half, 0, geom, item))
    ct, rt, lag, ring = list(geom)
    assert (ct, rt) == (8, 9) and ring == 2 * lag and lag >= 1
    assert ring * 129 * 256 * 8 <= 52 << 20  # the ring of half-frames (rows 0 .. 128) stays near the L2 budget the design states
    for frames in sorted({1, 2, lag - 1, lag, lag + 1, ring - 1, ring, ring + 1, 2 * ring + 3}):
        total = lag * ct + frames * (ct + rt)
        col_ticket, row_ticket = {}, {}
        for q in range(total):
            K.check(L.sdsp_b200_debug_fft_real_queue_item(half, q, geom, item))
            is_col, tile, f = list(item)
            if f >= frames:
                assert is_col  # empty slots are column slots past the last frame
                continue
            (col_ticket if is_col else row_ticket).setdefault(f, []).append((q, tile))
        for f in range(frames):
            assert sorted(t for _, t in col_ticket[f]) == list(range(ct))
            assert sorted(t for _, t in row_ticket[f]) == list(range(rt))
            assert max(q for q, _ in col_ticket[f]) < min(q for q, _ in row_ticket[f])
            if f >= ring:
                assert max(q for q, _ in row_ticket[f - ring]) < min(q for q, _ in col_ticket[f])


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("n1", [32, 64, 128, 256, 512, 1024])
def test_fft_work_queue_order_makes_every_wait_point_at_a_smaller_ticket(n1, prec):
    """Frames larger than one CTA run as a queue of column tiles and row tiles handed out in ticket order
    (fft_fused_kernel / fft_fused_tma_kernel).  Its freedom from deadlock rests on the order alone: a row tile waits for the
    column tiles of its frame, a column tile for the row tiles of the frame that held its scratch-ring slot before -- both
    must have been handed out EARLIER, whatever the batch size.  Checked here on the kernels' own decode function
    (sdsp_b200_debug_fft_queue_item), for batches below, at and past the lag and the ring size."""
    L = K.lib()
    geom, item = (C.c_int * 4)(), (C.c_int * 3)()
    p = K.F32 if prec == "f32" else K.F64
    K.check(L.sdsp_b200_debug_fft_queue_item(n1 * 256, p, 0, geom, item))
    tiles, cols, lag, ring = list(geom)
    assert tiles == n1 // 16 and cols * tiles == 256 and ring >= lag + 1 and lag >= 1
    # the scratch ring stays within the L2-resident budget the design states (32 MB; 48 MB where consumed lines are discarded)
    assert ring * n1 * 256 * (8 if prec == "f32" else 16) <= 48 << 20
    for frames in sorted({1, 2, lag - 1, lag, lag + 1, ring - 1, ring, ring + 1, 2 * ring + 3}):
        if frames < 1:
            continue
        total = (lag + 2 * frames) * tiles
        col_ticket, row_ticket = {}, {}
        for q in range(total):
            K.check(L.sdsp_b200_debug_fft_queue_item(n1 * 256, p, q, geom, item))
            is_col, tile, f = list(item)
            if f >= frames:
                assert is_col  # empty slots are column slots past the last frame
                continue
            (col_ticket if is_col else row_ticket).setdefault(f, []).append((q, tile))
        for f in range(frames):
            assert sorted(t for _, t in col_ticket[f]) == list(range(tiles))  # every tile exactly once
            assert sorted(t for _, t in row_ticket[f]) == list(range(tiles))
            first_row = min(q for q, _ in row_ticket[f])
            assert max(q for q, _ in col_ticket[f]) < first_row  # rows after their frame's columns
            if f >= ring:  # the slot's previous tenant has been handed out in full before the new columns
                assert max(q for q, _ in row_ticket[f - ring]) < min(q for q, _ in col_ticket[f])
            # the lead the design relies on: a frame's rows follow its columns by about 2 x lag x tiles tickets
            if frames > lag:
                assert first_row - min(q for q, _ in col_ticket[f]) >= (2 * lag - 1) * tiles or f < lag
    # sizes the queue kernels do not take are refused
    assert L.sdsp_b200_debug_fft_queue_item(4096, p, 0, geom, item) != 0


def test_generic_kernel_without_the_b2_multiply_gives_the_same_bits():
    """A generic bank whose every b2 is exactly 1 (all Butterworth low-/high-pass designs) runs kernels that add in2 instead of
    multiplying it by b2 (iir_core.cuh, NUM_GENERIC_B2ONE = 4): fma(1, in2, t) is t + in2, so nothing may change."""
    L = K.lib()
    rng = np.random.default_rng(4)
    for ftype in (1, 2):
        for sections in (2, 4, 8):
            g, b, a = S.design(ftype, sections, 700.0 * ftype, 39e3)
            assert (b[:, 2] == 1.0).all()
            for prec, dt in ((K.F64, np.float64), (K.F32, np.float32)):
                x = rng.standard_normal(777).astype(dt)
                outs = []
                for kind in (0, 4):
                    y, mem, dif = x.copy(), np.zeros((sections + 1, 2)), np.zeros(sections)
                    K.check(L.sdsp_b200_debug_emulate_iir_diff(sections, kind, prec, g, b.ctypes.data_as(dp), a.ctypes.data_as(dp),
                                                               mem.ctypes.data_as(dp), dif.ctypes.data_as(dp), y.ctypes.data, y.size))
                    outs.append((y, mem, dif))
                assert all(np.array_equal(p, q) for p, q in zip(*outs))
