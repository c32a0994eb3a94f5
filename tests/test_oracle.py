"""Pins the CPU oracle (oracle/sdsp_oracle.c) -- CPU only.

Checked against: the reference's golden impulse responses, the analytic known answers of the
reference's FFT tests, vectors produced by the unmodified reference headers (tests/golden/ref_vectors.npz)
and, when it has been built, the compiled reference itself (oracle/_ref/libsdsp_ref.so)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import IIR_GOLDEN_ABS, golden_impulses, ref_vectors, rel_l2

EPS = np.finfo(np.float64).eps
KIND_OF = {1: "lp", 2: "hp", 3: "bp"}


# ------------------------------------------------------------------ integer path: bit exact
@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_swap_lookup_matches_reference_vectors(n):
    z = ref_vectors()
    for base in (2, 4):
        key = f"swap_b{base}_{n}"
        if key not in z:
            continue
        assert np.array_equal(O.swap_lookup(n, base), z[key])


@pytest.mark.parametrize("n", [16384, 65536])
def test_swap_lookup_checksums_large(n):
    z = ref_vectors()
    for base in (2, 4):
        t = O.swap_lookup(n, base).astype(np.uint64)
        got = np.array([t.sum(), (t * (np.arange(n, dtype=np.uint64) + 1)).sum()], dtype=np.uint64)
        assert np.array_equal(got, z[f"swapsum_b{base}_{n}"])


def test_swap_table_n16_base2_known():
    # SURVEY 8(a) F6, probe of calc_swap_lookup<16,2>
    assert O.swap_lookup(16, 2).tolist() == [0, 8, 4, 12, 4, 10, 6, 14, 8, 9, 10, 13, 12, 13, 14, 15]


@pytest.mark.parametrize("n,base", [(64, 2), (64, 4), (1024, 2), (1024, 4), (4096, 4), (2048, 2)])
def test_digit_reverse_is_an_involution(n, base):
    r = O.digit_reverse(n, base)
    assert np.array_equal(r[r], np.arange(n))
    assert sorted(r.tolist()) == list(range(n))


# ------------------------------------------------------------------ FFT
@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_fft_port_matches_reference_vectors(n):
    z = ref_vectors()
    x = z[f"fft_in_{n}"]
    for radix in (2, 4):
        for d, inv in (("fwd", False), ("inv", True)):
            key = f"fft_r{radix}_{d}_{n}"
            if key not in z:
                continue
            got = O.fft(x, radix, inv)
            if n <= 256:  # libm and GCC's constant folding agree on every twiddle: bit exact
                assert np.array_equal(got, z[key]), key
            else:         # a few twiddles differ by one ulp (oracle/README.md)
                assert rel_l2(got, z[key]) < 4 * EPS, key


@pytest.mark.parametrize("radix", [2, 4])
def test_fft_known_answer_tone(radix):
    # reference test/testFFT.cpp:17-68 / 127-178
    N, n = 64, 7
    i = np.arange(N)
    s = np.cos(n * 2 * np.pi * i / N).astype(np.complex128)
    S = np.zeros(N, dtype=np.complex128)
    S[n] = S[N - n] = N / 2
    tol = 4 * N * EPS
    assert np.abs(O.fft(s, radix) - S).max() < tol
    assert np.abs(O.fft(S, radix, inverse=True) - s).max() < tol
    s2 = np.cos(n * 2 * np.pi * i / N + np.pi / 2).astype(np.complex128)
    S2 = np.zeros(N, dtype=np.complex128)
    S2[n], S2[N - n] = 1j * N / 2, -1j * N / 2
    assert np.abs(O.fft(s2, radix) - S2).max() < tol


@pytest.mark.parametrize("radix", [2, 4])
def test_fft_linearity(radix):
    # reference test/testFFT.cpp:70-125 / 180-235
    N, fs, a1, a2 = 256, 8000.0, 1.5, 2.5
    i = np.arange(N)
    x1 = np.sin(2 * np.pi * 1000.0 / fs * i).astype(np.complex128)
    x2 = np.sin(2 * np.pi * 500.0 / fs * i).astype(np.complex128)
    lhs = O.fft(a1 * x1 + a2 * x2, radix)
    rhs = a1 * O.fft(x1, radix) + a2 * O.fft(x2, radix)
    assert np.abs(lhs - rhs).max() < 4 * N * EPS


def test_fft_against_numpy_and_compiled_reference():
    rng = np.random.default_rng(7)
    for n in (64, 1024, 4096):
        x = rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))
        for radix in (2, 4):
            assert rel_l2(O.fft(x, radix), np.fft.fft(x)) < 1e-15
            if O.have_ref():
                assert rel_l2(O.fft(x, radix), O.fft(x, radix, impl="reference")) < 4 * EPS


@pytest.mark.parametrize("n", [16384, 65536])
def test_fft_port_pinned_above_4096(n):
    """The sizes BASELINE configs 2 and 5 are parity-checked at beyond the reference-vector fixtures: the port against
    the UNMODIFIED reference headers compiled for N = 16384 / 65536 (oracle/_ref/libsdsp_ref_big.so, `make refbig`:
    minutes of constexpr table folding, so it is built once and travels) and against numpy's pocketfft."""
    import os

    rng = np.random.default_rng(n)
    x = rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))
    have_big = os.path.exists(O.REF_BIG_SO)
    for radix in (2, 4):
        for inv in (False, True):
            got = O.fft(x, radix, inv)
            want = np.fft.ifft(x) if inv else np.fft.fft(x)
            assert rel_l2(got, want) < 4 * EPS, (n, radix, inv)
            if have_big:
                assert rel_l2(got, O.fft(x, radix, inv, impl="reference")) < 4 * EPS, (n, radix, inv)
    if not have_big:
        pytest.skip("oracle/_ref/libsdsp_ref_big.so not built: numpy only")
    for base in (2, 4):
        assert np.array_equal(O.swap_lookup(n, base), _ref_big_swap(n, base))


def _ref_big_swap(n, base):
    import ctypes as C

    out = np.zeros(n, dtype=np.uint32)
    assert O.ref_lib(big=True).sdsp_ref_swap_lookup(n, base, out.ctypes.data_as(C.POINTER(C.c_uint32))) == 0
    return out


def test_fft_rejects_bad_sizes():
    with pytest.raises(ValueError):
        O.fft(np.zeros(12, dtype=np.complex128), 2)
    with pytest.raises(ValueError):
        O.fft(np.zeros(32, dtype=np.complex128), 4)  # fft.h:304: radix 4 needs a power of 4


# ------------------------------------------------------------------ IIR
@pytest.mark.parametrize("impl", ["port", "reference"])
def test_iir_golden_impulse_responses(impl):
    # reference test/testIIR.cpp:32-77 and 223-430
    if impl == "reference" and not O.have_ref():
        pytest.skip("oracle/_ref not built")
    count = 0
    for name, ftype, fs, f0, q, n, h in golden_impulses():
        for kind in ("generic", KIND_OF[ftype]):
            f = O.Iir(4, kind, impl)
            f.design(ftype, f0, fs, q)
            f2 = f.copy()
            x = np.zeros(n)
            x[0] = 1.0
            y = f.process(x)
            assert np.abs(y - h).max() < IIR_GOLDEN_ABS, (name, kind)
            parts = [f2.process(x[i:i + 32]) for i in range(0, n, 32)]  # 32-sample blocks + 8-sample tail
            assert np.array_equal(np.concatenate(parts), y), (name, kind)
            count += 1
    assert count == 18


def test_iir_port_matches_reference_vectors():
    z = ref_vectors()
    x = z["iir_in"][0]
    for key in z["iir_cases"]:
        key = str(key)
        _, m, kind, t, f = key.split("_")
        sections, ftype, f0 = int(m[1:]), int(t[1:]), float(f[1:])
        flt = O.Iir(sections, kind, "port")
        flt.design(ftype, f0, 100e3, 1.1, 1.0 if f0 > 1e3 else 0.75)
        y = np.concatenate([flt.process(x[:300]), flt.process(x[300:])])
        assert np.array_equal(y, z[key]), key


def test_iir_gain_and_preload():
    # reference test/testIIR.cpp:79-218
    fs, f0, q = 100e3, 10e3, 1.1
    x = np.zeros(1024)
    x[0] = 1.0
    for ftype in (1, 2, 3):
        f1, f2 = O.Iir(4), O.Iir(4)
        f1.design(ftype, f0, fs, q, 1.0)
        f2.design(ftype, f0, fs, q, 2.0)
        assert np.abs(2.0 * f1.process(x) - f2.process(x)).max() < 1e-12
        f = O.Iir(4)
        f.design(ftype, f0, fs, q)
        f.preload_filter(10.0)
        y = f.process(np.full(1024, 10.0))
        assert np.abs(y - (10.0 if ftype == 1 else 0.0)).max() < 1e-12
        assert np.array_equal(y, ref_vectors()[f"iir_preload_t{ftype}"])


def test_scipy_fixture_generator_reproduces_the_reference_fixtures():
    """tools/make_fixtures.py restates the reference's Octave recipe (test_data/WriteImpulse.m, findIIRCutoffFreq.m) with
    scipy; it must regenerate the nine golden impulse responses, and its CSV writer / parser must round-trip the
    reference's one-line wire format (test/testIIR.cpp:7-28)."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from tools import make_fixtures as M

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "impulse_response.npz"))
    for name in z["names"]:
        hd, h = z[str(name) + "_header"], z[str(name) + "_h"]
        kind = {1: "lp", 2: "hp", 3: "bp"}[int(hd[0])]
        got = M.impulse_response(kind, hd[1], hd[2], hd[3], 8, int(hd[4]))
        assert np.abs(got - h).max() <= 1e-12 * np.abs(h).max(), name
        t, fs, f0, q, back = M.parse_csv_line(M.csv_line(kind, hd[1], hd[2], hd[3], got))
        assert (t, fs, f0, q) == (int(hd[0]), hd[1], hd[2], hd[3]) and np.allclose(back, got, rtol=1e-14, atol=0)


@pytest.mark.parametrize("sections", [2, 4, 6, 8])
def test_designers_and_recurrence_against_scipy_beyond_the_fixtures(sections):
    """An independent second oracle: scipy's butter -> sos -> sosfilt (the recipe of tools/make_fixtures.py)
    against the restated designers + recurrence for orders, sample rates, cut-offs and Qs the nine golden files do not
    cover.  Both build the same Butterworth filter, so the impulse responses agree to rounding."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from tools import make_fixtures as M

    x = np.zeros(600)
    x[0] = 1.0
    for kind, t in (("lp", 1), ("hp", 2), ("bp", 3)):
        for f0, fs, q in ((500.0, 48e3, 0.9), (3000.0, 44.1e3, 1.7), (11e3, 96e3, 2.5), (150.0, 8e3, 1.2)):
            f = O.Iir(sections)
            f.design(t, f0, fs, q)
            h = f.process(x)
            want = M.impulse_response(kind, fs, f0, q, 2 * sections, 600)
            assert np.abs(h - want).max() <= 1e-11 * np.abs(want).max(), (sections, kind, f0, fs, q)
