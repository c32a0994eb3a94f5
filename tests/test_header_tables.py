"""The compile-time tables of the drop-in sdsp/fft.h against the reference's own header -- CPU only.

tests/cpp/tables_dump.cpp is compiled against include/ (this repo) and, where /root/reference is present (the build
container; the GPU box has no copy), against the reference's include directory: calc_trigs, calc_trigs_naive, calc_wCoeffs
(both directions), calc_swap_lookup (bases 2 and 4), digit_reverse and the integer helpers must agree BYTE FOR BYTE,
signs of zeros included (reference include/sdsp/fft.h:12-43, 54-119, 148-256).  Everywhere, the dump is also checked
against the CPU oracle's restatement of the same tables."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import ROOT

REFERENCE = os.environ.get("SDSP_REFERENCE", "/root/reference")
SIZES = [2, 4, 8, 64, 128, 256, 1024]


def _build_and_run(tmp_path, name, include_dir):
    exe = tmp_path / name
    subprocess.run(["g++", "-std=gnu++20", "-O1", "-fconstexpr-ops-limit=400000000", "-I", include_dir,
                    os.path.join(ROOT, "tests", "cpp", "tables_dump.cpp"), "-o", str(exe)], check=True)
    return subprocess.run([str(exe)], check=True, capture_output=True).stdout


def _parse(blob):
    out, pos = {}, 0

    def take(dtype, count):
        nonlocal pos
        a = np.frombuffer(blob, dtype=dtype, count=count, offset=pos)
        pos += a.nbytes
        return a

    for n in SIZES:
        rows = n.bit_length() - 1
        pow4 = rows % 2 == 0
        d = {"cos": take(np.float64, rows * n).reshape(rows, n), "sin": take(np.float64, rows * n).reshape(rows, n),
             "naive": take(np.float64, rows * n).reshape(rows, n), "wf": take(np.complex128, rows * n).reshape(rows, n),
             "wr": take(np.complex128, rows * n).reshape(rows, n), "swap2": take(np.uint32, n)}
        if pow4:
            d["swap4"] = take(np.uint32, n)
        d["ints"] = take(np.uint32, 4)
        d["rev"] = take(np.uint32, 2 * len(range(0, n, 7))).reshape(-1, 2)
        out[n] = d
    assert pos == len(blob)
    return out


def test_dropin_tables_equal_the_reference_header_byte_for_byte(tmp_path):
    ours = _build_and_run(tmp_path, "ours", os.path.join(ROOT, "include"))
    if os.path.isdir(os.path.join(REFERENCE, "include", "sdsp")):
        theirs = _build_and_run(tmp_path, "theirs", os.path.join(REFERENCE, "include"))
        assert ours == theirs
    t = _parse(ours)
    for n in SIZES:
        d = t[n]
        rows = n.bit_length() - 1
        assert d["ints"].tolist() == [rows, rows // 2, 1, int(rows % 2 == 0)]
        assert np.array_equal(d["swap2"], O.swap_lookup(n, 2))
        if "swap4" in d:
            assert np.array_equal(d["swap4"], O.swap_lookup(n, 4))
        idx = np.arange(0, n, 7)
        assert np.array_equal(d["rev"][:, 0], O.digit_reverse(n, 2)[idx])
        if rows % 2 == 0:
            assert np.array_equal(d["rev"][:, 1], O.digit_reverse(n, 4)[idx])
        # the oracle's tables come from run-time libm, the header's from GCC's constant folding: equal to an ulp, and
        # exactly equal wherever the value is exact (0, +-1)
        w = O.wcoeffs(n)
        assert np.abs(d["wf"] - w).max() <= 2.3e-16
        assert np.abs(d["wr"] - np.conj(w)).max() <= 2.3e-16
        exact = (np.abs(w.real) == 1) | (w.real == 0)
        assert np.array_equal(d["wf"].real[exact], w.real[exact])
        assert np.abs(d["cos"] - d["naive"]).max() <= n * 2.3e-16  # against plain libm, whose argument 2 pi j / P is rounded (j up to n)
        assert np.array_equal(d["cos"], d["wf"].real) and np.array_equal(d["sin"], -d["wf"].imag)


@pytest.mark.parametrize("n", [2, 4, 64, 256, 1024, 4096])
def test_library_twiddle_table_against_the_oracle(n):
    """sdsp_b200_twiddle_table (the host copy of the generator that fills the device tables; octant-symmetric, long double)
    against the oracle's calc_wCoeffs (reference fft.h:197-214): within one ulp, both directions."""
    import ctypes as C

    from simpledsp_b200 import _capi as K

    rows = n.bit_length() - 1
    for direction, want in ((K.FORWARD, O.wcoeffs(n)), (K.REVERSE, O.wcoeffs(n, inverse=True))):
        out = np.zeros((rows, n), dtype=np.complex128)
        K.check(K.lib().sdsp_b200_twiddle_table(n, direction, out.ctypes.data_as(C.POINTER(C.c_double))))
        assert np.abs(out - want).max() <= 2.3e-16, (n, direction)  # one ulp of a value in [0.5, 1]
        assert np.abs(np.abs(out) - 1).max() <= 2.3e-16
