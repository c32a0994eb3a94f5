#!/usr/bin/env python
"""Mint impulse-response fixtures in the reference's wire format without Octave.

The reference's nine golden files (test_data/impulse_response/*.csv, one line each:
``type,fs,f0,Q,n,h[0..n-1]``, type = sdsp::filter_type, 16 significant digits) were produced by Octave
``butter`` -> ``zp2sos`` -> ``sosfilt`` (test_data/WriteImpulse.m:17-33) with the band edges of the band-pass
found by a zero-crossing search (test_data/findIIRCutoffFreq.m:19).  This restates that recipe with scipy:
``butter(..., output='sos')`` + ``sosfilt``, band edges from a bracketing root finder on the same function.
scipy reproduces all nine committed fixtures to <= 1e-12 of peak (tests/test_oracle.py checks it), so new
fixtures -- other orders, sample rates, fp32 variants -- can be minted where no Octave exists.

    python tools/make_fixtures.py --type lp --fs 39000 --f0 200 --q 1.4 --order 8 --n 1000 out.csv
"""
from __future__ import annotations

import argparse

import numpy as np
from scipy import optimize, signal

TYPES = {"lp": 1, "hp": 2, "bp": 3}


def band_edges(fs: float, f0: float, q: float):
    """f1, f2 of findIIRCutoffFreq.m: the -3 dB point below f0 of the prototype response, then f2 = f0/Q + f1."""
    theta0 = 2 * np.pi * f0 / fs
    k = np.tan(theta0 / (2 * q))

    def g(x):
        s = np.sin(x) * k
        return s / np.sqrt(s * s + (np.cos(x) - np.cos(theta0)) ** 2) - 1 / np.sqrt(2)

    theta1 = optimize.brentq(g, 1e-12, theta0, xtol=1e-16, rtol=8.9e-16, maxiter=500)
    f1 = theta1 * fs / (2 * np.pi)
    return f1, f0 / q + f1


def impulse_response(ftype: str, fs: float, f0: float, q: float, order: int = 8, n: int = 1000) -> np.ndarray:
    """Impulse response of the order-`order` Butterworth filter the reference's designers target
    (casc_2o_iir.h:82-194 with m_t = order/2 sections)."""
    x = np.zeros(n)
    x[0] = 1.0
    if ftype == "bp":
        f1, f2 = band_edges(fs, f0, q)
        sos = signal.butter(order // 2, [f1 / (fs / 2), f2 / (fs / 2)], btype="bandpass", output="sos")
    else:
        sos = signal.butter(order, f0 / (fs / 2), btype="low" if ftype == "lp" else "high", output="sos")
    return signal.sosfilt(sos, x)


def csv_line(ftype: str, fs: float, f0: float, q: float, h: np.ndarray) -> str:
    """The reference's one-line format, parsed by test/testIIR.cpp:7-28."""
    head = [TYPES[ftype], fs, f0, q, len(h)]
    return ",".join([format(v, ".15g") for v in head] + [format(v, ".15e") for v in h]) + "\n"


def parse_csv_line(line: str):
    """Inverse of csv_line (and of the reference's files): -> (type, fs, f0, Q, h)."""
    v = np.array([float(f) for f in line.strip().split(",")])
    n = int(v[4])
    assert v.size == 5 + n, "field count does not match n"
    return int(v[0]), v[1], v[2], v[3], v[5:]


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--type", choices=sorted(TYPES), required=True)
    ap.add_argument("--fs", type=float, default=39e3)
    ap.add_argument("--f0", type=float, required=True)
    ap.add_argument("--q", type=float, default=1.0)
    ap.add_argument("--order", type=int, default=8)
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("out")
    a = ap.parse_args()
    h = impulse_response(a.type, a.fs, a.f0, a.q, a.order, a.n)
    with open(a.out, "w") as fh:
        fh.write(csv_line(a.type, a.fs, a.f0, a.q, h))


if __name__ == "__main__":
    main()
