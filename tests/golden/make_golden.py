#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the reference (run in the build container only).

Needs /root/reference (fixtures + headers) and oracle/_ref/libsdsp_ref.so (``make -C oracle ref``).
Nothing here runs on the GPU box: the committed .npz files are what the tests read.

* impulse_response.npz -- the reference's nine golden impulse responses
  (test_data/impulse_response/*.csv: type,fs,f0,Q,n,h[0..n-1]) re-packed losslessly as float64.
* ref_vectors.npz      -- seeded inputs and the outputs the UNMODIFIED reference headers produce
  for them (FFT radix-2/radix-4 forward/reverse, swap tables, IIR noise responses, preload).
  Inputs are fp32-representable so the same vectors pin the fp32 kernels
  (fp32 oracle = fp64 reference on fp32-rounded input).
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import oracle as O  # noqa: E402

REFERENCE = os.environ.get("SDSP_REFERENCE", "/root/reference")


def impulse_responses():
    out = {}
    names = []
    for path in sorted(glob.glob(os.path.join(REFERENCE, "test_data", "impulse_response", "*.csv"))):
        name = os.path.splitext(os.path.basename(path))[0]
        with open(path) as fh:
            fields = fh.read().strip().split(",")
        v = np.array([float(f) for f in fields], dtype=np.float64)
        assert int(v[4]) == v.size - 5
        out[name + "_header"] = v[:5]  # type, fs, f0, Q, n
        out[name + "_h"] = v[5:]
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "impulse_response.npz"), **out)
    print("impulse_response.npz:", names)


def f32_noise(rng, shape):
    return rng.standard_normal(shape).astype(np.float32).astype(np.float64)


def ref_vectors():
    rng = np.random.default_rng(1234)
    out = {}
    for n in (4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
        frames = 2 if n <= 256 else 1
        x = f32_noise(rng, (frames, n)) + 1j * f32_noise(rng, (frames, n))
        out[f"fft_in_{n}"] = x
        for radix in (2, 4):
            if radix == 4 and (n.bit_length() - 1) % 2:
                continue
            out[f"fft_r{radix}_fwd_{n}"] = O.fft(x, radix, False, "reference")
            out[f"fft_r{radix}_inv_{n}"] = O.fft(x, radix, True, "reference")
            out[f"swap_b{radix}_{n}"] = O.swap_lookup(n, radix, "reference")
    # 65536-point tables are 256 KiB each: keep a checksum of the reference's table instead
    for n in (16384, 65536):  # calc_swap_lookup evaluated at run time by the shim
        for base in (2, 4):
            t = O.swap_lookup(n, base, "reference").astype(np.uint64)
            out[f"swapsum_b{base}_{n}"] = np.array([t.sum(), (t * (np.arange(n, dtype=np.uint64) + 1)).sum()],
                                                   dtype=np.uint64)

    fs = 100e3
    x = f32_noise(rng, (2, 1024))
    out["iir_in"] = x
    cases = []
    for sections in (2, 4, 6, 8):
        for kind in ("generic", "lp", "hp", "bp"):
            for ftype, f0 in ((1, 10e3), (2, 10e3), (3, 10e3), (1, 500.0)):
                want = {1: "lp", 2: "hp", 3: "bp"}[ftype]
                if kind != "generic" and kind != want:
                    continue
                f = O.Iir(sections, kind, "reference")
                f.design(ftype, f0, fs, 1.1, 1.0 if f0 > 1e3 else 0.75)
                key = f"iir_m{sections}_{kind}_t{ftype}_f{int(f0)}"
                # streamed in two unequal pieces through one object: state carries across calls
                y = np.concatenate([f.process(x[0, :300]), f.process(x[0, 300:])])
                out[key] = y
                cases.append(key)
    out["iir_cases"] = np.array(cases)
    # preload (casc_2o_iir.h:196-214): constant input after preload_filter
    for ftype in (1, 2, 3):
        f = O.Iir(4, "generic", "reference")
        f.design(ftype, 10e3, fs, 1.1)
        f.preload_filter(10.0)
        out[f"iir_preload_t{ftype}"] = f.process(np.full(1024, 10.0))
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("ref_vectors.npz:", len(out), "arrays,", O.ref_lib().sdsp_ref_describe().decode())


if __name__ == "__main__":
    impulse_responses()
    ref_vectors()
