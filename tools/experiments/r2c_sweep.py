"""Throughput of the half-spectrum calls over the frame sizes (device-resident, CUDA events): real samples/s and GB/s at 8 B/sample (f32)."""
import sys

import numpy as np
import torch

import simpledsp_b200 as S
from simpledsp_b200 import _capi as K

total = 1 << 28  # real samples per call
print("# n  r2c ms  Msamples/s  GB/s(8B)  |  c2r ms  Msamples/s  GB/s  | rel err vs torch.fft.rfft / irfft")
for lg in range(2, 17):
    n = 1 << lg
    frames = total // n
    x = torch.randn(frames, n, device="cuda", dtype=torch.float32)
    fwd = S.FftPlan(n, 2, K.F32, K.FORWARD)
    y = fwd.half_spectrum(x)
    row = [n]
    ref = torch.fft.rfft(x[:4].double(), dim=1)
    err = float(((y[:4].to(torch.complex128) - ref).abs().pow(2).sum().sqrt() / ref.abs().pow(2).sum().sqrt()))
    def timeit(fn, reps=10):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    ms = timeit(lambda: fwd.half_spectrum(x, out=y))
    row += [round(ms, 4), round(total / ms / 1e3), round(total * 8 / ms / 1e6, 1)]
    if n <= 32768:
        inv = S.FftPlan(n, 2, K.REVERSE and K.F32, K.REVERSE) if False else S.FftPlan(n, 2, K.F32, K.REVERSE)
        z = inv.real_from_half_spectrum(y)
        err2 = float((z[:4].double() - x[:4].double()).pow(2).sum().sqrt() / x[:4].double().pow(2).sum().sqrt())
        ms2 = timeit(lambda: inv.real_from_half_spectrum(y, out=z))
        row += ["|", round(ms2, 4), round(total / ms2 / 1e3), round(total * 8 / ms2 / 1e6, 1), "|", "%.2e" % err, "%.2e" % err2]
    else:
        row += ["|", "-", "-", "-", "|", "%.2e" % err]
    print(*row)
    del x, y
    torch.cuda.empty_cache()
