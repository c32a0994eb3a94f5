"""simpledsp_b200 -- B200 (sm_100a) implementation of simpledsp's two hot paths, batched.

The product is ``lib/libsdsp_b200.so`` (C ABI: ``include/sdsp_b200.h``) plus the header-only C++ drop-in
under ``include/sdsp/``.  This Python package is the thin host-side mirror used by the tests and the
benchmark: same names and argument meaning as the reference's ``sdsp::`` API (``fft_radix2``,
``fft_radix4``, ``casc_2o_iir`` with ``set_lp_coeff`` / ``set_hp_coeff`` / ``set_bp_coeff`` /
``preload_filter`` / ``copy_coeff_from`` / ``process``), plus the batched handles (``FftPlan``, ``IirBank``).
Nothing here computes: every call goes through the C ABI to the CUDA kernels and fails loudly when the
library or the device is missing.
"""
from . import _capi
from ._capi import (BAND_PASS, F32, F64, FORWARD, HIGH_PASS, IIR_AUTO, IIR_SCAN, IIR_SCAN_LOOKBACK, IIR_SCAN_SPLIT, IIR_SEQUENTIAL, LOW_PASS,
                    NUM_BP, NUM_GENERIC, NUM_HP, NUM_LP, REVERSE, SdspError)
from .fft import FftPlan, digit_reverse_permute, digit_reverse_table, fft_radix2, fft_radix4
from .iir import IirBank, casc_2o_iir, casc_2o_iir_bp, casc_2o_iir_hp, casc_2o_iir_lp, design

__all__ = [
    "FftPlan", "fft_radix2", "fft_radix4", "digit_reverse_table", "digit_reverse_permute",
    "IirBank", "casc_2o_iir", "casc_2o_iir_lp", "casc_2o_iir_hp", "casc_2o_iir_bp", "design",
    "SdspError", "F32", "F64", "FORWARD", "REVERSE", "LOW_PASS", "HIGH_PASS", "BAND_PASS",
    "NUM_GENERIC", "NUM_LP", "NUM_HP", "NUM_BP", "IIR_AUTO", "IIR_SEQUENTIAL", "IIR_SCAN", "IIR_SCAN_LOOKBACK", "IIR_SCAN_SPLIT",
]
