"""ctypes binding of libsdsp_b200.so (the C ABI declared in include/sdsp_b200.h).

The library is the product; this module only loads it.  There is no fallback of any kind: if the
shared object is missing, or a compute call is made without a usable sm_100 device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SDSP_B200_LIB: kernel-tuning aid only (benchmarking alternative builds of the same library)
LIB_PATH = os.environ.get("SDSP_B200_LIB") or os.path.join(HERE, "lib", "libsdsp_b200.so")

# enums of include/sdsp_b200.h
OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_OOM, ERR_NO_DEVICE = range(6)
F32, F64 = 0, 1
FORWARD, REVERSE = 0, 1
PTR_HOST, PTR_DEVICE = 0, 1
FILTER_NONE, LOW_PASS, HIGH_PASS, BAND_PASS = 0, 1, 2, 3
NUM_GENERIC, NUM_LP, NUM_HP, NUM_BP = 0, 1, 2, 3
IIR_AUTO, IIR_SEQUENTIAL, IIR_SCAN, IIR_SCAN_LOOKBACK, IIR_SCAN_SPLIT = 0, 1, 2, 3, 4

_vp, _dp, _u32p, _sz = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.c_size_t

# every symbol include/sdsp_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "sdsp_b200_version": (C.c_int, []),
    "sdsp_b200_last_error": (C.c_char_p, []),
    "sdsp_b200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "sdsp_b200_init": (C.c_int, [C.c_int]),
    "sdsp_b200_shutdown": (C.c_int, []),
    "sdsp_b200_host_alloc": (C.c_int, [C.POINTER(_vp), _sz]),
    "sdsp_b200_host_free": (C.c_int, [_vp]),
    "sdsp_b200_device_alloc": (C.c_int, [C.POINTER(_vp), _sz, C.c_int]),
    "sdsp_b200_device_free": (C.c_int, [_vp, C.c_int]),
    "sdsp_b200_memcpy": (C.c_int, [_vp, _vp, _sz, C.c_int]),
    "sdsp_b200_device_synchronize": (C.c_int, [C.c_int]),
    "sdsp_b200_fft_plan_create": (C.c_int, [C.POINTER(_vp), C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int]),
    "sdsp_b200_fft_plan_destroy": (C.c_int, [_vp]),
    "sdsp_b200_fft_exec": (C.c_int, [_vp, _vp, _sz, C.c_int, _vp]),
    "sdsp_b200_fft_exec_real": (C.c_int, [_vp, _vp, _vp, _sz, C.c_int, _vp]),
    "sdsp_b200_fft_exec_r2c": (C.c_int, [_vp, _vp, _vp, _sz, C.c_int, _vp]),
    "sdsp_b200_fft_exec_c2r": (C.c_int, [_vp, _vp, _vp, _sz, C.c_int, _vp]),
    "sdsp_b200_fft_plan_describe": (C.c_int, [_vp, C.c_char_p, _sz]),
    "sdsp_b200_fft_plan_launches": (C.c_int, [_vp, _sz, C.POINTER(C.c_int)]),
    "sdsp_b200_twiddle_table": (C.c_int, [C.c_uint32, C.c_int, _dp]),
    "sdsp_b200_digit_reverse_table": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, _u32p, C.c_int]),
    "sdsp_b200_digit_reverse_permute": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_int, _sz, C.c_int, C.c_int, _vp]),
    "sdsp_b200_iir_bank_create": (C.c_int, [C.POINTER(_vp), C.c_int, _sz, C.c_int, C.c_int, C.c_int]),
    "sdsp_b200_iir_bank_destroy": (C.c_int, [_vp]),
    "sdsp_b200_iir_bank_set_coeffs": (C.c_int, [_vp, _sz, _sz, _dp, _dp, _dp]),
    "sdsp_b200_iir_bank_set_state": (C.c_int, [_vp, _sz, _sz, _dp]),
    "sdsp_b200_iir_bank_get_state": (C.c_int, [_vp, _sz, _sz, _dp]),
    "sdsp_b200_iir_bank_reset_state": (C.c_int, [_vp]),
    "sdsp_b200_shard_range": (C.c_int, [_sz, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "sdsp_b200_iir_bank_set_state_diff": (C.c_int, [_vp, _sz, _sz, _dp]),
    "sdsp_b200_iir_bank_get_state_diff": (C.c_int, [_vp, _sz, _sz, _dp]),
    "sdsp_b200_iir_bank_process": (C.c_int, [_vp, _vp, _sz, _sz, C.c_int, C.c_int, _vp]),
    "sdsp_b200_iir_bank_describe": (C.c_int, [_vp, _sz, _sz, C.c_int, C.c_char_p, _sz]),
    "sdsp_b200_iir_design_lp": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp]),
    "sdsp_b200_iir_design_hp": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp]),
    "sdsp_b200_iir_design_bp": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp]),
    "sdsp_b200_iir_preload_state": (C.c_int, [C.c_int, C.c_int, C.c_double, _dp, _dp, C.c_double, _dp]),
    "sdsp_b200_iir_process_once": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, _dp, _dp, _dp, _vp, _sz, C.c_int]),
    "sdsp_b200_debug_emulate_fft": (C.c_int, [C.c_uint32, C.c_int, C.c_int, _vp, _sz]),
    "sdsp_b200_debug_emulate_r2c": (C.c_int, [C.c_uint32, C.c_int, C.c_int, _vp, _vp, _sz]),
    "sdsp_b200_debug_emulate_iir": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, _dp, _dp, _dp, _vp, _sz]),
    "sdsp_b200_debug_emulate_iir_diff": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, _dp, _dp, _dp, _dp, _vp, _sz]),
    "sdsp_b200_debug_iir_decay_length": (C.c_int, [C.c_int, C.c_int, C.c_int, _dp, _dp, C.POINTER(C.c_ulonglong)]),
    "sdsp_b200_debug_fft_queue_item": (C.c_int, [C.c_uint, C.c_int, C.c_ulonglong, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sdsp_b200_debug_fft_real_queue_item": (C.c_int, [C.c_int, C.c_ulonglong, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "sdsp_b200_debug_emulate_iir_scan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, _dp, _dp, _dp, _vp, _sz,
                                                   C.c_int, C.c_int]),
}


class SdspError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libsdsp_b200 status {status}: {message}")
        self.status = status


_lib = None


def lib():
    """Load libsdsp_b200.so (once).  Raises if it has not been built -- there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C simpledsp_b200/csrc`).  simpledsp_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI and the header drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status != OK:
        raise SdspError(status, lib().sdsp_b200_last_error().decode(errors="replace"))
