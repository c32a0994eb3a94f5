"""Cascaded biquad IIR: host-side mirror of sdsp::casc_2o_iir<m_t> and its fixed-numerator siblings
(reference include/sdsp/casc_2o_iir.h:8-468), plus the batched channel bank."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as K
from ._buffers import describe

_REAL = {"float32": K.F32, "float64": K.F64}
_dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_dp)


def design(filter_type: int, sections: int, f0: float, fs: float, q: float = 1.0, gain: float = 1.0):
    """Butterworth designers of the reference (set_lp_coeff / set_hp_coeff / set_bp_coeff,
    casc_2o_iir.h:168-194 / 140-166 / 82-138) -> (gain, b[sections][3], a[sections][3])."""
    g = C.c_double()
    b = np.zeros((sections, 3))
    a = np.zeros((sections, 3))
    L = K.lib()
    if filter_type == K.LOW_PASS:
        K.check(L.sdsp_b200_iir_design_lp(sections, f0, fs, gain, C.byref(g), _p(b), _p(a)))
    elif filter_type == K.HIGH_PASS:
        K.check(L.sdsp_b200_iir_design_hp(sections, f0, fs, gain, C.byref(g), _p(b), _p(a)))
    elif filter_type == K.BAND_PASS:
        K.check(L.sdsp_b200_iir_design_bp(sections, f0, fs, q, gain, C.byref(g), _p(b), _p(a)))
    else:
        raise ValueError("filter_type must be LOW_PASS, HIGH_PASS or BAND_PASS")
    return g.value, b, a


class IirBank:
    """n_channels independent filter objects resident on the device (``sdsp_b200_iir_bank_*``)."""

    def __init__(self, sections: int, n_channels: int, precision: int = K.F32, numerator: int = K.NUM_GENERIC, device: int = 0):
        self._h = C.c_void_p()
        K.check(K.lib().sdsp_b200_iir_bank_create(C.byref(self._h), sections, n_channels, precision, numerator, device))
        self.sections, self.n_channels, self.precision, self.numerator, self.device = sections, n_channels, precision, numerator, device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib = getattr(K, "lib", None)  # (module globals are already gone when this runs at interpreter exit)
            if lib is not None:
                lib().sdsp_b200_iir_bank_destroy(self._h)
                self._h = C.c_void_p()

    __del__ = close

    def set_coeffs(self, gain, b, a, first: int = 0):
        gain = np.ascontiguousarray(gain, dtype=np.float64).reshape(-1)
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(gain.size, self.sections, 3)
        bp = None
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.float64).reshape(gain.size, self.sections, 3)
            bp = _p(b)
        K.check(K.lib().sdsp_b200_iir_bank_set_coeffs(self._h, first, gain.size, _p(gain), bp, _p(a)))

    def set_state(self, mem, first: int = 0):
        mem = np.ascontiguousarray(mem, dtype=np.float64).reshape(-1, self.sections + 1, 2)
        K.check(K.lib().sdsp_b200_iir_bank_set_state(self._h, first, mem.shape[0], _p(mem)))

    def get_state(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n_channels - first if count is None else count
        mem = np.zeros((count, self.sections + 1, 2))
        K.check(K.lib().sdsp_b200_iir_bank_get_state(self._h, first, count, _p(mem)))
        return mem

    def set_state_diff(self, diff, first: int = 0):
        """fp32 banks only carry these (difference-form recurrence): with get_state()/set_state() an exact checkpoint."""
        diff = np.ascontiguousarray(diff, dtype=np.float64).reshape(-1, self.sections)
        K.check(K.lib().sdsp_b200_iir_bank_set_state_diff(self._h, first, diff.shape[0], _p(diff)))

    def get_state_diff(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n_channels - first if count is None else count
        diff = np.zeros((count, self.sections))
        K.check(K.lib().sdsp_b200_iir_bank_get_state_diff(self._h, first, count, _p(diff)))
        return diff

    def reset_state(self):
        K.check(K.lib().sdsp_b200_iir_bank_reset_state(self._h))

    def describe(self, n_samples: int, channel_stride: int | None = None, path: int = K.IIR_AUTO) -> str:
        buf = C.create_string_buffer(1024)
        K.check(K.lib().sdsp_b200_iir_bank_describe(self._h, n_samples, channel_stride or n_samples, path, buf, len(buf)))
        return buf.value.decode()

    def process_ptr(self, ptr: int, n_samples: int, channel_stride: int, ptr_kind: int, path: int = K.IIR_AUTO, stream=None):
        K.check(K.lib().sdsp_b200_iir_bank_process(self._h, ptr, n_samples, channel_stride, ptr_kind, path, stream))

    def process(self, data, path: int = K.IIR_AUTO):
        """Filter data[channel, :] in place; history carries over to the next call.  data: numpy
        (staged) or torch CUDA tensor of shape [n_channels, n_samples] (or [n_samples] for one channel)."""
        ptr, kind, stream, prec, dev = describe(data, _REAL)
        if prec != self.precision:
            raise TypeError("dtype does not match the bank's precision")
        if kind == K.PTR_DEVICE and dev is not None and dev != self.device:
            raise ValueError(f"tensor lives on cuda:{dev}, the bank on cuda:{self.device}")
        shape = tuple(int(s) for s in data.shape)
        if len(shape) == 1:
            shape = (1,) + shape
        if len(shape) != 2 or shape[0] != self.n_channels:
            raise ValueError(f"expected shape [{self.n_channels}, n_samples]")
        self.process_ptr(ptr, shape[1], shape[1], kind, path, stream)
        return data


class casc_2o_iir:
    """One filter object with the reference's interface (casc_2o_iir.h:8-215); state lives on the host
    between calls, exactly as the reference object carries m_mem, and every process() runs on the GPU."""

    _numerator = K.NUM_GENERIC

    def __init__(self, sections: int = 4, precision: int = K.F64, device: int = 0):
        if sections % 2:
            raise ValueError("M must be even!")  # casc_2o_iir.h:25
        self.sections, self.precision, self.device = sections, precision, device
        self.gain = 1.0
        self.b = np.zeros((sections, 3))
        self.a = np.zeros((sections, 3))
        self.mem = np.zeros((sections + 1, 2))
        self.f_type = K.FILTER_NONE

    def copy(self):
        o = type(self).__new__(type(self))
        o.sections, o.precision, o.device = self.sections, self.precision, self.device
        o.gain, o.f_type = self.gain, self.f_type
        o.b, o.a, o.mem = self.b.copy(), self.a.copy(), self.mem.copy()
        return o

    def copy_coeff_from(self, other):  # casc_2o_iir.h:28-34: coefficients and type, not history
        self.gain, self.f_type = other.gain, other.f_type
        self.b, self.a = other.b.copy(), other.a.copy()

    def _set(self, ftype, f0, fs, q, gain):
        self.gain, self.b, self.a = design(ftype, self.sections, f0, fs, q, gain)
        self.f_type = ftype

    def set_lp_coeff(self, f0, fs, gain=1.0):
        self._set(K.LOW_PASS, f0, fs, 1.0, gain)

    def set_hp_coeff(self, f0, fs, gain=1.0):
        self._set(K.HIGH_PASS, f0, fs, 1.0, gain)

    def set_bp_coeff(self, f0, fs, q, gain=1.0):
        self._set(K.BAND_PASS, f0, fs, q, gain)

    def preload_filter(self, value):  # casc_2o_iir.h:196-214
        K.check(K.lib().sdsp_b200_iir_preload_state(self.sections, self.f_type, self.gain, _p(self.b), _p(self.a), value, _p(self.mem)))

    def process(self, data):
        """In place over a 1-D numpy array (the reference's process(begin, end))."""
        if not isinstance(data, np.ndarray) or data.ndim != 1:
            raise TypeError("expected a 1-D numpy array")
        ptr, kind, stream, prec, dev = describe(data, _REAL)
        if prec != self.precision:
            raise TypeError("dtype does not match the filter's precision")
        K.check(K.lib().sdsp_b200_iir_process_once(self.sections, self._numerator, prec, self.gain, _p(self.b), _p(self.a),
                                                   _p(self.mem), ptr, data.size, self.device))
        return data


class _fixed(casc_2o_iir):
    def preload_filter(self, value):  # the fixed-numerator classes have no preload (casc_2o_iir.h:266-468)
        raise AttributeError("preload_filter is only defined for casc_2o_iir")


class casc_2o_iir_lp(_fixed):
    _numerator = K.NUM_LP
    set_hp_coeff = set_bp_coeff = None


class casc_2o_iir_hp(_fixed):
    _numerator = K.NUM_HP
    set_lp_coeff = set_bp_coeff = None


class casc_2o_iir_bp(_fixed):
    _numerator = K.NUM_BP
    set_lp_coeff = set_hp_coeff = None
