"""ctypes bindings for the CPU checkers.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two libraries sit behind this module:

* ``libsdsp_oracle.so``  -- the plain-C restatement (``oracle/sdsp_oracle.c``), kind ``"port"``;
* ``_ref/libsdsp_ref.so`` -- the unmodified reference headers compiled behind ``extern "C"``
  (``oracle/ref_shim.cpp``), kind ``"reference"``.  Built in the build container where
  ``/root/reference`` exists; the prebuilt file travels to the GPU box.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline / --impl reference) may
import this module.  ``simpledsp_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libsdsp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsdsp_ref.so")
REF_BIG_SO = os.path.join(HERE, "_ref", "libsdsp_ref_big.so")

MAX_SECTIONS = 16
_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)


def build(ref: bool = True) -> None:
    """(Re)build the checkers with oracle/Makefile.  The reference target is skipped when
    /root/reference is absent (GPU box): the prebuilt oracle/_ref is used as is."""
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    if ref and os.path.isdir(os.environ.get("SDSP_REFERENCE", "/root/reference") + "/include/sdsp"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


class _IirStruct(C.Structure):
    _fields_ = [
        ("sections", C.c_int),
        ("kind", C.c_int),
        ("pos", C.c_int),
        ("ftype", C.c_int),
        ("gain", C.c_double),
        ("mem", (C.c_double * 3) * (MAX_SECTIONS + 1)),
        ("b", (C.c_double * 3) * MAX_SECTIONS),
        ("a", (C.c_double * 3) * MAX_SECTIONS),
    ]


_port = None
_ref = None


def port_lib():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build(ref=False)
        lib = C.CDLL(PORT_SO)
        lib.sdsp_oracle_digit_reverse.restype = C.c_uint32
        lib.sdsp_oracle_digit_reverse.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        lib.sdsp_oracle_swap_lookup.argtypes = [C.c_uint32, C.c_uint32, _u32p]
        lib.sdsp_oracle_calc_trigs.argtypes = [C.c_uint32, C.c_int, _dp]
        lib.sdsp_oracle_calc_wcoeffs.argtypes = [C.c_uint32, C.c_int, _dp]
        lib.sdsp_oracle_fft_radix2.argtypes = [_dp, C.c_uint32, C.c_int]
        lib.sdsp_oracle_fft_radix4.argtypes = [_dp, C.c_uint32, C.c_int]
        lib.sdsp_oracle_fft_batch.argtypes = [_dp, C.c_uint32, C.c_size_t, C.c_int, C.c_int]
        lib.sdsp_oracle_iir_init.argtypes = [C.POINTER(_IirStruct), C.c_int, C.c_int]
        lib.sdsp_oracle_iir_copy_coeff_from.argtypes = [C.POINTER(_IirStruct), C.POINTER(_IirStruct)]
        lib.sdsp_oracle_iir_set_lp.argtypes = [C.POINTER(_IirStruct), C.c_double, C.c_double, C.c_double]
        lib.sdsp_oracle_iir_set_hp.argtypes = [C.POINTER(_IirStruct), C.c_double, C.c_double, C.c_double]
        lib.sdsp_oracle_iir_set_bp.argtypes = [C.POINTER(_IirStruct), C.c_double, C.c_double, C.c_double, C.c_double]
        lib.sdsp_oracle_iir_preload.argtypes = [C.POINTER(_IirStruct), C.c_double]
        lib.sdsp_oracle_iir_process.argtypes = [C.POINTER(_IirStruct), _dp, C.c_size_t]
        lib.sdsp_oracle_iir_sizeof.restype = C.c_size_t
        assert lib.sdsp_oracle_iir_sizeof() == C.sizeof(_IirStruct)
        _port = lib
    return _port


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref_lib(big: bool = False):
    """The compiled reference.  Raises FileNotFoundError when it has not been built."""
    global _ref
    path = REF_BIG_SO if big else REF_SO
    if big or _ref is None:
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        lib = C.CDLL(path)
        lib.sdsp_ref_describe.restype = C.c_char_p
        lib.sdsp_ref_fft.argtypes = [_dp, C.c_uint32, C.c_int, C.c_int]
        lib.sdsp_ref_fft_batch.argtypes = [_dp, C.c_uint32, C.c_size_t, C.c_int, C.c_int, C.c_int]
        lib.sdsp_ref_swap_lookup.argtypes = [C.c_uint32, C.c_uint32, _u32p]
        lib.sdsp_ref_iir_create.restype = C.c_void_p
        lib.sdsp_ref_iir_create.argtypes = [C.c_int, C.c_int]
        lib.sdsp_ref_iir_destroy.argtypes = [C.c_void_p]
        lib.sdsp_ref_iir_clone.restype = C.c_void_p
        lib.sdsp_ref_iir_clone.argtypes = [C.c_void_p]
        lib.sdsp_ref_iir_set_lp.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        lib.sdsp_ref_iir_set_hp.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        lib.sdsp_ref_iir_set_bp.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
        lib.sdsp_ref_iir_preload.argtypes = [C.c_void_p, C.c_double]
        lib.sdsp_ref_iir_process.argtypes = [C.c_void_p, _dp, C.c_size_t]
        lib.sdsp_ref_iir_bank_run.argtypes = [_dp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                              C.POINTER(C.c_int), _dp, C.c_double, C.c_double, C.c_int]
        if big:
            return lib
        _ref = lib
    return _ref


def _as_frames(x: np.ndarray) -> np.ndarray:
    a = np.array(x, dtype=np.complex128, order="C", copy=True)
    return a


# ----------------------------------------------------------------------------- FFT
def fft(x, radix: int = 2, inverse: bool = False, impl: str = "port", threads: int = 1) -> np.ndarray:
    """Transform the last axis of ``x`` (complex128) with the reference's algorithm.

    impl = "port" (C restatement) or "reference" (compiled reference headers)."""
    a = _as_frames(x)
    n = a.shape[-1]
    frames = a.size // n if n else 0
    p = a.ctypes.data_as(_dp)
    if impl == "port":
        rc = port_lib().sdsp_oracle_fft_batch(p, n, frames, radix, int(inverse))
    elif impl == "reference":
        lib = ref_lib(big=n > 4096)
        rc = lib.sdsp_ref_fft_batch(p, n, frames, radix, int(inverse), threads)
    else:
        raise ValueError(impl)
    if rc != 0:
        raise ValueError(f"oracle fft: n={n} radix={radix} not supported by impl={impl} (rc={rc})")
    return a


def digit_reverse(n: int, base: int) -> np.ndarray:
    """rev(i) for every i < n -- fft.h:217-236."""
    lib = port_lib()
    return np.array([lib.sdsp_oracle_digit_reverse(n, base, i) for i in range(n)], dtype=np.uint32)


def swap_lookup(n: int, base: int, impl: str = "port") -> np.ndarray:
    """The half swap table of fft.h:238-256."""
    out = np.zeros(n, dtype=np.uint32)
    if impl == "port":
        rc = port_lib().sdsp_oracle_swap_lookup(n, base, out.ctypes.data_as(_u32p))
    else:
        rc = ref_lib().sdsp_ref_swap_lookup(n, base, out.ctypes.data_as(_u32p))
    if rc != 0:
        raise ValueError(f"swap_lookup: n={n} base={base} unsupported (rc={rc})")
    return out


def wcoeffs(n: int, inverse: bool = False) -> np.ndarray:
    rows = int(n).bit_length() - 1
    out = np.zeros((rows, n), dtype=np.complex128)
    rc = port_lib().sdsp_oracle_calc_wcoeffs(n, int(inverse), out.ctypes.data_as(_dp))
    if rc != 0:
        raise ValueError(n)
    return out


# ----------------------------------------------------------------------------- IIR
KINDS = {"generic": 0, "lp": 1, "hp": 2, "bp": 3}


class Iir:
    """One cascaded-biquad filter object (casc_2o_iir<m_t> or its fixed-numerator siblings)."""

    def __init__(self, sections: int = 4, kind: str = "generic", impl: str = "port"):
        self.impl = impl
        self.sections = sections
        self.kind = KINDS[kind]
        if impl == "port":
            self._s = _IirStruct()
            if port_lib().sdsp_oracle_iir_init(C.byref(self._s), sections, self.kind) != 0:
                raise ValueError((sections, kind))
        else:
            self._h = ref_lib().sdsp_ref_iir_create(sections, self.kind)
            if not self._h:
                raise ValueError((sections, kind))

    def __del__(self):
        if getattr(self, "impl", None) == "reference" and getattr(self, "_h", None):
            ref_lib().sdsp_ref_iir_destroy(self._h)
            self._h = None

    def copy(self) -> "Iir":
        o = Iir.__new__(Iir)
        o.impl, o.sections, o.kind = self.impl, self.sections, self.kind
        if self.impl == "port":
            o._s = _IirStruct()
            C.memmove(C.byref(o._s), C.byref(self._s), C.sizeof(_IirStruct))
        else:
            o._h = ref_lib().sdsp_ref_iir_clone(self._h)
        return o

    def _call(self, name, *args):
        if self.impl == "port":
            rc = getattr(port_lib(), "sdsp_oracle_iir_" + name)(C.byref(self._s), *args)
        else:
            rc = getattr(ref_lib(), "sdsp_ref_iir_" + name)(self._h, *args)
        if rc not in (0, None):
            raise ValueError(f"{name} not available for this filter class")

    def set_lp_coeff(self, f0, fs, gain=1.0):
        self._call("set_lp", f0, fs, gain)

    def set_hp_coeff(self, f0, fs, gain=1.0):
        self._call("set_hp", f0, fs, gain)

    def set_bp_coeff(self, f0, fs, q, gain=1.0):
        self._call("set_bp", f0, fs, q, gain)

    def preload_filter(self, value):
        self._call("preload", value)

    def design(self, ftype: int, f0: float, fs: float, q: float = 1.0, gain: float = 1.0):
        if ftype == 1:
            self.set_lp_coeff(f0, fs, gain)
        elif ftype == 2:
            self.set_hp_coeff(f0, fs, gain)
        elif ftype == 3:
            self.set_bp_coeff(f0, fs, q, gain)
        else:
            raise ValueError(ftype)

    def process(self, x) -> np.ndarray:
        a = np.array(x, dtype=np.float64, order="C", copy=True)
        p = a.ctypes.data_as(_dp)
        if self.impl == "port":
            port_lib().sdsp_oracle_iir_process(C.byref(self._s), p, a.size)
        else:
            ref_lib().sdsp_ref_iir_process(self._h, p, a.size)
        return a

    # port only: the coefficients the designers produced, for uploading to the GPU bank
    def coefficients(self):
        assert self.impl == "port"
        m = self.sections
        b = np.array([[self._s.b[j][i] for i in range(3)] for j in range(m)])
        a = np.array([[self._s.a[j][i] for i in range(3)] for j in range(m)])
        return float(self._s.gain), b, a


def iir_bank_reference(x, ftype, f0, fs, q=1.0, sections=4, kind="generic", threads=1) -> np.ndarray:
    """Planar [channels][n] float64 run through one compiled-reference filter object per channel."""
    a = np.array(x, dtype=np.float64, order="C", copy=True)
    ch, n = a.shape
    ft = np.ascontiguousarray(ftype, dtype=np.int32)
    ff = np.ascontiguousarray(f0, dtype=np.float64)
    rc = ref_lib().sdsp_ref_iir_bank_run(a.ctypes.data_as(_dp), ch, n, n, sections, KINDS[kind],
                                         ft.ctypes.data_as(C.POINTER(C.c_int)), ff.ctypes.data_as(_dp),
                                         fs, q, threads)
    if rc != 0:
        raise ValueError("reference bank run failed")
    return a


def iir_bank_port(x, ftype, f0, fs, q=1.0, sections=4, kind="generic") -> np.ndarray:
    a = np.array(x, dtype=np.float64, order="C", copy=True)
    for c in range(a.shape[0]):
        f = Iir(sections, kind, "port")
        f.design(int(ftype[c]), float(f0[c]), fs, q)
        a[c] = f.process(a[c])
    return a
