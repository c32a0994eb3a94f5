"""Batched FFT: host-side mirror of sdsp::fft_radix2 / sdsp::fft_radix4 (reference include/sdsp/fft.h:258-360)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as K
from ._buffers import describe

_COMPLEX = {"complex64": K.F32, "complex128": K.F64}


class FftPlan:
    """A transform of ``n`` points, batched over frames: ``sdsp_b200_fft_plan_*`` of include/sdsp_b200.h.

    radix is the reference's entry point (2 -> fft_radix2, any power of 2; 4 -> fft_radix4, powers of 4).
    direction: FORWARD = forward_fft (fft.h:135-146), REVERSE = reverse_fft with the 1/N scale (:121-133).
    """

    def __init__(self, n: int, radix: int = 2, precision: int = K.F32, direction: int = K.FORWARD, device: int = 0):
        self._h = C.c_void_p()
        K.check(K.lib().sdsp_b200_fft_plan_create(C.byref(self._h), n, radix, precision, direction, device))
        self.n, self.radix, self.precision, self.direction, self.device = n, radix, precision, direction, device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib = getattr(K, "lib", None)  # (module globals are already gone when this runs at interpreter exit)
            if lib is not None:
                lib().sdsp_b200_fft_plan_destroy(self._h)
                self._h = C.c_void_p()

    __del__ = close

    def describe(self) -> str:
        buf = C.create_string_buffer(1024)
        K.check(K.lib().sdsp_b200_fft_plan_describe(self._h, buf, len(buf)))
        return buf.value.decode()

    def launches(self, n_frames: int) -> int:
        out = C.c_int()
        K.check(K.lib().sdsp_b200_fft_plan_launches(self._h, n_frames, C.byref(out)))
        return out.value

    def exec_ptr(self, ptr: int, n_frames: int, ptr_kind: int, stream=None) -> None:
        K.check(K.lib().sdsp_b200_fft_exec(self._h, ptr, n_frames, ptr_kind, stream))

    def exec_real_ptr(self, real_ptr: int, out_ptr: int, n_frames: int, ptr_kind: int, stream=None) -> None:
        """n_frames real frames of n scalars at real_ptr -> n_frames spectra of n complex at out_ptr (out of place)."""
        K.check(K.lib().sdsp_b200_fft_exec_real(self._h, real_ptr, out_ptr, n_frames, ptr_kind, stream))

    def real(self, real_in, out=None):
        """Transform of real frames (imaginary part zero, the way the reference's callers fill their complex arrays,
        test/testFFT.cpp:24): numpy in -> numpy out, torch CUDA tensor in -> torch CUDA tensor out."""
        n = self.n
        if isinstance(real_in, np.ndarray):
            want = np.float32 if self.precision == K.F32 else np.float64
            x = np.ascontiguousarray(real_in, dtype=want)
            if out is None:
                out = np.empty(x.shape, dtype=np.complex64 if self.precision == K.F32 else np.complex128)
            self.exec_real_ptr(x.ctypes.data, out.ctypes.data, x.size // n, K.PTR_HOST, None)
            return out
        import torch

        if out is None:
            out = torch.empty(real_in.shape, device=real_in.device, dtype=torch.complex64 if self.precision == K.F32 else torch.complex128)
        self.exec_real_ptr(real_in.data_ptr(), out.data_ptr(), real_in.numel() // n, K.PTR_DEVICE, torch.cuda.current_stream().cuda_stream)
        return out

    def exec_r2c_ptr(self, real_ptr: int, out_ptr: int, n_frames: int, ptr_kind: int, stream=None) -> None:
        """n_frames real frames of n scalars at real_ptr -> n_frames half spectra of n/2 + 1 complex at out_ptr (out of place)."""
        K.check(K.lib().sdsp_b200_fft_exec_r2c(self._h, real_ptr, out_ptr, n_frames, ptr_kind, stream))

    def half_spectrum(self, real_in, out=None):
        """Bins 0 .. n/2 of the transform of real frames (the rest is the conjugate mirror): shape (..., n) -> (..., n/2 + 1).
        numpy in -> numpy out (staged through the device), torch CUDA tensor in -> torch CUDA tensor out."""
        n = self.n
        shape = tuple(real_in.shape[:-1]) + (n // 2 + 1,)
        if real_in.shape[-1] != n:
            raise ValueError(f"last dimension must be {n}")
        if isinstance(real_in, np.ndarray):
            want = np.float32 if self.precision == K.F32 else np.float64
            x = np.ascontiguousarray(real_in, dtype=want)
            if out is None:
                out = np.empty(shape, dtype=np.complex64 if self.precision == K.F32 else np.complex128)
            self.exec_r2c_ptr(x.ctypes.data, out.ctypes.data, x.size // n, K.PTR_HOST, None)
            return out
        import torch

        if real_in.device.index != self.device:
            raise ValueError("tensor lives on another device than the plan")
        if out is None:
            out = torch.empty(shape, device=real_in.device, dtype=torch.complex64 if self.precision == K.F32 else torch.complex128)
        self.exec_r2c_ptr(real_in.data_ptr(), out.data_ptr(), real_in.numel() // n, K.PTR_DEVICE, torch.cuda.current_stream().cuda_stream)
        return out

    def exec_c2r_ptr(self, half_ptr: int, real_ptr: int, n_frames: int, ptr_kind: int, stream=None) -> None:
        """n_frames half spectra of n/2 + 1 complex at half_ptr -> n_frames real frames of n scalars at real_ptr (reverse plans)."""
        K.check(K.lib().sdsp_b200_fft_exec_c2r(self._h, half_ptr, real_ptr, n_frames, ptr_kind, stream))

    def real_from_half_spectrum(self, half_in, out=None):
        """The way back from half_spectrum (reverse plans, 1/n included): shape (..., n/2 + 1) complex -> (..., n) real."""
        n = self.n
        if half_in.shape[-1] != n // 2 + 1:
            raise ValueError(f"last dimension must be {n // 2 + 1}")
        shape = tuple(half_in.shape[:-1]) + (n,)
        frames = 1
        for d in shape[:-1]:
            frames *= int(d)
        if isinstance(half_in, np.ndarray):
            x = np.ascontiguousarray(half_in, dtype=np.complex64 if self.precision == K.F32 else np.complex128)
            if out is None:
                out = np.empty(shape, dtype=np.float32 if self.precision == K.F32 else np.float64)
            self.exec_c2r_ptr(x.ctypes.data, out.ctypes.data, frames, K.PTR_HOST, None)
            return out
        import torch

        if half_in.device.index != self.device:
            raise ValueError("tensor lives on another device than the plan")
        if out is None:
            out = torch.empty(shape, device=half_in.device, dtype=torch.float32 if self.precision == K.F32 else torch.float64)
        self.exec_c2r_ptr(half_in.data_ptr(), out.data_ptr(), frames, K.PTR_DEVICE, torch.cuda.current_stream().cuda_stream)
        return out

    def __call__(self, data):
        """Transform every length-n row of ``data`` in place (numpy: staged through the device;
        torch CUDA tensor: in place on the current stream, asynchronously).  Returns ``data``."""
        ptr, kind, stream, prec, dev = describe(data, _COMPLEX)
        if prec != self.precision:
            raise TypeError("dtype does not match the plan's precision")
        if dev is not None and dev != self.device:
            raise ValueError("tensor lives on another device than the plan")
        if data.shape[-1] != self.n:
            raise ValueError(f"last dimension must be {self.n}")
        total = 1
        for s in data.shape:
            total *= int(s)
        self.exec_ptr(ptr, total // self.n, kind, stream)
        return data


_plans = {}


def _plan_for(data, radix, inverse):
    ptr, kind, stream, prec, dev = describe(data, _COMPLEX)
    key = (int(data.shape[-1]), radix, prec, K.REVERSE if inverse else K.FORWARD, dev or 0)
    p = _plans.get(key)
    if p is None:
        p = _plans[key] = FftPlan(*key)
    return p


def fft_radix2(data, inverse: bool = False):
    """sdsp::fft_radix2<T>(data) on every row of ``data``, in place (fft.h:258-299)."""
    return _plan_for(data, 2, inverse)(data)


def fft_radix4(data, inverse: bool = False):
    """sdsp::fft_radix4<T>(data) on every row of ``data``, in place (fft.h:301-360)."""
    return _plan_for(data, 4, inverse)(data)


def digit_reverse_table(n: int, base: int, half_table: bool = False, device: int = 0) -> np.ndarray:
    """digit_reverse<N,base> for every index, or calc_swap_lookup<N,base> when half_table
    (fft.h:217-256), computed on the device."""
    out = np.zeros(n, dtype=np.uint32)
    K.check(K.lib().sdsp_b200_digit_reverse_table(n, base, int(half_table), out.ctypes.data_as(C.POINTER(C.c_uint32)), device))
    return out


def digit_reverse_permute(data, base: int, device: int = 0):
    """out[rev(i)] = in[i] on every row, in place: the swap sweep of fft.h:269-273 / 351-355."""
    ptr, kind, stream, prec, dev = describe(data, _COMPLEX)
    n = int(data.shape[-1])
    total = 1
    for s in data.shape:
        total *= int(s)
    K.check(K.lib().sdsp_b200_digit_reverse_permute(ptr, n, base, prec, total // n, kind, dev if dev is not None else device, stream))
    return data
