"""Turning numpy arrays / torch tensors into (pointer, ptr_kind, stream) for the C ABI."""
from __future__ import annotations

import numpy as np

from . import _capi as K


def is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def describe(x, dtypes):
    """-> (pointer, ptr_kind, stream, precision, device).  x must be C-contiguous and one of `dtypes`
    (a dict numpy-dtype-name -> precision)."""
    if is_torch(x):
        import torch

        name = str(x.dtype).replace("torch.", "")
        if name not in dtypes:
            raise TypeError(f"unsupported dtype {x.dtype}; expected one of {sorted(dtypes)}")
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous (the transform runs in place)")
        if x.is_cuda:
            stream = torch.cuda.current_stream(x.device).cuda_stream
            return x.data_ptr(), K.PTR_DEVICE, stream, dtypes[name], x.device.index or 0
        return x.data_ptr(), K.PTR_HOST, None, dtypes[name], None
    if not isinstance(x, np.ndarray):
        raise TypeError("expected a numpy array or a torch tensor")
    if x.dtype.name not in dtypes:
        raise TypeError(f"unsupported dtype {x.dtype}; expected one of {sorted(dtypes)}")
    if not x.flags.c_contiguous or not x.flags.writeable:
        raise ValueError("array must be C-contiguous and writeable (the transform runs in place)")
    return x.ctypes.data, K.PTR_HOST, None, dtypes[x.dtype.name], None
