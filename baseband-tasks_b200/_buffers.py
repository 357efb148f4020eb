"""Device buffers: torch is used for allocation, streams and copies only."""
import ctypes
import sys

import numpy as np

from . import _cabi

_TORCH = None


def torch():
    global _TORCH
    if _TORCH is None:
        import torch as _t
        _TORCH = _t
    return _TORCH


def torch_dtype(dtype):
    return getattr(torch(), np.dtype(dtype).name)


def is_tensor(x):
    t = _TORCH or sys.modules.get('torch')
    return t is not None and isinstance(x, t.Tensor)


def empty(shape, dtype):
    return torch().empty(tuple(shape), dtype=torch_dtype(dtype),
                         device=_cabi.device())


def zeros(shape, dtype):
    return torch().zeros(tuple(shape), dtype=torch_dtype(dtype),
                         device=_cabi.device())


def as_device(x, dtype=None):
    """A contiguous tensor on the compute device (uploads host arrays)."""
    t = torch()
    if not is_tensor(x):
        x = np.ascontiguousarray(x, dtype=dtype)
        if not x.flags.writeable:
            x = x.copy()
        x = t.from_numpy(x)
    elif dtype is not None and x.dtype != torch_dtype(dtype):
        x = x.to(torch_dtype(dtype))
    dev = _cabi.device()
    if x.device != dev:
        x = x.to(dev, non_blocking=True)
    return x.contiguous()


def as_host(x):
    """A numpy array (downloads device tensors)."""
    if is_tensor(x):
        return x.cpu().numpy()
    return np.asarray(x)


def ptr(x, byte_offset=0):
    """Raw pointer to the first element of a contiguous tensor."""
    if x is None:
        return None
    return ctypes.c_void_p(x.data_ptr() + byte_offset)
