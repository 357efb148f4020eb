"""The 'cuda' FFT maker: hand-written sm_100a Stockham kernels via the C ABI.

Registered as ``'cuda'`` through `FFTMakerMeta`, exactly as the reference's
makers register themselves (fourier/base.py:235-253), and selectable with
``fft_maker.set('cuda')``.  Replaces `NumpyFFTBase` (fourier/numpy.py:13-49):
unnormalised forward transform, 1/n on the inverse, 1/sqrt(n) both ways with
``ortho`` (fourier/base.py:95-104); real transforms keep n//2+1 bins.

Any length and axis is transformed on the GPU: powers of two up to 16384 in
one kernel, larger ones in four steps, other lengths through Bluestein's
algorithm on a power-of-two length (about three times the work, so
``next_fast_len`` rounds up to a power of two and padded tasks choose such
lengths themselves, dispersion.py:99, base.py:757-758); real transforms beyond
the single-kernel sizes go through the complex transform.  Arithmetic is
single precision (double-precision input is converted, with a warning).  There
is no CPU fallback.
"""
import ctypes
import math
import warnings

import numpy as np

from .. import _buffers as B
from .. import _cabi
from .base import FFTBase, FFTMakerBase

__all__ = ['CudaFFTBase', 'CudaFFTMaker']

_SINGLE = {'f': np.dtype('f4'), 'c': np.dtype('c8')}


class CudaFFTBase(FFTBase):
    """Single pre-defined FFT running on the GPU.

    Accepts device tensors (returns a device tensor, data stay in HBM) or
    numpy arrays (uploaded, transformed, downloaded).
    """

    def __init__(self, direction='forward'):
        super().__init__(direction=direction)
        self._plan = None
        self._work = None
        self._work_bytes = 0
        shape, axis = self._time_shape, self._axis % len(self._time_shape)
        self._n = shape[axis]
        self._outer = int(np.prod(shape[:axis], dtype=np.int64))
        self._inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
        real = self._time_dtype.kind == 'f'
        forward = self.direction == 'forward'
        if real:
            self._kind = _cabi.BBT_R2C if forward else _cabi.BBT_C2R
        else:
            self._kind = _cabi.BBT_C2C
        if self._ortho:
            self._scale = 1. / math.sqrt(self._n)
        else:
            self._scale = 1. if forward else 1. / self._n
        if forward:
            self._in_shape, self._in_dtype = self._time_shape, self._time_dtype
            self._out_shape, self._out_dtype = (self._frequency_shape,
                                                self._frequency_dtype)
        else:
            self._in_shape, self._in_dtype = (self._frequency_shape,
                                              self._frequency_dtype)
            self._out_shape, self._out_dtype = (self._time_shape,
                                                self._time_dtype)

    def _get_plan(self):
        if self._plan is None:
            lib = _cabi.lib()
            plan = ctypes.c_void_p()
            outer, inner = self._outer, self._inner
            lib.check(lib.bbt_fft_plan_create(
                ctypes.byref(plan), self._n, outer, inner, self._kind,
                _cabi.BBT_BACKWARD if self.direction == 'backward'
                else _cabi.BBT_FORWARD, self._scale))
            self._plan = plan
            self._work_bytes = lib.bbt_fft_plan_work_bytes(plan)
        return self._plan

    def _get_work(self):
        """Scratch of the large transforms, allocated once per FFT object."""
        if not self._work_bytes:
            return None
        if self._work is None or self._work.device != _cabi.device():
            self._work = B.empty((self._work_bytes,), np.uint8)
        return self._work

    def __call__(self, a, out=None):
        """Transform ``a``; ``out`` may name a device tensor to fill."""
        return self._fft(a, out=out)

    def _fft(self, a, out=None):
        host = not B.is_tensor(a)
        if tuple(a.shape) != tuple(self._in_shape):
            raise ValueError(f"input has shape {tuple(a.shape)}, transform "
                             f"was set up for {tuple(self._in_shape)}.")
        single_in = _SINGLE[self._in_dtype.kind]
        single_out = _SINGLE[self._out_dtype.kind]
        if self._in_dtype != single_in:
            warnings.warn("the cuda FFT maker computes in single precision; "
                          f"{self._in_dtype} data are converted.",
                          stacklevel=3)
        x = B.as_device(a, dtype=single_in)
        lib = _cabi.lib()
        plan = self._get_plan()
        if (out is None or not out.is_contiguous()
                or out.dtype != B.torch_dtype(single_out)
                or out.numel() != int(np.prod(self._out_shape,
                                              dtype=np.int64))):
            out = B.empty(self._out_shape, single_out)
            given = False
        else:
            given = True
        lib.check(lib.bbt_fft_exec(plan, B.ptr(x), B.ptr(out),
                                   B.ptr(self._get_work()),
                                   _cabi.stream_ptr()))
        if given:
            return out
        if self._out_dtype != single_out:
            out = out.to(B.torch_dtype(self._out_dtype))
        return B.as_host(out) if host else out

    def __del__(self):
        plan, self._plan = getattr(self, '_plan', None), None
        if plan is not None:
            try:
                _cabi.lib().bbt_fft_plan_destroy(plan)
            except Exception:
                pass


def smooth_fast_len(n):
    """Smallest length >= ``n`` with no prime factor above 7 (``n`` itself up
    to 7): what the reference's numpy maker pads frames to
    (fourier/numpy.py:99-126), so tasks framed with it read exactly the
    samples per frame the reference would."""
    n = int(n)
    if n <= 7:
        return n
    best = 1 << (n - 1).bit_length()          # a power of two always qualifies
    p7 = 1
    while p7 < best:
        p57 = p7
        while p57 < best:
            p357 = p57
            while p357 < best:
                # Fill up with twos.
                rest = -(-n // p357)
                candidate = p357 << max(0, (rest - 1).bit_length())
                if n <= candidate < best:
                    best = candidate
                p357 *= 3
            p57 *= 5
        p7 *= 7
    return best


class CudaFFTMaker(FFTMakerBase):
    """FFT factory for the hand-written CUDA kernels (key ``'cuda'``).

    ``fast_len='pow2'`` (default): padded tasks round their frames up to a
    power of two, the lengths the fused kernels take.  ``fast_len='reference'``
    (``fft_maker.set('cuda', fast_len='reference')``): frames are padded as the
    reference's numpy maker pads them (2-3-5-7-smooth lengths), for results on
    identical framing; such frames run through the general FFT plans
    (Bluestein) at several times the cost.
    """
    _FFTBase = CudaFFTBase

    def __init__(self, fast_len='pow2'):
        if fast_len not in ('pow2', 'reference'):
            raise ValueError("fast_len should be 'pow2' or 'reference'.")
        self._repr_kwargs = {} if fast_len == 'pow2' else {
            'fast_len': fast_len}
        if fast_len == 'reference':
            self.next_fast_len = smooth_fast_len

    def __call__(self, shape, dtype, direction='forward', axis=0, ortho=False,
                 sample_rate=None):
        n = tuple(shape)[axis]
        if n < 2:
            raise NotImplementedError("the cuda FFT maker needs at least two "
                                      "points along the transform axis.")
        return super().__call__(shape=shape, dtype=dtype, direction=direction,
                                axis=axis, ortho=ortho,
                                sample_rate=sample_rate)

    @staticmethod
    def next_fast_len(n):
        """Smallest power of two >= n (the lengths the kernels handle)."""
        n = int(n)
        return 1 if n <= 1 else 1 << (n - 1).bit_length()
