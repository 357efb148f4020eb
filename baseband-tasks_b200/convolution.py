"""Convolution with an arbitrary response, in the Fourier domain, on the GPU.

Mirrors `Convolve` of the reference (convolution.py:65-127): overlap-save
frames, ``ifft(fft(x) * fft(response zero-padded to the frame))`` with the
first ``len(response) - 1`` samples of every frame discarded.  It reuses the
three-pass dedispersion kernels with the transformed response in place of the
dispersion chirp (`bbt_dedisperse_plan_set_response`), for complex and for
real-valued streams (the transform of a real response is Hermitian).
"""
import ctypes

import numpy as np

from . import _buffers as B
from . import _cabi
from .base import PaddedTaskBase, check_broadcast_to
from .fourier import fft_maker
from .fourier.cuda import CudaFFTMaker

__all__ = ['Convolve']


def adjust_response_dims(response, ih):
    """One-dimensional responses apply along time (convolution.py:13-19)."""
    response = np.asanyarray(response)
    if response.ndim == 1 and ih.ndim > 1:
        response = response.reshape(response.shape[:1]
                                    + (1,) * (ih.ndim - 1))
    else:
        check_broadcast_to(response, response.shape[:1]
                           + tuple(ih.sample_shape))
    return response


class Convolve(PaddedTaskBase):
    """Convolve a time stream with a response, in the Fourier domain.

    Parameters
    ----------
    ih : task or stream reader
        Input data stream, with time as the first axis.
    response : `~numpy.ndarray`
        Response to convolve the time stream with.  If one-dimensional,
        assumed to apply to the sample axis of ``ih``.
    offset : int, optional
        Where samples should be considered to be taken from.  For the default
        of 0, a given sample has the same time as the convolution of the
        filter with all preceding samples.
    samples_per_frame : int, optional
        Number of output samples which should be produced in each frame.
        Default: the smallest power-of-two frame with at least 75% efficiency.
    """
    _on_device = True
    _multi_frame = True

    def __init__(self, ih, response, *, offset=0, samples_per_frame=None):
        self._response = adjust_response_dims(response, ih)
        pad = self._response.shape[0] - 1
        FFT = fft_maker.get()
        if not isinstance(FFT, CudaFFTMaker):
            raise TypeError("convolution on the GPU needs the 'cuda' FFT "
                            "maker (fft_maker.set('cuda')).")
        super().__init__(ih, pad_start=pad - offset, pad_end=offset,
                         samples_per_frame=samples_per_frame,
                         next_fast_len=FFT.next_fast_len)
        self._FFT = FFT
        self._real = not self.ih.complex_data
        if self._real and np.iscomplexobj(self._response):
            raise ValueError("a real stream needs a real response.")
        self._n_series = int(np.prod(self.ih.sample_shape, dtype=np.int64))
        self._plan = None
        self._work = None

    @property
    def _ft_response(self):
        """Transform of the zero-padded response, one row per distinct
        series response, and the map from series to row."""
        n = self._ih_samples_per_frame
        resp = np.broadcast_to(self._response, self._response.shape[:1]
                               + tuple(self.ih.sample_shape))
        resp = resp.reshape(resp.shape[0], -1).T        # [series][taps]
        uniq, index = np.unique(resp, axis=0, return_inverse=True)
        long_response = np.zeros((uniq.shape[0], n), np.complex64)
        long_response[:, :uniq.shape[1]] = uniq
        ft = np.fft.fft(long_response, axis=1).astype(np.complex64)
        return np.ascontiguousarray(ft), index.astype(np.int32).ravel()

    def _get_plan(self):
        if self._plan is None:
            lib = _cabi.lib()
            ft, series_map = self._ft_response
            plan = ctypes.c_void_p()
            pad = self._pad_start + self._pad_end
            lib.check(lib.bbt_dedisperse_plan_create(
                ctypes.byref(plan), self._ih_samples_per_frame,
                self._n_series, pad, self.samples_per_frame, ft.shape[0],
                series_map.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                None, None, None, 0., 1., 0., 0))
            lib.check(lib.bbt_dedisperse_plan_set_response(
                plan, ft.ctypes.data_as(ctypes.c_void_p)))
            self._plan = plan
        return self._plan

    def task_frames(self, data, n_frames, out=None):
        lib = _cabi.lib()
        plan = self._get_plan()
        host = not B.is_tensor(data)
        S, spf = self._n_series, self.samples_per_frame
        if self._real:
            xr = B.as_device(data, dtype=np.float32)
            x = B.empty(xr.shape, np.complex64)
            lib.check(lib.bbt_convert_exec(B.ptr(xr), B.ptr(x), xr.numel(), 0,
                                           _cabi.stream_ptr()))
        else:
            x = B.as_device(data, dtype=np.complex64)
        direct = (out is not None and not self._real
                  and out.dtype == B.torch_dtype(np.complex64))
        y = out if direct else B.empty((n_frames * spf,) + self.sample_shape,
                                       np.complex64)
        wb = lib.bbt_dedisperse_work_bytes(plan, n_frames)
        if self._work is None or self._work.numel() < wb:
            self._work = None
            self._work = B.empty((max(wb, 16),), np.uint8)
        lib.check(lib.bbt_dedisperse_exec(
            plan, B.ptr(x), spf * S, n_frames, 0, B.ptr(y), spf * S,
            B.ptr(self._work), _cabi.stream_ptr()))
        if self._real:
            yr = B.empty(y.shape, np.float32)
            lib.check(lib.bbt_convert_exec(B.ptr(y), B.ptr(yr), y.numel(), 1,
                                           _cabi.stream_ptr()))
            y = yr
        if out is not None:
            if y is not out:
                out.copy_(y)
            return out
        if np.dtype(self.dtype) not in (np.dtype('c8'), np.dtype('f4')):
            y = y.to(B.torch_dtype(self.dtype))
        return B.as_host(y) if host else y

    def task(self, data, out=None):
        return self.task_frames(data, 1, out=out)

    def close(self):
        super().close()
        plan, self._plan = self._plan, None
        self._work = None
        if plan is not None:
            _cabi.lib().bbt_dedisperse_plan_destroy(plan)

    def __del__(self):
        plan, self._plan = getattr(self, '_plan', None), None
        if plan is not None:
            try:
                _cabi.lib().bbt_dedisperse_plan_destroy(plan)
            except Exception:
                pass
