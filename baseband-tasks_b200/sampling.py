"""Integer sample shifts on the GPU.

Mirrors `ShiftSamples` of the reference (sampling.py:380-425): channels are
shifted by whole numbers of samples (positive shifts delay a channel); the
stream is padded by the spread of the shifts and starts ``shift.max()``
samples later.  The resampling tasks of the reference's sampling module are
outside the accelerated path.
"""
import numpy as np

from . import _buffers as B
from . import _cabi
from ._units import to_float
from .base import PaddedTaskBase, check_broadcast_to

__all__ = ['ShiftSamples', 'to_sample']


def to_sample(ih, offset):
    """The offset in units of samples: numbers are samples already,
    quantities with units of time are multiplied by the sample rate."""
    unit = getattr(offset, 'unit', None)
    if unit is not None and getattr(unit, 'physical_type', '') == 'time':
        return to_float(offset * ih.sample_rate)
    return np.asarray(to_float(offset) if unit is not None else offset,
                      dtype=float)


class ShiftSamples(PaddedTaskBase):
    """Shift channels in a stream by integer numbers of samples.

    Parameters
    ----------
    ih : task or stream reader
        Input data stream, with time as the first axis.
    shift : int or float array-like, or time quantity
        Amount by which to shift samples along the stream (rounded to the
        nearest integer).  Should broadcast to the sample shape.
    samples_per_frame : int
        Number of shifted samples which should be produced in one go.
    """
    _on_device = True
    _multi_frame = True

    def __init__(self, ih, shift, *, samples_per_frame=None):
        shift = self._shift = np.round(to_sample(ih, shift)).astype(int)
        check_broadcast_to(shift, ih.sample_shape)
        start_time = ih.start_time + int(shift.max()) / ih.sample_rate
        super().__init__(ih, pad_start=0, pad_end=int(np.ptp(shift)),
                         samples_per_frame=samples_per_frame,
                         start_time=start_time)
        offsets = np.broadcast_to(shift.max() - shift, self.sample_shape)
        self._offsets = np.ascontiguousarray(offsets, dtype=np.int64).ravel()
        self._d_offsets = None
        if np.dtype(self.dtype).itemsize not in (4, 8):
            raise NotImplementedError("samples should be float32 or "
                                      "complex64.")

    def task_frames(self, data, n_frames, out=None):
        lib = _cabi.lib()
        host = not B.is_tensor(data)
        x = B.as_device(data)
        if self._d_offsets is None:
            self._d_offsets = B.as_device(self._offsets)
        n_out = n_frames * self.samples_per_frame
        result = out
        if result is None:
            result = B.empty((n_out,) + self.sample_shape, self.dtype)
        lib.check(lib.bbt_shift_exec(
            B.ptr(x), B.ptr(result), B.ptr(self._d_offsets), n_out,
            max(len(self._offsets), 1), np.dtype(self.dtype).itemsize,
            _cabi.stream_ptr()))
        return B.as_host(result) if (host and out is None) else result

    def task(self, data, out=None):
        return self.task_frames(data, 1, out=out)

    def close(self):
        super().close()
        self._d_offsets = None
