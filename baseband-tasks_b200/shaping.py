"""Slicing of streams: what ``stream[item]`` returns.

Minimal counterpart of `GetSlice` of the reference (shaping.py, used by
`Base.__getitem__`, base.py:469-472): a time slice (``start:stop`` in samples,
step 1) optionally followed by an index into the samples.  The other shaping
tasks of the reference are view/metadata operations outside the accelerated
path and are not provided.
"""
import operator

import numpy as np

from . import _buffers as B
from .base import TaskBase

__all__ = ['GetSlice']


class GetSlice(TaskBase):
    """Stream restricted to a slice in time (and an item of each sample)."""

    def __init__(self, ih, item):
        if not isinstance(item, tuple):
            item = (item,)
        time_item, sample_item = item[0], item[1:]
        if isinstance(time_item, slice):
            start, stop, step = time_item.indices(ih.shape[0])
            if step != 1:
                raise ValueError("time slices should have a step of 1.")
        else:
            start = operator.index(time_item)
            if start < 0:
                start += ih.shape[0]
            stop = start + 1
        if not 0 <= start < stop <= ih.shape[0]:
            raise IndexError("slice is empty or out of range.")
        self._start, self._stop = start, stop
        self._grid_shift = start
        self._sample_item = (slice(None),) + tuple(sample_item)
        probe = np.empty((1,) + tuple(ih.sample_shape), np.int8)
        sample_shape = probe[self._sample_item].shape[1:]
        meta = {}
        for attr in ('frequency', 'sideband', 'polarization'):
            value = getattr(ih, attr, None)
            if value is not None and sample_item:
                full = np.broadcast_to(value, ih.sample_shape, subok=True)
                value = full[tuple(sample_item)]
            if value is not None:
                meta[attr] = value
        spf = min(ih.samples_per_frame, stop - start)
        super().__init__(ih, ih_samples_per_frame=spf, samples_per_frame=spf,
                         shape=(stop - start,) + sample_shape,
                         start_time=ih._tell_time(start),
                         **meta)
        self._on_device = hasattr(ih, 'read_device')
        self._time_from_ih = getattr(ih, '_time_from_ih', False)

    def _tell_time(self, offset):
        # Through the underlying stream (shaping.py:412-413): its samples need
        # not be evenly spaced in time (phase-stepped Integrate, PulseStack).
        return self.ih._tell_time(self._start + offset)

    def _read_data(self, count, out=None):
        data = self._ih_read(self._start + self.offset, count)
        self.offset += count
        data = data[self._sample_item]
        if out is not None:
            out[...] = data
            return out
        if B.is_tensor(data):
            return data.contiguous()
        return data
