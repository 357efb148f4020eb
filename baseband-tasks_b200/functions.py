"""Detection: `Square` and `Power` on the GPU.

Mirrors baseband_tasks/functions.py of the reference: `Square` (:19-56)
squares real samples or takes |z|^2 of complex ones; `Power` (:59-143) turns
two polarizations X, Y into [|X|^2, |Y|^2, Re(X conj Y), Im(X conj Y)] along
the polarization axis, which is found from the polarization labels (:102).
"""
import numpy as np

from . import _buffers as B
from . import _cabi
from .base import TaskBase, simplify_shape

__all__ = ['Square', 'Power']


def _pair_labels(first, second):
    """Element-wise concatenation of two arrays of polarization labels."""
    return np.char.add(np.asarray(first, dtype=str),
                       np.asarray(second, dtype=str))


def _real_dtype(dtype):
    """The real dtype matching a (possibly complex) sample dtype."""
    dtype = np.dtype(dtype)
    return np.dtype(f'f{dtype.itemsize // 2}') if dtype.kind == 'c' else dtype


class _Detector(TaskBase):
    """What `Square` and `Power` share: labels for the repr and a float
    result, computed by one kernel over whatever block of frames is read."""
    _on_device = True
    _multi_frame = True
    _grid_shift = 0      # sample k of the output is sample k of the input

    def _labels_from(self, ih):
        raise NotImplementedError

    def _repr_item(self, key, default, value=None):
        # Labels derived from the input are the default, so they are not
        # repeated in the repr (functions.py:52-55,125-129).
        if key == 'polarization' and default is None \
                and hasattr(self.ih, 'polarization'):
            default = self._labels_from(self.ih)
        return super()._repr_item(key, default=default, value=value)


class Square(_Detector):
    """Intensities: x**2 for real samples, |z|**2 for complex ones.

    Parameters
    ----------
    ih : task or stream reader
        The stream to detect.
    polarization : array or nested list of str, optional
        Labels of the output.  Default: each input label doubled ('X' ->
        'XX'), or none if the input carries no labels.
    """

    def __init__(self, ih, polarization=None):
        self._complex = np.dtype(ih.dtype).kind == 'c'
        labels = polarization if polarization is not None \
            else self._labels_from(ih)
        super().__init__(ih, dtype=_real_dtype(ih.dtype), polarization=labels)

    def _labels_from(self, ih):
        pol = getattr(ih, 'polarization', None)
        return None if pol is None else _pair_labels(pol, pol)

    def task(self, data, out=None):
        host = not B.is_tensor(data)
        x = B.as_device(data, dtype=np.complex64 if self._complex
                        else np.float32)
        result = out
        if result is None or result.dtype != B.torch_dtype(np.float32):
            result = B.empty(x.shape, np.float32)
        lib = _cabi.lib()
        lib.check(lib.bbt_square_exec(B.ptr(x), B.ptr(result), x.numel(),
                                      int(self._complex),
                                      _cabi.stream_ptr()))
        return _finish(result, out, host, self.dtype)


class Power(_Detector):
    """Powers and cross terms of a pair of polarizations.

    From X and Y the four products ``|X|^2``, ``|Y|^2``, ``Re(X conj Y)`` and
    ``Im(X conj Y)`` are formed along the polarization axis, labelled 'XX',
    'YY', 'XY' and 'YX' (functions.py:59-143).

    Parameters
    ----------
    ih : task or stream reader
        Complex stream with two polarizations along one axis.
    polarization : array or nested list of str, optional
        The four output labels, shaped so that they identify the polarization
        axis.  Default: derived from the labels of ``ih`` as above; then
        ``ih`` has to have them (`AttributeError` otherwise).

    A `ValueError` is raised for real streams, for anything but two input
    polarizations, and for output labels that are not four distinct ones
    along a single axis.
    """

    def __init__(self, ih, polarization=None):
        if polarization is None:
            labels = self._labels_from(ih)
        else:
            labels = simplify_shape(np.asanyarray(polarization))
            distinct = len(np.unique(labels))
            if labels.size != 4 or distinct != 4 or 4 not in labels.shape:
                raise ValueError("need 4 distinct output polarizations along "
                                 "one axis.")
        # Labels are aligned with the trailing axes of a sample.
        axis = ih.ndim - labels.ndim + labels.shape.index(4)
        if ih.shape[axis] != 2:
            raise ValueError(f"axis {axis} of the input holds "
                             f"{ih.shape[axis]} polarizations, not 2.")
        if np.dtype(ih.dtype).kind != 'c':
            raise ValueError("cross products need a complex stream.")
        self._axis = axis
        before, after = ih.shape[:axis], ih.shape[axis + 1:]
        super().__init__(ih, shape=before + (4,) + after, polarization=labels,
                         dtype=_real_dtype(ih.dtype))
        # Kernel view: (A, 2, B) -> (A, 4, B) with A including time.
        self._inner = int(np.prod(after, dtype=np.int64))
        self._outer_per_sample = int(np.prod(before[1:], dtype=np.int64))
        # Straight after (de)dispersion the products are formed by its last
        # pass (bbt_dedisperse_power_exec) and the voltages never reach HBM.
        self._fused = (self._inner == 1 and hasattr(ih, 'can_detect')
                       and ih.can_detect())

    def _ih_read(self, start, count, device=None):
        if self._fused:
            return self.ih.read_detected(start, count)
        return super()._ih_read(start, count, device)

    def _run_frames(self, f0, f1, out=None):
        if not self._fused:
            return super()._run_frames(f0, f1, out)
        start = f0 * self._ih_samples_per_frame
        stop = min(f1 * self._ih_samples_per_frame, self._ih_stop)
        return self.ih.read_detected(start, stop - start, out=out)

    def _labels_from(self, ih):
        pol = ih.polarization        # AttributeError if there is none
        if pol.size != 2:
            raise ValueError("exactly 2 input polarizations are needed; "
                             "reshape the stream first.")
        return _pair_labels(pol[[0, 1, 0, 1]], pol[[0, 1, 1, 0]])

    def task(self, data, out=None):
        """The four products for a block of samples."""
        if (self._fused and B.is_tensor(data)
                and data.dtype == B.torch_dtype(np.float32)):
            # Detected already on the way here (_ih_read).
            if out is not None:
                out.copy_(data)
                return out
            return data
        host = not B.is_tensor(data)
        x = B.as_device(data, dtype=np.complex64)
        n = x.shape[0]
        result = out
        if result is None or result.dtype != B.torch_dtype(np.float32):
            result = B.empty((n,) + self.shape[1:], np.float32)
        lib = _cabi.lib()
        lib.check(lib.bbt_power_exec(B.ptr(x), B.ptr(result),
                                     n * self._outer_per_sample, self._inner,
                                     _cabi.stream_ptr()))
        return _finish(result, out, host, self.dtype)


def _finish(result, out, host, dtype):
    if out is not None:
        if result is not out:
            out.copy_(result)
        return out
    if result.dtype != B.torch_dtype(dtype):
        result = result.to(B.torch_dtype(dtype))
    return B.as_host(result) if host else result
