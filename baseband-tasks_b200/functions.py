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


def _char_add(a, b):
    return np.char.add(np.asarray(a, dtype=str), np.asarray(b, dtype=str))


class Square(TaskBase):
    """Converts samples to intensities by squaring.

    Parameters
    ----------
    ih : task or stream reader
        Input data stream.
    polarization : array or (nested) list of char, optional
        Output polarization labels.  By default, doubled labels from the
        underlying stream (and ignored if not given).
    """
    _on_device = True
    _multi_frame = True

    def __init__(self, ih, polarization=None):
        if polarization is None:
            polarization = self._default_polarization(ih)
        ih_dtype = np.dtype(ih.dtype)
        self._complex = ih_dtype.kind == 'c'
        dtype = np.zeros(1, dtype=ih_dtype).real.dtype
        super().__init__(ih, dtype=dtype, polarization=polarization)

    def _default_polarization(self, ih):
        if not hasattr(ih, 'polarization'):
            return None
        return _char_add(ih.polarization, ih.polarization)

    def _repr_item(self, key, default, value=None):
        if key == 'polarization':
            default = self._default_polarization(self.ih)
        return super()._repr_item(key, default=default, value=value)

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


class Power(TaskBase):
    """Calculate powers and cross terms for two polarizations.

    For polarizations X and Y, 4 terms are produced: ``XX = |X|^2``,
    ``YY = |Y|^2``, ``XY = Re(X conj Y)`` and ``YX = Im(X conj Y)``.

    Parameters
    ----------
    ih : task or stream reader
        Input data stream.
    polarization : array or (nested) list of char, optional
        Output polarization labels.  By default, inferred from the
        underlying stream, using the scheme described above.

    Raises
    ------
    AttributeError
        If no polarization information is given.
    ValueError
        If the underlying stream is not complex, the number of polarizations
        not equal to two, or the polarization labels not unique.
    """
    _on_device = True
    _multi_frame = True

    def __init__(self, ih, polarization=None):
        if polarization is None:
            polarization = self._default_polarization(ih)
        else:
            polarization = simplify_shape(np.asanyarray(polarization))
            if not (polarization.size == 4 == len(np.unique(polarization))
                    and 4 in polarization.shape):
                raise ValueError('output polarizations should have 4 unique '
                                 'elements along one axis.')

        self._axis = ih.ndim - polarization.ndim + polarization.shape.index(4)
        if ih.shape[self._axis] != 2:
            raise ValueError(f"input shape should be 2 along polarization axis"
                             f" ({self._axis}), not {ih.shape[self._axis]}.")

        shape = ih.shape[:self._axis] + (4,) + ih.shape[self._axis + 1:]
        ih_dtype = np.dtype(ih.dtype)
        if ih_dtype.kind != 'c':
            raise ValueError("Power only works on a complex timestream.")
        dtype = np.zeros(1, ih_dtype).real.dtype
        super().__init__(ih, shape=shape, polarization=polarization,
                         dtype=dtype)
        # Kernel view: (A, 2, B) -> (A, 4, B) with A including time.
        self._inner = int(np.prod(ih.shape[self._axis + 1:], dtype=np.int64))
        self._outer_per_sample = int(np.prod(ih.shape[1:self._axis],
                                             dtype=np.int64))

    def _default_polarization(self, ih):
        if ih.polarization.size != 2:
            raise ValueError("stream should have exactly 2 polarizations. "
                             "Reshape appropriately.")
        return _char_add(ih.polarization[[0, 1, 0, 1]],
                         ih.polarization[[0, 1, 1, 0]])

    def _repr_item(self, key, default, value=None):
        if (key == 'polarization' and hasattr(self.ih, 'polarization')
                and default is None):
            default = self._default_polarization(self.ih)
        return super()._repr_item(key, default=default, value=value)

    def task(self, data, out=None):
        """Calculate the polarization powers and cross terms."""
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
