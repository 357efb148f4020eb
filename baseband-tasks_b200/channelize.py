"""Channelizer and dechannelizer on the GPU.

Mirrors `Channelize` / `Dechannelize` of the reference (channelize.py:13-178):
blocks of ``n`` time samples are Fourier transformed along axis 1 of
``(samples_per_frame, n) + ih.sample_shape``; the output sample shape is
``(n,) + ih.sample_shape`` (``n // 2 + 1`` channels for real input), the rate
``ih.sample_rate / n`` and the channel frequencies
``frequency + fftfreq * sideband`` in FFT order (:57-71).
"""
import operator

import numpy as np

from . import _buffers as B
from .base import TaskBase, getattr_if_none
from .fourier import fft_maker

__all__ = ['Channelize', 'Dechannelize']


class Channelize(TaskBase):
    """Basic channelizer.

    Parameters
    ----------
    ih : task or stream reader
        Input data stream, with time as the first axis.
    n : int
        Number of input samples to channelize.
    samples_per_frame : int, optional
        Number of complete output samples per frame.  Default: 1.  (Reads of
        many frames are transformed in one launch whatever this is.)
    frequency, sideband : optional
        Frequencies and sidebands of the channels of ``ih``.  Default: taken
        from ``ih`` (if available).
    """
    _on_device = True
    _multi_frame = True

    def __init__(self, ih, n, samples_per_frame=1, *,
                 frequency=None, sideband=None):
        self._n = n = operator.index(n)
        samples_per_frame = operator.index(samples_per_frame)
        # The check TaskBase makes (base.py:681-683), before any plan exists.
        assert ih.shape[0] >= n * samples_per_frame, (
            "not enough samples to fill one frame of spectra.")
        self._FFT = fft_maker.get()
        self._fft = self._FFT((samples_per_frame, n) + tuple(ih.sample_shape),
                              ih.dtype, axis=1, sample_rate=ih.sample_rate)
        self._ffts = {samples_per_frame: self._fft}

        # Channel frequencies in FFT order, on either side of each input
        # frequency according to its sideband (channelize.py:60-64).
        f_in = getattr_if_none(ih, 'frequency', frequency, required=False)
        sb_in = getattr_if_none(ih, 'sideband', sideband, required=False)
        f_out = None if f_in is None else f_in + self._fft.frequency * sb_in
        super().__init__(ih, shape=(-1,) + self._fft.frequency_shape[1:],
                         sample_rate=ih.sample_rate / n,
                         samples_per_frame=samples_per_frame,
                         dtype=self._fft.frequency_dtype,
                         frequency=f_out, sideband=sb_in)

    def _fft_for(self, n_spec):
        fft = self._ffts.get(n_spec)
        if fft is None:
            fft = self._FFT((n_spec, self._n) + tuple(self.ih.sample_shape),
                            self.ih.dtype, axis=1,
                            sample_rate=self.ih.sample_rate)
            if len(self._ffts) > 8:
                self._ffts = {self.samples_per_frame: self._fft}
            self._ffts[n_spec] = fft
        return fft

    def task(self, data, out=None):
        n_spec = data.shape[0] // self._n
        fft = self._fft_for(n_spec)
        result = fft(data.reshape(fft.time_shape), out=out)
        if out is not None and result is not out:
            out.copy_(result)
            return out
        return result

    def inverse(self, ih):
        """Create a Dechannelize instance that undoes this Channelization."""
        with fft_maker.set(self._FFT):
            return Dechannelize(ih, n=self._fft.time_shape[1],
                                dtype=self._fft.time_dtype)

    def close(self):
        super().close()
        self._ffts = {}
        self._fft = None


class Dechannelize(TaskBase):
    """Basic dechannelizer (channelize.py:90-178).

    Inverse Fourier transform on first sample axis (which gets removed).
    """
    _on_device = True
    _multi_frame = True

    def __init__(self, ih, n=None, samples_per_frame=None, *,
                 dtype=None, frequency=None, sideband=None):
        assert ih.complex_data, "Dechannelization needs complex spectra."
        dtype = np.dtype(ih.dtype if dtype is None else dtype)
        if n is not None:
            n = operator.index(n)
        elif dtype.kind == 'c':
            n = ih.sample_shape[0]      # complex: as many samples as channels
        else:
            raise ValueError("a real-valued output needs the number of "
                             "samples per spectrum, 'n'.")
        # Frames hold whole spectra: at least one.
        ih_samples_per_frame = (ih.samples_per_frame
                                if samples_per_frame is None
                                else max(int(round(samples_per_frame / n)), 1))

        self._FFT = fft_maker.get()
        self._ifft = self._FFT((ih_samples_per_frame, n)
                               + tuple(ih.sample_shape[1:]),
                               dtype=dtype, axis=1, direction='backward')
        self._iffts = {ih_samples_per_frame: self._ifft}
        self._n = n
        if frequency is None and getattr(ih, 'frequency', None) is not None:
            frequency = ih.frequency[0]      # the zero-frequency channel
        super().__init__(ih, ih_samples_per_frame=ih_samples_per_frame,
                         shape=(-1,) + tuple(ih.shape[2:]),
                         sample_rate=ih.sample_rate * n,
                         dtype=self._ifft.time_dtype,
                         frequency=frequency, sideband=sideband)

    def task(self, data, out=None):
        n_spec = data.shape[0]
        ifft = self._iffts.get(n_spec)
        if ifft is None:
            ifft = self._FFT((n_spec, self._n) + tuple(self.ih.sample_shape[1:]),
                             dtype=self.dtype, axis=1, direction='backward')
            self._iffts = {self._ih_samples_per_frame: self._ifft,
                           n_spec: ifft}
        result = ifft(data, out=out)
        if out is not None:
            if result is not out:
                out.copy_(result.reshape(out.shape))
            return out
        return result.reshape((-1,) + self.sample_shape)

    def inverse(self, ih):
        """Create a Channelize instance that undoes this Dechannelization."""
        with fft_maker.set(self._FFT):
            return Channelize(ih, n=self._ifft.time_shape[1])
