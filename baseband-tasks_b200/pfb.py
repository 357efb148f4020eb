"""Polyphase filter bank on the GPU: FIR and FFT in one kernel.

Mirrors `sinc_hamming`, `PolyphaseFilterBankSamples` and `PolyphaseFilterBank`
of the reference (pfb.py:14-154): spectrum ``j`` is the Fourier transform of
``sum_t response[t] * x[(j + t) * n : (j + t + 1) * n]``; the stream is padded by
``(n_tap - 1) * n`` samples, half at each side, so that time stamps refer to
the centre of the filter (:74-89).  The reference's two classes differ only in
how the FIR is evaluated (time domain or Fourier domain along the block axis,
equal to rounding); both map to the same fused kernel here.
"""
import operator

import numpy as np

from . import _buffers as B
from . import _cabi
from .base import TaskBase, getattr_if_none
from .fourier import fft_maker

__all__ = ['sinc_hamming', 'PolyphaseFilterBankSamples',
           'PolyphaseFilterBank']


def sinc_hamming(n_tap, n_sample, sinc_scale=1.):
    r"""Construct a sinc-hamming polyphase filter (pfb.py:14-45).

    ``sinc(n_tap * sinc_scale * (k/N - 0.5)) * hamming(N)`` for
    ``N = n_tap * n_sample``, reshaped to ``(n_tap, n_sample)``; for example
    ``sinc_hamming(4, 2048)`` (CHIME) or ``sinc_hamming(12, 64, 0.95)`` (GUPPI).
    """
    n = n_tap * n_sample
    x = n_tap * sinc_scale * np.linspace(-0.5, 0.5, n, endpoint=False)
    return (np.sinc(x) * np.hamming(n)).reshape(n_tap, n_sample)


class PolyphaseFilterBankSamples(TaskBase):
    """Channelize using a polyphase filter bank.

    Parameters
    ----------
    ih : task or stream reader
        Input data stream, with time as the first axis.
    response : `~numpy.ndarray`
        Polyphase filter.  The first dimension is taken to be the
        number of taps, and the second the number of channels.
    samples_per_frame : int, optional
        Number of complete output samples per frame.  Default: inferred from
        padding, ensuring an efficiency of at least 75%.
    frequency, sideband : optional
        Frequencies and sidebands of the channels of ``ih``.  Default: taken
        from ``ih`` (if available).
    """
    _on_device = True
    _multi_frame = True

    def __init__(self, ih, response, samples_per_frame=None,
                 frequency=None, sideband=None):
        response = np.asarray(response)
        n_tap, n = response.shape
        n = operator.index(n)
        pad = (n_tap - 1) * n
        assert pad % 2 == 0
        # Framing of the reference's padded task (pfb.py:76-83,
        # base.py:750-768) and of the Channelize on top of it (:86-87).
        if samples_per_frame is None:
            padded_ih_spf = max(ih.samples_per_frame, pad * 4)
        else:
            padded_ih_spf = samples_per_frame * n + pad
        padded_spf = padded_ih_spf - pad
        spf = padded_spf // n
        if spf < 1:
            raise ValueError("frames should hold at least one spectrum.")
        n_padded = ih.shape[0] - pad
        self._n = n
        self._n_tap = n_tap
        self._response = response
        self._d_response = None
        # 0 complex, 1 float, 2 raw 8-bit integer samples (kept as int8 on
        # the device: a quarter of the bytes to copy and read).
        kind = np.dtype(ih.dtype).kind
        self._real = 2 if np.dtype(ih.dtype) == np.int8 else int(kind != 'c')
        fft_dtype = np.float32 if self._real == 2 else ih.dtype
        self._FFT = fft_maker.get()
        self._fft = self._FFT((spf, n) + tuple(ih.sample_shape), fft_dtype,
                              axis=1, sample_rate=ih.sample_rate)

        frequency = getattr_if_none(ih, 'frequency', frequency, required=False)
        sideband = getattr_if_none(ih, 'sideband', sideband, required=False)
        if frequency is not None:
            frequency = frequency + self._fft.frequency * sideband

        n_chan = self._fft.frequency_shape[1]
        shape = ((n_padded // (n * spf)) * spf, n_chan) + tuple(ih.sample_shape)
        start_time = ih.start_time + (pad // 2) / ih.sample_rate
        super().__init__(ih, shape=shape, sample_rate=ih.sample_rate / n,
                         samples_per_frame=spf,
                         ih_samples_per_frame=spf * n + pad,
                         start_time=start_time, frequency=frequency,
                         sideband=sideband, dtype=self._fft.frequency_dtype)
        self._inner = int(np.prod(ih.sample_shape, dtype=np.int64))

    def _run_frames(self, f0, f1, out=None):
        spf, n = self.samples_per_frame, self._n
        n_spec = (f1 - f0) * spf
        data = self._ih_read(f0 * spf * n, (n_spec + self._n_tap - 1) * n)
        return self.ppf_fft(data, out=out)

    def _read_frame(self, frame_index):
        return self._run_frames(frame_index, frame_index + 1)

    def ppf_fft(self, data, out=None):
        """Filter and transform all complete spectra in ``data``."""
        lib = _cabi.lib()
        host = not B.is_tensor(data)
        x = B.as_device(data, dtype=(np.complex64, np.float32,
                                     np.int8)[self._real])
        if self._d_response is None:
            self._d_response = B.as_device(
                np.ascontiguousarray(self._response, dtype=np.float32))
        n = self._n
        n_spec = x.shape[0] // n - (self._n_tap - 1)
        result = out
        if result is None or result.dtype != B.torch_dtype(np.complex64):
            result = B.empty((n_spec,) + self.sample_shape, np.complex64)
        lib.check(lib.bbt_pfb_exec(
            B.ptr(x), B.ptr(result), B.ptr(self._d_response), n, self._n_tap,
            self._inner, n_spec, int(self._real), _cabi.stream_ptr()))
        if out is not None:
            if result is not out:
                out.copy_(result)
            return out
        if self.dtype != np.dtype(np.complex64):
            result = result.to(B.torch_dtype(self.dtype))
        return B.as_host(result) if host else result

    task = ppf_fft

    def close(self):
        super().close()
        self._d_response = None


class PolyphaseFilterBank(PolyphaseFilterBankSamples):
    """Channelize using a polyphase filter bank.

    In the reference this class applies the filter in the Fourier domain
    (pfb.py:103-154), which is equal to rounding to the time-domain
    definition evaluated by the fused kernel used here.
    """
