"""Coherent (de)dispersion on the GPU.

Mirrors `Disperse` / `Dedisperse` of the reference (dispersion.py:16-190):
same arguments, padding from the band-edge delays (:54-93), start-time shift
(:96), FFT length from the maker's ``next_fast_len`` (:97-101) and the same
per-frame result ``ifft(fft(x) * phase_factor)[pad_start:pad_start+spf]``
(:135-139).  The three numpy passes of the reference are one plan in the CUDA
library (include/bbt_b200.h, bbt_dedisperse_*): the chirp is generated on the
device in float64 and multiplied inside the middle FFT pass, whole runs of
overlap-save frames are processed per launch, and only the valid samples of
each frame are written.
"""
import ctypes

import numpy as np

from . import _buffers as B
from . import _cabi
from ._units import to_float, to_mhz
from .base import PaddedTaskBase, SetAttribute, getattr_if_none
from .dm import DispersionMeasure
from .fourier import fft_maker
from .fourier.cuda import CudaFFTMaker
from .sampling import ShiftSamples

__all__ = ['Disperse', 'Dedisperse', 'DisperseSamples', 'DedisperseSamples']


class Disperse(PaddedTaskBase):
    """Coherently disperse a time stream.

    Parameters
    ----------
    ih : task or stream reader
        Input data stream, with time as the first axis.
    dm : float or `~baseband_tasks_b200.dm.DispersionMeasure`
        Dispersion measure (pc/cm^3).  If negative, will dedisperse.
    reference_frequency : frequency, optional
        Frequency to which the data should be dispersed.  Can be an array.
        By default, the mean frequency.
    samples_per_frame : int, optional
        Number of dispersed samples which should be produced in one go.  The
        number of input samples used will be larger to avoid wrapping.  If
        not given, the minimum power-of-two frame that gives at least 75%
        efficiency.
    frequency, sideband : optional
        Frequencies and sidebands of the channels of ``ih``.  Default: taken
        from ``ih``.
    """
    _on_device = True
    _multi_frame = True

    def __init__(self, ih, dm, *, reference_frequency=None,
                 samples_per_frame=None, frequency=None, sideband=None):
        dm = DispersionMeasure(dm)
        frequency = getattr_if_none(ih, 'frequency', frequency)
        sideband = getattr_if_none(ih, 'sideband', sideband)
        sideband = np.where(np.asanyarray(sideband) > 0, 1, -1)
        # Work in MHz and microseconds, like the oracle.
        freq_mhz = to_mhz(frequency)
        rate_mhz = to_mhz(ih.sample_rate)
        half_rate = rate_mhz / 2.
        if ih.complex_data:
            freq_low = freq_mhz - half_rate
            freq_high = freq_mhz + half_rate
        else:
            freq_low = freq_mhz + np.minimum(sideband, 0.) * half_rate
            freq_high = freq_mhz + np.maximum(sideband, 0.) * half_rate

        if reference_frequency is None:
            fref_mhz = np.mean(freq_low + freq_high) / 2.
            reference_frequency = fref_mhz * 1e6 * _unit_of(frequency)
        else:
            fref_mhz = to_mhz(reference_frequency)

        delay_low = _time_delay(dm, freq_low, fref_mhz)
        delay_high = _time_delay(dm, freq_high, fref_mhz)
        delay_max = max(np.max(delay_low), np.max(delay_high))
        delay_min = min(np.min(delay_low), np.min(delay_high))
        rate_hz = rate_mhz * 1e6
        pad_start = int(np.ceil(delay_max * rate_hz))
        pad_end = int(np.ceil(-delay_min * rate_hz))
        if pad_start < 0:
            # Both delays less than 0: shift, and pad less at the end.
            assert pad_end > 0
            sample_offset = pad_start
            pad_end += pad_start
            pad_start = 0
        elif pad_end < 0:
            sample_offset = -pad_end
            pad_start += pad_end
            pad_end = 0
        else:
            sample_offset = 0

        FFT = fft_maker.get()
        if not isinstance(FFT, CudaFFTMaker):
            raise TypeError("coherent dedispersion on the GPU needs the "
                            "'cuda' FFT maker (fft_maker.set('cuda')).")
        start_time = ih.start_time + sample_offset / ih.sample_rate
        super().__init__(ih, pad_start=pad_start, pad_end=pad_end,
                         samples_per_frame=samples_per_frame,
                         next_fast_len=FFT.next_fast_len,
                         frequency=frequency, sideband=sideband,
                         start_time=start_time)
        # Real-valued streams (rfft/irfft in the reference) run through the
        # complex kernels with a Hermitian response.
        self._real = not self.ih.complex_data
        self._FFT = FFT
        self._dm = dm
        self.reference_frequency = reference_frequency
        self._sample_offset = sample_offset
        self._grid_shift = self._pad_start + sample_offset
        self._pad_slice = slice(self._pad_start,
                                self._pad_start + self.samples_per_frame)
        self._rate_mhz = rate_mhz
        # One chirp per distinct (frequency, reference, sideband).
        shape = self.ih.sample_shape
        f = np.broadcast_to(freq_mhz, shape).ravel()
        r = np.broadcast_to(fref_mhz, shape).ravel()
        s = np.broadcast_to(sideband, shape).ravel()
        keys = list(zip(f.tolist(), r.tolist(), s.tolist()))
        uniq = sorted(set(keys))
        index = {k: i for i, k in enumerate(uniq)}
        self._series_map = np.array([index[k] for k in keys], np.int32)
        self._chirp_par = (np.array([u[0] for u in uniq], np.float64),
                           np.array([u[1] for u in uniq], np.float64),
                           np.array([u[2] for u in uniq], np.int8))
        self._n_series = len(keys)
        self._plan = None
        self._work = None
        # Frame lengths the fused plan does not take (not a power of two: the
        # reference's own framing, CudaFFTMaker(fast_len='reference')) go
        # frame by frame through the FFT plans: fft, x phase factor, ifft.
        n = self._ih_samples_per_frame
        self._generic = bool(n & (n - 1))
        self._generic_parts = None

    # ------------------------------------------------- Power fused in
    # While set, frames come out of the last pass as the four products of
    # each polarization pair (float32, in the bytes of the two voltages):
    # `functions.Power` reads its input this way when `can_detect` says the
    # plan has that path, and the voltages never go to memory.
    _detect = False

    def can_detect(self):
        """Whether Power can be fused into the last dedispersion pass: the
        polarization pair has to be the last axis of a sample."""
        if (self._generic or self._real or len(self.sample_shape) == 0
                or self.sample_shape[-1] != 2
                or np.dtype(self.dtype) != np.complex64):
            return False
        try:
            lib = _cabi.lib()
            return bool(lib.bbt_dedisperse_power_supported(self._get_plan()))
        except _cabi.BBTError:
            # No plan for this frame length (or no device yet): the tasks
            # run one after the other and report the problem when read.
            return False

    def read_detected(self, start, count, out=None):
        """Samples [start, start+count) with Power applied, as a device
        tensor of shape (count,) + sample_shape[:-1] + (4,), float32
        (written straight into ``out`` if that is given)."""
        single = B.torch_dtype(np.float32)
        target = None
        if (out is not None and out.is_contiguous() and out.dtype == single):
            target = out.view(B.torch_dtype(np.complex64))
        self._detect = True
        self._frame_index = None        # cached frames hold voltages
        try:
            self.seek(start)
            count = self._check_read(count, None)
            data = B.as_device(self._read_data(count, target))
        finally:
            self._detect = False
            self._frame_index = None
        if target is not None:
            return out
        data = data.contiguous().view(single).reshape(
            (count,) + self.sample_shape[:-1] + (4,))
        if out is not None:
            out.copy_(data)
            return out
        return data

    # ------------------------------------------------------------- the plan
    def _get_plan(self):
        if self._plan is None:
            lib = _cabi.lib()
            f, r, s = self._chirp_par
            plan = ctypes.c_void_p()
            dbl = ctypes.POINTER(ctypes.c_double)
            lib.check(lib.bbt_dedisperse_plan_create(
                ctypes.byref(plan), self._ih_samples_per_frame,
                self._n_series, self._pad_start, self.samples_per_frame,
                len(f),
                self._series_map.ctypes.data_as(
                    ctypes.POINTER(ctypes.c_int32)),
                None if self._real else f.ctypes.data_as(dbl),
                None if self._real else r.ctypes.data_as(dbl),
                None if self._real else s.ctypes.data_as(
                    ctypes.POINTER(ctypes.c_int8)),
                float(self._dm), float(self._rate_mhz),
                float(self._sample_offset), 0))
            if self._real:
                response = self._hermitian_response()
                lib.check(lib.bbt_dedisperse_plan_set_response(
                    plan, response.ctypes.data_as(ctypes.c_void_p)))
            self._plan = plan
        return self._plan

    def _real_phase_factor(self):
        """Phase factors at the rfft frequencies of each distinct chirp,
        float64 phases rounded to complex64 (dispersion.py:115-129)."""
        n = self._ih_samples_per_frame
        fftfreq = np.fft.rfftfreq(n) * self._rate_mhz
        f, r, s = self._chirp_par
        d = DispersionMeasure.dispersion_delay_constant * float(self._dm)
        out = np.empty((len(f), n // 2 + 1), np.complex64)
        for c in range(len(f)):
            freq = f[c] + fftfreq * s[c]
            phase = d * freq * (1. / r[c] - 1. / freq)**2 * 1e6 * s[c]
            if self._sample_offset != 0:
                phase = phase + (self._sample_offset / self._rate_mhz
                                 * fftfreq)
            out[c] = np.exp(phase * (2j * np.pi))
        return out

    def _hermitian_response(self):
        """Full-length response reproducing rfft -> x phase factor -> irfft:
        Hermitian extension, with the imaginary parts that irfft ignores
        (bins 0 and N/2) dropped."""
        half = self._real_phase_factor()
        n = self._ih_samples_per_frame
        full = np.empty((half.shape[0], n), np.complex64)
        full[:, :n // 2 + 1] = half
        full[:, n // 2 + 1:] = half[:, n // 2 - 1:0:-1].conj()
        full[:, 0] = half[:, 0].real
        full[:, n // 2] = half[:, n // 2].real
        return np.ascontiguousarray(full)

    @property
    def _fft(self):
        """Forward transform of one frame (dispersion.py:105-107)."""
        return self._FFT(shape=(self._ih_samples_per_frame,)
                         + self.ih.sample_shape, dtype=self.ih.dtype,
                         sample_rate=self.ih.sample_rate)

    @property
    def _ifft(self):
        return self._fft.inverse()

    @property
    def phase_factor(self):
        """Phase factors of the Fourier-transformed frame, as used by the
        kernels: float64 phases rounded to complex64 (dispersion.py:115-129),
        shape ``(N,) + sample_shape``."""
        lib = _cabi.lib()
        n = self._ih_samples_per_frame
        if self._generic:
            return self._host_phase_factor()
        if self._real:
            return self._real_phase_factor()[self._series_map].T.reshape(
                (n // 2 + 1,) + self.ih.sample_shape)
        host = np.empty((len(self._chirp_par[0]), n), np.complex64)
        lib.check(lib.bbt_dedisperse_plan_get_response(
            self._get_plan(), host.ctypes.data_as(ctypes.c_void_p)))
        return host[self._series_map].T.reshape(
            (n,) + self.ih.sample_shape)

    @property
    def dm(self):
        return self._dm

    # ------------------------------------------------------------ the work
    def task_frames(self, data, n_frames, out=None):
        """(De)disperse ``n_frames`` overlapping frames of a run of input.

        ``data`` holds ``(n_frames-1)*samples_per_frame + N`` samples; frame i
        starts at ``i*samples_per_frame``.  Returns (or fills ``out`` with)
        the ``n_frames*samples_per_frame`` valid output samples.
        """
        host = not B.is_tensor(data)
        if self._generic:
            return self._task_frames_generic(data, n_frames, out, host)
        lib = _cabi.lib()
        plan = self._get_plan()
        if self._real:
            return self._task_frames_real(data, n_frames, out, host)
        result = self._exec_frames(B.as_device(data, dtype=np.complex64),
                                   n_frames, out)
        if out is None and host:
            return B.as_host(result)
        return result

    def _exec_frames(self, x, n_frames, out):
        lib = _cabi.lib()
        plan = self._get_plan()
        S, spf = self._n_series, self.samples_per_frame
        N = self._ih_samples_per_frame
        assert x.shape[0] == (n_frames - 1) * spf + N
        # The kernels write complex64: a caller's buffer of another dtype
        # (a complex128 stream) or layout is filled through a temporary.
        direct = (out is not None and out.is_contiguous()
                  and out.dtype == B.torch_dtype(np.complex64))
        result = out if direct else B.empty(
            (n_frames * spf,) + self.sample_shape, np.complex64)
        assert result.shape[0] == n_frames * spf
        wb = lib.bbt_dedisperse_work_bytes(plan, n_frames)
        if self._work is None or self._work.numel() < wb:
            self._work = None
            self._work = B.empty((max(wb, 16),), np.uint8)
        run = (lib.bbt_dedisperse_power_exec if self._detect
               else lib.bbt_dedisperse_exec)
        lib.check(run(plan, B.ptr(x), spf * S, n_frames, 0, B.ptr(result),
                      spf * S, B.ptr(self._work), _cabi.stream_ptr()))
        if out is not None and not direct:
            out.copy_(result)
            return out
        return result

    def _task_frames_generic(self, data, n_frames, out, host):
        """Any frame length: what `Disperse.task` of the reference does
        (dispersion.py:135-139), with the transforms on the GPU."""
        lib = _cabi.lib()
        single = np.float32 if self._real else np.complex64
        x = B.as_device(data, dtype=single)
        if self._generic_parts is None:
            fft = self._fft
            factor = B.as_device(np.ascontiguousarray(
                np.broadcast_to(self._host_phase_factor(),
                                fft.frequency_shape), dtype=np.complex64))
            self._generic_parts = (fft, fft.inverse(), factor)
        fft, ifft, factor = self._generic_parts
        N, spf = self._ih_samples_per_frame, self.samples_per_frame
        direct = (out is not None and out.is_contiguous()
                  and out.dtype == B.torch_dtype(single))
        result = out if direct else B.empty(
            (n_frames * spf,) + self.sample_shape, single)
        for f in range(n_frames):
            ft = fft(x[f * spf:f * spf + N])
            lib.check(lib.bbt_multiply_exec(B.ptr(ft), B.ptr(factor),
                                            B.ptr(ft), ft.numel(),
                                            _cabi.stream_ptr()))
            back = ifft(ft)
            result[f * spf:(f + 1) * spf] = back[self._pad_slice]
        if out is not None:
            if not direct:
                out.copy_(result)
            return out
        if self.dtype != np.dtype(single):
            result = result.to(B.torch_dtype(self.dtype))
        return B.as_host(result) if host else result

    def _host_phase_factor(self):
        """`phase_factor` computed on the host in float64 and rounded to
        complex64 (dispersion.py:115-129), shape (N or N//2+1,) + sample_shape
        -- for the frame lengths that do not go through the fused plan."""
        n = self._ih_samples_per_frame
        if self._real:
            half = self._real_phase_factor()
            return half[self._series_map].T.reshape(
                (n // 2 + 1,) + self.ih.sample_shape)
        fftfreq = np.fft.fftfreq(n) * self._rate_mhz
        f, r, s = self._chirp_par
        d = DispersionMeasure.dispersion_delay_constant * float(self._dm)
        table = np.empty((len(f), n), np.complex64)
        for c in range(len(f)):
            freq = f[c] + fftfreq * s[c]
            phase = d * freq * (1. / r[c] - 1. / freq)**2 * 1e6 * s[c]
            if self._sample_offset != 0:
                phase = phase + (self._sample_offset / self._rate_mhz
                                 * fftfreq)
            table[c] = np.exp(phase * (2j * np.pi))
        return table[self._series_map].T.reshape((n,) + self.ih.sample_shape)

    def _task_frames_real(self, data, n_frames, out, host):
        """Real-valued streams: frames 2p and 2p+1 go through the complex
        kernels as the real and imaginary part of one frame (the Hermitian
        response is a real convolution, which treats the two separately)."""
        lib = _cabi.lib()
        plan = self._get_plan()
        x = B.as_device(data, dtype=np.float32)
        S, spf = self._n_series, self.samples_per_frame
        N = self._ih_samples_per_frame
        assert x.shape[0] == (n_frames - 1) * spf + N
        n_pairs = (n_frames + 1) // 2
        z = B.empty((n_pairs * N,) + self.sample_shape, np.complex64)
        lib.check(lib.bbt_pair_frames_exec(B.ptr(x), B.ptr(z), x.shape[0],
                                           spf, N, S, n_frames,
                                           _cabi.stream_ptr()))
        w = B.empty((n_pairs * spf,) + self.sample_shape, np.complex64)
        wb = lib.bbt_dedisperse_work_bytes(plan, n_pairs)
        if self._work is None or self._work.numel() < wb:
            self._work = None
            self._work = B.empty((max(wb, 16),), np.uint8)
        lib.check(lib.bbt_dedisperse_exec(
            plan, B.ptr(z), N * S, n_pairs, 0, B.ptr(w), spf * S,
            B.ptr(self._work), _cabi.stream_ptr()))
        del z
        single = B.torch_dtype(np.float32)
        direct = (out is not None and out.is_contiguous()
                  and out.dtype == single)
        result = out if direct else B.empty(
            (n_frames * spf,) + self.sample_shape, np.float32)
        lib.check(lib.bbt_unpair_frames_exec(B.ptr(w), B.ptr(result), spf, S,
                                             n_frames, _cabi.stream_ptr()))
        if out is not None:
            if not direct:
                out.copy_(result)
            return out
        if self.dtype != np.dtype(np.float32):
            result = result.to(B.torch_dtype(self.dtype))
        return B.as_host(result) if host else result

    def task(self, data, out=None):
        """One frame: N input samples to ``samples_per_frame`` outputs."""
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


class Dedisperse(Disperse):
    """Coherently dedisperse a time stream (dispersion.py:149-190).

    Parameters are as for `Disperse`; the dispersion measure is removed
    rather than added.
    """

    def __init__(self, ih, dm, *, reference_frequency=None,
                 samples_per_frame=None, frequency=None, sideband=None):
        super().__init__(ih, -DispersionMeasure(dm),
                         reference_frequency=reference_frequency,
                         samples_per_frame=samples_per_frame,
                         frequency=frequency, sideband=sideband)

    @property
    def dm(self):
        return -self._dm


class DisperseSamples(ShiftSamples):
    """Incoherently shift a time stream to give it a dispersive time delay
    (dispersion.py:193-251): only whole-sample shifts by the delay at the
    mid-channel frequency, no in-channel smearing.

    Parameters are as for `Disperse`.
    """

    def __init__(self, ih, dm, *, reference_frequency=None,
                 samples_per_frame=None, frequency=None, sideband=None):
        if frequency is not None or sideband is not None:
            ih = SetAttribute(ih, frequency=frequency, sideband=sideband)
        frequency = ih.frequency
        freq_mhz = to_mhz(frequency)
        if not ih.complex_data:
            # Mid-channel frequency for real data.
            freq_mhz = freq_mhz + ih.sideband * to_mhz(ih.sample_rate) / 2.
        if reference_frequency is None:
            fref_mhz = np.mean(freq_mhz)
            reference_frequency = fref_mhz * 1e6 * _unit_of(frequency)
        else:
            fref_mhz = to_mhz(reference_frequency)
        dm = DispersionMeasure(dm)
        time_delay = _time_delay(dm, freq_mhz, fref_mhz)
        shift = time_delay * (to_mhz(ih.sample_rate) * 1e6)
        super().__init__(ih, shift, samples_per_frame=samples_per_frame)
        self.reference_frequency = reference_frequency
        self._dm = dm

    @property
    def dm(self):
        return self._dm


class DedisperseSamples(DisperseSamples):
    """Incoherently shift a time stream to correct for a dispersive time
    delay (dispersion.py:254-298)."""

    def __init__(self, ih, dm, *, reference_frequency=None,
                 samples_per_frame=None, frequency=None, sideband=None):
        super().__init__(ih, -DispersionMeasure(dm),
                         reference_frequency=reference_frequency,
                         samples_per_frame=samples_per_frame,
                         frequency=frequency, sideband=sideband)

    @property
    def dm(self):
        return -self._dm


def _time_delay(dm, f_mhz, fref_mhz):
    d = DispersionMeasure.dispersion_delay_constant * float(dm)
    return d * (1. / np.asarray(f_mhz)**2 - 1. / np.asarray(fref_mhz)**2)


def _unit_of(quantity):
    """1 Hz in the kind of object ``quantity`` is (Quantity or number)."""
    unit = getattr(quantity, 'unit', None)
    if unit is None:
        return 1.
    import astropy.units as u
    return u.Hz
