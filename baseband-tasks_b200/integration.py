"""Integration over time and folding on pulse phase, on the GPU.

Mirrors `Integrate` and `Fold` of the reference (integration.py:52-395): same
arguments and bin bookkeeping -- the bin edges in the upstream stream are
``around(k / mean_offset_size + ih_start)`` (:184-186) computed on the host in
float64 and handed to the kernels as one int64 table, so counts and bin
assignment are bit-exact; empty bins average to NaN (:268-269);
``average=False`` gives a structured array with ``data`` and ``count``.

The segmented sums (:290-303) and the phase-bin scatter-add (:380-395) are CUDA
kernels (bbt_integrate_exec, bbt_fold_exec).  When the input is
``Power(Channelize(x))`` (or ``Power(x)`` for `Fold`) with the polarization
axis last, the fused kernels read ``x`` directly, so detected spectra are
never written to HBM.
"""
import ctypes
import warnings

import numpy as np

from . import _buffers as B
from . import _cabi
from . import base as _base
from ._units import is_index, to_float
from .base import BaseTaskBase
from .channelize import Channelize
from .functions import Power

__all__ = ['Integrate', 'Fold', 'PulseStack', 'PolynomialPhase']

_MAX_BINS_PER_LAUNCH = 32768
# Fewest output values for which averages are formed inside the kernels.
_AVERAGE_IN_KERNEL_MIN = 262144


class PolynomialPhase:
    """Pulse phase as a polynomial in time, evaluable on the device.

    ``phase(t) = sum_k coef[k] * dt**k`` cycles, with ``dt`` the time in
    seconds since ``reference_time``.  As a ``phase`` callable it takes times
    (like any other callable passed to `Fold`); `Fold` recognises it and
    evaluates it inside the fold kernel from the sample index instead, as
    ``dt = (float64(i) - i_ref) / sample_rate`` with Horner's rule and
    individually rounded float64 operations -- `of_index` is the same
    arithmetic in numpy (the documented time convention for bit-exact bins).
    Here ``i`` counts samples on the grid of the whole observation
    (``Base._sample_grid``: tasks that cut or pad a stream pass the grid of
    their input on with an exact integer shift), and ``i_ref`` is the index on
    that grid at which ``dt`` is zero; a block of a stream shared out in time
    therefore gets the very same bins as the whole stream would.
    """

    def __init__(self, coef, reference_time):
        self.coef = np.atleast_1d(np.asarray(coef, dtype=np.float64))
        if not 1 <= len(self.coef) <= 8:
            raise ValueError("need between 1 and 8 coefficients.")
        self.reference_time = reference_time

    def __call__(self, time):
        dt = np.asarray(to_float(time - self.reference_time),
                        dtype=np.float64)
        return self._horner(dt)

    def _horner(self, dt):
        phase = np.full(np.shape(dt), self.coef[-1])
        for c in self.coef[-2::-1]:
            phase = phase * dt + c
        return phase

    def i_ref(self, start_time, sample_rate):
        """Index (fractional) at which ``dt`` is zero on a grid of samples
        that starts at ``start_time``."""
        return float(to_float((self.reference_time - start_time)
                              * sample_rate))

    def grid(self, stream):
        """``(i_ref, i_0)`` for folding ``stream``: the index of the reference
        time and of the stream's first sample on the stream's sample grid."""
        time, index = stream._sample_grid()
        return self.i_ref(time, stream.sample_rate), index

    def of_index(self, index, i_ref, sample_rate):
        rate = float(to_float(sample_rate * 1.))
        dt = (np.asarray(index).astype(np.float64) - i_ref) / rate
        return self._horner(dt)


class Integrate(BaseTaskBase):
    """Integrate a stream stepwise.

    Parameters
    ----------
    ih : task or stream reader
        Input data stream, with time as the first axis.
    step : int or time interval, optional
        Interval over which to integrate: a number of samples of the
        underlying stream, a time (seconds or a quantity), or, with ``phase``,
        a phase interval.  Default: all samples.
    phase : callable, optional
        Should return full pulse phase (including cycle count) for the times
        passed in; integration is then over intervals of ``step`` in phase.
    start : time or int, optional
        Time or offset at which to start the integration.  Default: 0.
    average : bool, optional
        Whether to return the average (default) or, in a structured array
        with ``'data'`` and ``'count'``, the sums and numbers of samples.
    samples_per_frame : int, optional
        Number of output samples per frame.  Default: 1.  (Reads of many
        samples are integrated in one pass whatever this is.)
    dtype : `~numpy.dtype`, optional
        Output dtype.
    """
    _on_device = True

    def __init__(self, ih, step=None, phase=None, *,
                 start=0, average=True, samples_per_frame=1, dtype=None):
        self._start, self._step = start, step
        # Where in ``ih`` the integration starts: a whole sample the pointer
        # can be put at, plus -- when ``start`` is a time between samples --
        # the fraction the bin edges are to be counted from.
        first = ih.seek(start)
        n_in = ih.shape[0] - first
        if first < 0 or n_in < 0:
            raise ValueError("'start' is not within the underlying stream.")
        if is_index(start):
            edge0, t_start = first, ih.time
        else:
            edge0 = first + to_float((start - ih.time) * ih.sample_rate)
            t_start = start
        # The output grid: how many samples, at what rate, starting where.
        if step is None:
            step = n_in
        if is_index(step):
            assert phase is None, 'cannot pass in phase and integer step'
            rate_out, n_out = ih.sample_rate / step, n_in / step
            origin = t_start
        else:
            # A time interval, or with ``phase`` an interval in phase.
            at = (lambda t: t) if phase is None else phase
            origin, end = at(t_start), at(ih.stop_time)
            rate_out = 1 / step
            n_out = to_float((end - origin) * rate_out)
        # Output samples per input sample (not necessarily 1 / integer), and
        # the number of output samples: a last one counts if at least half
        # of it is there.
        self._mean_offset_size = n_out / n_in
        self._sample_start = origin
        n_out = int(n_out + self._mean_offset_size / 2)
        assert n_out >= 1, "time per frame larger than total time in stream"
        # Streams that count in cycles (``phase`` given, or the input already
        # does: a phase-stepped Integrate, a PulseStack) have no start time
        # of their own; times come from the input (the reference tells the
        # two apart by the unit of the rate, integration.py:146-151).
        self._time_from_ih = (phase is not None
                              or getattr(ih, '_time_from_ih', False))
        if dtype is None:
            dtype = ih.dtype if average else np.dtype(
                [('data', ih.dtype), ('count', int)])
        super().__init__(ih, shape=(n_out,) + tuple(ih.sample_shape),
                         sample_rate=rate_out,
                         samples_per_frame=samples_per_frame,
                         start_time=False if self._time_from_ih else origin,
                         dtype=dtype)
        self.average = average
        self._phase = phase
        self._ih_start = edge0
        self._setup_source()

    # ------------------------------------------------------------- sources
    def _setup_source(self):
        """Decide what the kernels read: ``ih`` itself or, fused, the input
        of ``Power(Channelize(x))``."""
        ih = self.ih
        self._fused = None
        self._src = ih
        self._src_ratio = 1      # source samples per ih sample
        if (type(ih) is Power and type(getattr(ih, 'ih', None)) is Channelize
                and np.dtype(ih.ih.ih.dtype) == np.complex64
                and ih._axis == ih.ndim - 1 and ih.ih.ih.ndim >= 2
                and _fusable_channelizer(ih.ih._n)):
            ch = ih.ih
            self._fused = 'chanpow'
            self._src = ch.ih
            self._src_ratio = ch._n
            self._chan_n = ch._n
            self._chan_m = int(np.prod(ch.ih.sample_shape[:-1],
                                       dtype=np.int64))

    def _tell_time(self, offset):
        if self._start_time is not False:
            return super()._tell_time(offset)
        return self.ih._tell_time(self._get_offsets(offset))

    def _get_offsets(self, samples, precision=1.e-3, max_iter=10):
        """Sample numbers of ``ih`` at which the output samples ``samples``
        begin (bin edges; fractional input allowed), as integers.

        Without a phase the edges are evenly spaced:
        ``around(samples / mean_offset_size + start)`` -- this expression is
        the contract for bit-exact counts (integration.py:184-186).  With a
        phase, the sample at which the phase reaches each edge is found by
        inverting ``phase(t)`` numerically, to ``precision`` samples.
        """
        if self._phase is None:
            return (np.around(np.asanyarray(samples) / self._mean_offset_size
                              + self._ih_start).astype(int))
        ih = self.ih
        shape = np.shape(samples)
        # Phase, from the start of the integration, each edge should have.
        want = np.asarray(to_float(np.ravel(samples) / self.sample_rate),
                          dtype=float)
        per_sample = float(to_float(self._mean_offset_size
                                    / self.sample_rate))  # mean phase step
        n_rel = ih.shape[0] - self._ih_start

        def phase_at(rel):
            t = ih.start_time + (rel + self._ih_start) / ih.sample_rate
            return np.asarray(to_float(self._phase(t) - self._sample_start),
                              dtype=float)

        # A secant search per edge, all edges at once: start from where the
        # mean phase step puts the edge, then follow the local slope.
        rel = np.clip(want / per_sample, 0., n_rel)
        got = phase_at(rel)
        slope = np.full(rel.shape, per_sample)
        todo = np.ones(rel.shape, bool)
        for _ in range(max_iter):
            step = (want[todo] - got[todo]) / slope[todo]
            new = np.clip(rel[todo] + step, 0., n_rel)
            new_got = phase_at(new)
            moved = new - rel[todo]
            with np.errstate(divide='ignore', invalid='ignore'):
                local = (new_got - got[todo]) / moved
            usable = np.isfinite(local) & (local > 0)
            slope[todo] = np.where(usable, local, slope[todo])
            rel[todo], got[todo] = new, new_got
            todo[todo] = np.abs(moved) > precision
            if not todo.any():
                break
        else:  # pragma: no cover
            warnings.warn('offset calculation did not converge. '
                          'This should not happen!')
        return ((rel + self._ih_start).round().astype(int).reshape(shape))

    # -------------------------------------------------------------- reading
    def _read_frame(self, frame_index):
        sample0 = frame_index * self.samples_per_frame
        n_sample = min(self.samples_per_frame, self.shape[0] - sample0)
        return self._integrate_samples(sample0, n_sample)

    def _read_data(self, count, out=None):
        # Many output samples in one pass over the upstream data (results
        # per bin do not depend on how output samples are grouped in frames).
        if count == 0 or self._frame_dependent():
            return super()._read_data(count, out)
        per_sample = max(1, 4 * int(np.prod(self.sample_shape,
                                            dtype=np.int64)))
        max_n = max(1, _base.BLOCK_BYTES // per_sample)
        a = self.offset
        if count <= max_n and out is None:
            result = self._integrate_samples(a, count)
        else:
            result = out
            pos = 0
            while pos < count:
                n = min(max_n, count - pos)
                part = self._integrate_samples(a + pos, n)
                if result is None:
                    result = _base._empty_like(
                        part, (count,) + tuple(part.shape[1:]))
                result[pos:pos + n] = part
                pos += n
        self.offset = a + count
        return result

    def _frame_dependent(self):
        return False

    def read_sums(self, count=None, within=None):
        """Sums and counts of the next ``count`` output samples, on the device.

        Returns ``(sums, counts)``: float32 sums with the shape of the output
        samples (complex streams as real, imaginary pairs in the last axis)
        and int64 counts per bin.  This is what is reduced over ranks when a
        stream is sharded in time (`baseband_tasks_b200.parallel`): with
        ``within=(first, last)`` only the samples ``[first, last)`` of the
        underlying stream are read, so a rank holding a block of the stream
        gets partial sums and counts for the bins its block cuts through.
        """
        count = self._check_read(count, None)
        a = self.offset
        saved, self._raw_sums = getattr(self, '_raw_sums', False), True
        self._within = within
        try:
            if self._frame_dependent():
                parts = []
                spf = self.samples_per_frame
                pos = a
                while pos < a + count:
                    n = min(a + count - pos, spf - pos % spf)
                    parts.append(self._integrate_samples(pos, n))
                    pos += n
                sums = B.torch().cat([p[0] for p in parts])
                cnts = B.torch().cat([p[1] for p in parts])
            else:
                sums, cnts = self._integrate_samples(a, count)
        finally:
            self._raw_sums = saved
            self._within = None
        self.offset = a + count
        return sums, cnts

    _within = None

    def _clip(self, start, stop):
        """Upstream range to read, restricted to ``within`` of `read_sums`."""
        if self._within is not None:
            start = max(start, int(self._within[0]))
            stop = min(stop, int(self._within[1]))
        return start, stop

    def _integrate_samples(self, sample0, n_sample):
        """Output samples [sample0, sample0 + n_sample)."""
        samples = np.arange(sample0, sample0 + n_sample + 1)
        offsets = self._get_offsets(samples).astype(np.int64)
        inner = int(np.prod(self.ih.sample_shape, dtype=np.int64))
        if self.ih.complex_data:
            inner *= 2
        sums = B.zeros((n_sample, inner), np.float32)
        count = B.zeros((n_sample,), np.int64)
        # Averages are formed inside the kernels (each contribution divided
        # by the width of its bin, known from the offsets).
        # Only with enough bins to fill the GPU at one partial sum per bin;
        # otherwise bins are split over many CTAs and divided afterwards.
        parallel = (n_sample * self._chan_m * 128 if self._fused == 'chanpow'
                    else n_sample * inner)
        self._average_in_kernel = (self.average
                                   and parallel >= _AVERAGE_IN_KERNEL_MIN
                                   and not getattr(self, '_raw_sums', False))
        self._accumulate(offsets, sums, count)
        if self._average_in_kernel:
            empty = np.flatnonzero(np.diff(offsets) == 0)
            if len(empty):      # no samples: NaN, as 0 / 0 gives
                sums[B.as_device(empty)] = float('nan')
        return self._finish(sums, count, (n_sample,) + self.sample_shape)

    def _finish(self, sums, count, shape):
        """Average on the device, or assemble the structured host array."""
        if getattr(self, '_raw_sums', False):
            return sums, count
        lib = _cabi.lib()
        ih_dtype = np.dtype(self.ih.dtype)
        single = np.complex64 if ih_dtype.kind == 'c' else np.float32
        if self.average:
            if getattr(self, '_average_in_kernel', False):
                out = sums
            else:
                out = B.empty(sums.shape, np.float32)
                lib.check(lib.bbt_average_exec(
                    B.ptr(sums), B.ptr(count), B.ptr(out), count.numel(),
                    sums.numel() // max(count.numel(), 1),
                    _cabi.stream_ptr()))
            out = _as_dtype(out, single).reshape(shape)
            if self.dtype != np.dtype(single):
                out = out.to(B.torch_dtype(self.dtype))
            return out
        data = B.as_host(_as_dtype(sums, single)).reshape(shape)
        frame = np.zeros(shape, dtype=self.dtype)
        frame['data'] = data
        cnt = B.as_host(count)
        frame['count'] = cnt.reshape(cnt.shape + (1,) * (
            len(shape) - cnt.ndim))
        return frame

    def _accumulate(self, offsets, sums, count):
        """Add upstream samples [offsets[0], offsets[-1]) to their bins."""
        lib = _cabi.lib()
        d_off = B.as_device(offsets)
        n_bins = len(offsets) - 1
        start, stop = self._clip(int(offsets[0]), int(offsets[-1]))
        ratio = self._src_ratio
        src = self._src
        per_src_sample = (int(np.prod(src.sample_shape, dtype=np.int64))
                          * np.dtype(src.dtype).itemsize)
        # Chunks are whole runs of frames of the source (so that it computes
        # every frame once): boundaries at multiples of ``chunk_src`` source
        # samples, rounded down to whole samples of ``ih``.
        src_frame = max(1, getattr(src, 'samples_per_frame', 1))
        chunk_src = src_frame * max(1, _base.BLOCK_BYTES
                                    // max(per_src_sample * src_frame, 1))
        chunk_src = max(chunk_src, ratio)
        pos = start
        while pos < stop:
            nxt = ((pos * ratio) // chunk_src + 1) * chunk_src // ratio
            nxt = min(stop, max(nxt, pos + 1))
            n = nxt - pos
            src.seek(pos * ratio)
            if hasattr(src, 'read_device'):
                x = src.read_device(n * ratio)
            else:
                x = B.as_device(src.read(n * ratio))
            # Bins with any overlap with [pos, nxt).
            b0 = int(np.searchsorted(offsets[1:], pos, side='right'))
            b1 = int(np.searchsorted(offsets[:-1], nxt, side='left'))
            for bb in range(b0, b1, _MAX_BINS_PER_LAUNCH):
                nb = min(_MAX_BINS_PER_LAUNCH, b1 - bb)
                if self._fused == 'chanpow':
                    lib.check(lib.bbt_channelize_power_integrate_exec(
                        B.ptr(x), self._chan_n, self._chan_m, n, pos,
                        B.ptr(d_off), bb, nb, B.ptr(sums), B.ptr(count),
                        int(self._average_in_kernel), _cabi.stream_ptr()))
                else:
                    x = _as_float32(x, src.dtype)
                    lib.check(lib.bbt_integrate_exec(
                        B.ptr(x), n, sums.shape[1], pos, B.ptr(d_off), bb,
                        nb, B.ptr(sums), B.ptr(count),
                        int(self._average_in_kernel), _cabi.stream_ptr()))
            pos = nxt
        assert n_bins == count.shape[0]


def _fusable_channelizer(n):
    """Lengths the fused Channelize -> Power -> Integrate kernel takes (one
    spectrum per block FFT); longer ones are read unfused through `ih`."""
    return 2 <= n <= 16384 and n & (n - 1) == 0


def _as_dtype(t, dtype):
    """View float32 pairs as complex64 where the stream is complex."""
    if np.dtype(dtype).kind == 'c':
        return B.torch().view_as_complex(t.reshape(t.shape[:-1] + (-1, 2)))
    return t


def _as_float32(x, dtype):
    dtype = np.dtype(dtype)
    t = B.torch()
    if dtype.kind == 'c':
        if dtype != np.complex64:
            x = x.to(t.complex64)
        return t.view_as_real(x)
    if dtype != np.float32:
        x = x.to(t.float32)
    return x


class Fold(Integrate):
    """Fold pulse profiles in fixed time intervals.

    Parameters
    ----------
    ih : task or stream reader
        Input data stream, with time as the first axis.
    n_phase : int
        Number of bins per pulse period.
    phase : callable
        Should return pulse phases (with or without cycle count, in cycles)
        for the times passed in.  A `PolynomialPhase` is evaluated inside the
        fold kernel; any other callable is evaluated on the host, per sample,
        like the reference does.
    step : int or time interval, optional
        Number of input samples or time interval over which to fold.
        Default: the whole stream, in a single profile.
    start, average, samples_per_frame, dtype
        As for `Integrate` (``dtype`` is ignored, as in the reference).
    """

    def __init__(self, ih, n_phase, phase, step=None, *,
                 start=0, average=True, samples_per_frame=1, dtype=None):
        super().__init__(ih, step=step, start=start, average=average,
                         samples_per_frame=samples_per_frame)
        self._shape = (self._shape[0], n_phase) + tuple(ih.sample_shape)
        self.n_phase = n_phase
        self.phase = phase

    def _setup_source(self):
        ih = self.ih
        self._fused = None
        self._src = ih
        self._src_ratio = 1
        if (type(ih) is Power and np.dtype(ih.ih.dtype) == np.complex64
                and ih._axis == ih.ndim - 1):
            self._fused = 'power'
            self._src = ih.ih

    def _frame_dependent(self):
        # With several time bins per frame the reference assigns a sample
        # exactly on an inner bin edge to the earlier bin (searchsorted
        # side='left', integration.py:386): results depend on the framing.
        return self.samples_per_frame > 1

    def _integrate_samples(self, sample0, n_sample):
        samples = np.arange(sample0, sample0 + n_sample + 1)
        offsets = self._get_offsets(samples).astype(np.int64)
        lo = offsets[:-1].copy()
        hi = offsets[1:].copy()
        if self.samples_per_frame > 1 and n_sample > 1:
            lo[1:] += 1
            hi[:-1] += 1
        inner = int(np.prod(self.ih.sample_shape, dtype=np.int64))
        if self.ih.complex_data:
            inner *= 2
        sums = B.zeros((n_sample, self.n_phase, inner), np.float32)
        count = B.zeros((n_sample, self.n_phase), np.int64)
        self._fold(offsets, lo, hi, inner, sums, count)
        return self._finish(sums, count,
                            (n_sample, self.n_phase) + self.sample_shape[1:])

    def _fold(self, offsets, lo, hi, inner, sums, count):
        lib = _cabi.lib()
        d_lo, d_hi = B.as_device(lo), B.as_device(hi)
        start, stop = self._clip(int(offsets[0]), int(offsets[-1]))
        src = self._src
        ih = self.ih
        per_sample = (int(np.prod(src.sample_shape, dtype=np.int64))
                      * np.dtype(src.dtype).itemsize)
        chunk = max(1, _base.BLOCK_BYTES // max(per_sample, 1))
        src_spf = max(1, getattr(src, 'samples_per_frame', 1))
        if chunk > src_spf:
            chunk = (chunk // src_spf) * src_spf
        poly = self.phase if isinstance(self.phase, PolynomialPhase) else None
        if poly is not None:
            rate = float(to_float(ih.sample_rate * 1.))
            i_ref, i_grid = poly.grid(ih)
            coef = poly.coef
            coef_p = coef.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        pos = start
        while pos < stop:
            nxt = min(stop, (pos // chunk + 1) * chunk)
            n = nxt - pos
            src.seek(pos)
            if hasattr(src, 'read_device'):
                x = src.read_device(n)
            else:
                x = B.as_device(src.read(n))
            if self._fused is None:
                x = _as_float32(x, src.dtype)
            if poly is None:
                # Times from the absolute sample index (the reference adds
                # offsets within the frame to the frame's start time,
                # integration.py:375-388): the phase bin of a sample then
                # does not depend on where frames or `start` happen to fall.
                raw_items = np.arange(pos, nxt)
                phases = self.phase(ih.start_time
                                    + raw_items / ih.sample_rate)
                phases = np.asarray(_cycles(phases), dtype=np.float64)
                pbin = ((phases % 1.) * self.n_phase).astype(np.int32)
                d_pbin = B.as_device(pbin)
            b0 = int(np.searchsorted(hi, pos, side='right'))
            b1 = int(np.searchsorted(lo, nxt, side='left'))
            for bb in range(b0, b1, _MAX_BINS_PER_LAUNCH):
                nb = min(_MAX_BINS_PER_LAUNCH, b1 - bb)
                if poly is None:
                    lib.check(lib.bbt_fold_exec(
                        B.ptr(x), int(self._fused == 'power'), n, inner, pos,
                        pos, B.ptr(d_lo), B.ptr(d_hi), bb, nb, B.ptr(d_pbin),
                        None, 0, 0., 1., self.n_phase, B.ptr(sums),
                        B.ptr(count), _cabi.stream_ptr()))
                else:
                    lib.check(lib.bbt_fold_exec(
                        B.ptr(x), int(self._fused == 'power'), n, inner, pos,
                        pos + i_grid, B.ptr(d_lo), B.ptr(d_hi), bb, nb, None,
                        coef_p,
                        len(coef), i_ref, rate, self.n_phase, B.ptr(sums),
                        B.ptr(count), _cabi.stream_ptr()))
            pos = nxt


def _cycles(phases):
    """Phases as plain numbers of cycles (quantities are converted)."""
    if hasattr(phases, 'to_value'):
        return phases.to_value('cycle')
    return phases


class PulseStack(BaseTaskBase):
    """Create a stream of pulse profiles (integration.py:398-477).

    One output sample is one pulse: ``n_phase`` bins in phase, each the
    integral of the input over 1 / n_phase of a cycle of ``phase`` (which has
    to include the cycle count).  Underneath is an `Integrate` in steps of
    ``1 / n_phase`` cycles; this class only groups its samples by pulse.

    Parameters are as for `Fold`, without ``step``.
    """

    def __init__(self, ih, n_phase, phase, *,
                 start=0, average=True, samples_per_frame=1, dtype=None):
        self.n_phase = n_phase
        bins = Integrate(ih, 1. / n_phase, phase, start=start,
                         average=average, dtype=dtype,
                         samples_per_frame=samples_per_frame * n_phase)
        n_pulse = bins.shape[0] // n_phase      # whole pulses only
        super().__init__(bins, shape=(n_pulse, n_phase) + bins.shape[1:],
                         sample_rate=bins.sample_rate / n_phase,
                         samples_per_frame=samples_per_frame, dtype=dtype)
        self._time_from_ih = True    # one sample per cycle, not per second

    def _read_frame(self, frame_index):
        # A frame here is the same frame of the binned stream, cut into
        # pulses; a last, shorter frame may end with part of a pulse.
        bins = self.ih._read_frame(frame_index)
        whole = len(bins) - len(bins) % self.n_phase
        return bins[:whole].reshape((-1,) + self.sample_shape)

    def _tell_time(self, offset):
        return self.ih._tell_time(offset * self.n_phase)
