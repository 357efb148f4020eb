"""CPU oracle for the baseband-tasks dedispersion / channelization hot path.

TEST INFRASTRUCTURE ONLY.  This module is a numpy-only *restatement* of the
arithmetic and index bookkeeping of mhvk/baseband-tasks (reference paths are
relative to /root/reference/baseband_tasks).  It is imported only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``; the product package
(``baseband_tasks_b200``) never imports it and has no CPU fallback.

Parity status: the reference itself cannot be imported in the build container
(astropy and baseband are absent), so this restatement is pinned against the
known-answer assertions of the reference's own tests (see
``tests/test_oracle_golden.py``: test_dm.py:33-73, test_dispersion.py:64-124,
test_pfb.py:26-102, test_integration.py:112-303, test_generators.py:253-316,
fourier/tests/test_fourier.py:77-166).  The FFT arithmetic is numpy.fft
(pocketfft), exactly what the reference's NumpyFFTMaker calls
(fourier/numpy.py:33-49).  Integer/time arithmetic at benchmark scale is
"parity unpinned" by the reference; it is pinned by sharing one host-computed
offsets table between oracle and kernels.

Conventions (the reference uses astropy units; here plain floats):
frequencies in MHz where a function says ``_mhz``, otherwise Hz; times in
seconds relative to the stream start; dispersion measure in pc/cm^3;
phases in cycles.
"""
import numpy as np

# dm.py:37 -- hardcoded tempo constant, s MHz^2 cm^3 / pc.
DISPERSION_DELAY_CONSTANT = 1. / 2.41e-4


# --------------------------------------------------------------------------
# fourier/numpy.py:99-126  (smallest 2-3-5-7 smooth number >= n)
def next_fast_len(n):
    n = int(n)
    if n <= 7:
        return n
    best = 2 * n
    p2 = 1
    while p2 < best:
        p3 = p2
        while p3 < best:
            p5 = p3
            while p5 < best:
                p7 = p5
                while p7 < n:
                    p7 *= 7
                if p7 < best:
                    best = p7
                p5 *= 5
            p3 *= 3
        p2 *= 2
    return best


def next_pow2(n):
    """What the 'cuda' maker uses as next_fast_len (dispersion.py:99 hook)."""
    n = int(n)
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


# --------------------------------------------------------------------------
# dm.py:42-120
def dm_time_delay(dm, f_mhz, fref_mhz=None):
    """Dispersion delay in seconds (dm.py:74-76)."""
    d = DISPERSION_DELAY_CONSTANT * dm
    ref_inv2 = 0. if fref_mhz is None else 1. / np.asarray(fref_mhz) ** 2
    return d * (1. / np.asarray(f_mhz) ** 2 - ref_inv2)


def dm_phase_delay(dm, f_mhz, fref_mhz=None):
    """Dispersion phase delay in cycles (dm.py:103-105).

    s MHz^2 * MHz * MHz^-2 = s MHz = 1e6 cycles.
    """
    d = DISPERSION_DELAY_CONSTANT * dm
    f_mhz = np.asarray(f_mhz)
    ref_inv = 0. if fref_mhz is None else 1. / np.asarray(fref_mhz)
    return d * f_mhz * (ref_inv - 1. / f_mhz) ** 2 * 1e6


def dm_phase_factor(dm, f_mhz, fref_mhz=None):
    """exp(i phase_delay) (dm.py:107-120)."""
    return np.exp(2j * np.pi * dm_phase_delay(dm, f_mhz, fref_mhz))


# --------------------------------------------------------------------------
# fourier/base.py:114-157,313-340 ; fourier/numpy.py:33-49
def fft_frequency(n, sample_rate, real, trailing=0):
    """FFT sample frequencies with ``trailing`` extra unit dimensions."""
    f = (np.fft.rfftfreq(n, d=1. / sample_rate) if real
         else np.fft.fftfreq(n, d=1. / sample_rate))
    return f.reshape(f.shape + (1,) * trailing)


def freq_dtype(dtype):
    dtype = np.dtype(dtype)
    if dtype.kind == 'f':
        return np.dtype('c{0:d}'.format(2 * dtype.itemsize))
    return dtype


def fft(a, axis=0, ortho=False):
    """Forward transform as NumpyFFTBase._cfft/_rfft (numpy.py:33-43)."""
    a = np.asarray(a)
    norm = 'ortho' if ortho else None
    if a.dtype.kind == 'f':
        return np.fft.rfft(a, axis=axis, norm=norm).astype(
            freq_dtype(a.dtype), copy=False)
    return np.fft.fft(a, axis=axis, norm=norm).astype(a.dtype, copy=False)


def ifft(a, time_dtype, n=None, axis=0, ortho=False):
    """Inverse transform as NumpyFFTBase._icfft/_irfft (numpy.py:37-49)."""
    time_dtype = np.dtype(time_dtype)
    norm = 'ortho' if ortho else None
    if time_dtype.kind == 'f':
        return np.fft.irfft(a, axis=axis, norm=norm, n=n).astype(
            time_dtype, copy=False)
    return np.fft.ifft(a, axis=axis, norm=norm).astype(time_dtype, copy=False)


# --------------------------------------------------------------------------
# generators.py:171-190
def noise_frame(seed, offset, samples_per_frame, sample_shape, dtype):
    """One frame of NoiseGenerator output, starting at sample ``offset``."""
    dtype = np.dtype(dtype)
    rng = np.random.Generator(np.random.Philox(seed))
    state = rng.bit_generator.state
    state['state']['counter'][1] = offset
    rng.bit_generator.state = state
    shape = (samples_per_frame,) + tuple(sample_shape)
    if dtype.kind == 'c':
        shape = shape[:-1] + (shape[-1] * 2,)
    numbers = rng.normal(size=shape)
    if dtype.kind == 'c':
        numbers = numbers.view(np.complex128)
    return numbers.astype(dtype, copy=False)


def noise_stream(seed, n, samples_per_frame, sample_shape=(), dtype='c8',
                 start=0):
    """Samples [start, start+n) of a NoiseGenerator stream."""
    spf = samples_per_frame
    out = np.empty((n,) + tuple(sample_shape), dtype)
    pos = start
    while pos < start + n:
        i0 = (pos // spf) * spf
        frame = noise_frame(seed, i0, spf, sample_shape, dtype)
        take = min(start + n - pos, i0 + spf - pos)
        out[pos - start:pos - start + take] = frame[pos - i0:pos - i0 + take]
        pos += take
    return out


# --------------------------------------------------------------------------
# base.py:743-795 (overlap-save framing)
def padded_framing(n_in, ih_samples_per_frame, pad_start, pad_end,
                   samples_per_frame=None, fast_len=None):
    """Frame sizes of a PaddedTaskBase (base.py:750-768).

    Returns (N, spf, n_out): input samples per frame, output samples per
    frame and total number of output samples.
    """
    if pad_start < 0 or pad_end < 0:
        raise ValueError("padding values must be 0 or positive.")
    pad = pad_start + pad_end
    if samples_per_frame is None:
        big_n = max(ih_samples_per_frame, pad * 4)
    else:
        big_n = samples_per_frame + pad
    if fast_len is not None:
        big_n = fast_len(big_n)
    assert big_n <= n_in, "time per frame larger than total time in stream"
    return big_n, big_n - pad, n_in - pad


def padded_apply(x, big_n, pad_start, pad_end, task, out_dtype=None,
                 out_sample_shape=None, ratio=1):
    """Run ``task`` over a whole stream with PaddedTaskBase framing.

    Frame i reads input [i*spf, i*spf+N); a last, partial frame is
    re-anchored to the end of the input and its first samples are skipped
    (base.py:775-795).  ``task`` gets N input samples and must return the
    spf//ratio valid output samples.
    """
    n_in = x.shape[0]
    pad = pad_start + pad_end
    spf = big_n - pad
    n_out = (n_in - pad) // ratio
    spf_out = spf // ratio
    res = None
    pos = 0
    frame = 0
    while pos < n_out:
        ih_index = frame * spf
        max_start = n_in - big_n
        if ih_index > max_start:
            skip = (ih_index - max_start) // ratio
            ih_index = max_start
        else:
            skip = 0
        y = task(x[ih_index:ih_index + big_n])
        if res is None:
            res = np.empty((n_out,) + y.shape[1:], y.dtype)
        take = min(spf_out - skip, n_out - pos)
        res[pos:pos + take] = y[skip:skip + take]
        pos += take
        frame += 1
    return res


# --------------------------------------------------------------------------
# dispersion.py:48-139
class DispersePlan:
    """Everything Disperse.__init__ decides (dispersion.py:48-113).

    Parameters mirror the reference; ``frequency_mhz`` and ``sideband``
    broadcast against ``sample_shape``; ``sample_rate_mhz`` is the rate of
    the input stream.  ``dm`` is the *dispersing* DM (Dedisperse passes -dm,
    dispersion.py:184).
    """

    def __init__(self, dm, frequency_mhz, sideband, sample_rate_mhz,
                 complex_data, n_in, ih_samples_per_frame, sample_shape=(),
                 reference_frequency_mhz=None, samples_per_frame=None,
                 fast_len=next_fast_len):
        frequency = np.asarray(frequency_mhz, dtype=float)
        sideband = np.where(np.asarray(sideband) > 0, 1, -1).astype(np.int8)
        half_rate = sample_rate_mhz / 2.
        if complex_data:
            freq_low = frequency - half_rate
            freq_high = frequency + half_rate
        else:
            freq_low = frequency + np.minimum(sideband, 0.) * half_rate
            freq_high = frequency + np.maximum(sideband, 0.) * half_rate
        if reference_frequency_mhz is None:
            reference_frequency_mhz = (freq_low + freq_high).mean() / 2.
        delay_low = dm_time_delay(dm, freq_low, reference_frequency_mhz)
        delay_high = dm_time_delay(dm, freq_high, reference_frequency_mhz)
        delay_max = max(np.max(delay_low), np.max(delay_high))
        delay_min = min(np.min(delay_low), np.min(delay_high))
        rate_hz = sample_rate_mhz * 1e6
        pad_start = int(np.ceil(delay_max * rate_hz))
        pad_end = int(np.ceil(-delay_min * rate_hz))
        if pad_start < 0:
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
        self.dm = dm
        self.frequency_mhz = frequency
        self.sideband = sideband
        self.sample_rate_mhz = sample_rate_mhz
        self.complex_data = complex_data
        self.reference_frequency_mhz = reference_frequency_mhz
        self.sample_shape = tuple(sample_shape)
        self.pad_start, self.pad_end = pad_start, pad_end
        self.sample_offset = sample_offset
        self.N, self.samples_per_frame, self.n_out = padded_framing(
            n_in, ih_samples_per_frame, pad_start, pad_end,
            samples_per_frame, fast_len)
        # Start time shift relative to the input start, in seconds
        # (dispersion.py:96 and base.py:769-770).
        self.start_offset = (sample_offset + pad_start) / rate_hz

    def phase_factor(self, dtype):
        """The chirp (dispersion.py:115-129), float64 then cast."""
        dtype = np.dtype(dtype)
        fftfreq = fft_frequency(self.N, self.sample_rate_mhz,
                                real=dtype.kind == 'f',
                                trailing=len(self.sample_shape))
        frequency = self.frequency_mhz + fftfreq * self.sideband
        phase = dm_phase_delay(self.dm, frequency,
                               self.reference_frequency_mhz)
        phase = phase * self.sideband
        if self.sample_offset != 0:
            # sample_offset / rate [us] * fftfreq [MHz] = cycles.
            phase = phase + (self.sample_offset / self.sample_rate_mhz
                             * fftfreq)
        factor = np.exp(phase * (2j * np.pi))
        return factor.astype(freq_dtype(dtype), copy=False)


def disperse(x, plan, phase_factor=None):
    """Disperse a whole stream (dispersion.py:135-139 per frame)."""
    if phase_factor is None:
        phase_factor = plan.phase_factor(x.dtype)
    sl = slice(plan.pad_start, plan.pad_start + plan.samples_per_frame)

    def task(data):
        ft = fft(data, axis=0)
        ft *= phase_factor
        return ifft(ft, data.dtype, n=plan.N, axis=0)[sl]

    return padded_apply(x, plan.N, plan.pad_start, plan.pad_end, task)


# --------------------------------------------------------------------------
# channelize.py:50-74,128-166
def channelize(x, n, samples_per_frame=1):
    """Channelize: blocks of n samples -> FFT (channelize.py:73-74)."""
    nspec = (x.shape[0] // (n * samples_per_frame)) * samples_per_frame
    data = x[:nspec * n].reshape((nspec, n) + x.shape[1:])
    return fft(data, axis=1)


def channelize_frequency(frequency, sideband, n, sample_rate, real,
                         sample_ndim):
    """Channel frequencies, FFT order (channelize.py:62-64)."""
    return (np.asarray(frequency)
            + fft_frequency(n, sample_rate, real, trailing=sample_ndim)
            * np.asarray(sideband))


def dechannelize(x, n=None, dtype=None):
    """Inverse of channelize (channelize.py:165-166)."""
    dtype = x.dtype if dtype is None else np.dtype(dtype)
    if n is None:
        n = x.shape[1]
    y = ifft(x, dtype, n=n, axis=1)
    return y.reshape((-1,) + y.shape[2:])


# --------------------------------------------------------------------------
# pfb.py:14-154
def sinc_hamming(n_tap, n_sample, sinc_scale=1.):
    """pfb.py:43-45."""
    n = n_tap * n_sample
    x = n_tap * sinc_scale * np.linspace(-0.5, 0.5, n, endpoint=False)
    return (np.sinc(x) * np.hamming(n)).reshape(n_tap, n_sample)


def pfb_framing(n_in, ih_samples_per_frame, response, samples_per_frame=None):
    """Framing of PolyphaseFilterBankSamples (pfb.py:74-89)."""
    n_tap, n = response.shape
    pad = (n_tap - 1) * n
    assert pad % 2 == 0
    if samples_per_frame is not None:
        samples_per_frame = samples_per_frame * n
    big_n, spf, n_out = padded_framing(n_in, ih_samples_per_frame,
                                       pad // 2, pad // 2, samples_per_frame)
    return big_n, spf, n_out


def pfb(x, response, ih_samples_per_frame=1, samples_per_frame=None,
        fourier=False):
    """Polyphase filter bank of a whole stream.

    Time-domain FIR (pfb.py:91-100) or Fourier-domain along the block axis
    (pfb.py:136-154), followed by Channelize(n) (pfb.py:86-87).
    """
    n_tap, n = response.shape
    pad = (n_tap - 1) * n
    big_n, spf, n_out = pfb_framing(x.shape[0], ih_samples_per_frame,
                                    response, samples_per_frame)
    n_blocks = big_n // n
    shape2 = (n_blocks, n) + x.shape[1:]
    resp = response.reshape(response.shape + (1,) * (x.ndim - 1))

    if fourier:
        long_response = np.zeros(shape2[:2] + (1,) * (x.ndim - 1), x.dtype)
        long_response[:n_tap] = resp
        ft_response_conj = fft(long_response, axis=0).conj()

    def task(data):
        data = data[:n_blocks * n].reshape(shape2)
        if fourier:
            ft = fft(data, axis=0)
            ft = ft * ft_response_conj
            result = ifft(ft, data.dtype, n=n_blocks, axis=0)
            result = result[:n_blocks + 1 - n_tap]
        else:
            result = np.empty((spf // n,) + data.shape[1:], data.dtype)
            for i in range(n_blocks + 1 - n_tap):
                result[i] = (data[i:i + n_tap] * resp).sum(0)
        return result.reshape((-1,) + result.shape[2:])

    filtered = padded_apply(x, big_n, pad // 2, pad // 2, task)
    return channelize(filtered, n, spf // n)


# --------------------------------------------------------------------------
# functions.py:15-16,132-143
def complex_square(z):
    return np.square(z.real) + np.square(z.imag)


def square(x):
    return complex_square(x) if x.dtype.kind == 'c' else np.square(x)


def power(x, axis=-1):
    """Power.task: [|X|^2, |Y|^2, Re(X Y*), Im(X Y*)] along ``axis``."""
    axis = axis % x.ndim
    assert x.shape[axis] == 2
    shape = x.shape[:axis] + (4,) + x.shape[axis + 1:]
    result = np.empty(shape, x.real.dtype)
    in_ = np.moveaxis(x, axis, 0)
    out = np.moveaxis(result, axis, 0)
    out[0] = complex_square(in_[0])
    out[1] = complex_square(in_[1])
    c = in_[0] * in_[1].conj()
    out[2] = c.real
    out[3] = c.imag
    return result


# --------------------------------------------------------------------------
# integration.py:106-303
class IntegratePlan:
    """Bin bookkeeping of Integrate without a phase callable.

    ``step`` is an integer number of upstream samples, a float time in
    seconds, or None for everything (integration.py:122-143).  ``start`` is
    an integer upstream offset or a float time in seconds since the upstream
    start (integration.py:110-120).
    """

    def __init__(self, n_in, sample_rate, step=None, start=0):
        if isinstance(start, (int, np.integer)):
            ih_start_int = int(start)
            self.ih_start = ih_start_int
            start_time = ih_start_int / sample_rate
        else:
            # seek() rounds to the nearest sample (base.py:341), after which
            # ih_start gets the fractional remainder (integration.py:115-118).
            ih_start_int = int(np.round(start * sample_rate))
            self.ih_start = (ih_start_int
                             + (start - ih_start_int / sample_rate)
                             * sample_rate)
            start_time = start
        ih_n_sample = n_in - ih_start_int
        if ih_start_int < 0 or ih_n_sample < 0:
            raise ValueError("'start' is not within the underlying stream.")
        if step is None:
            step = ih_n_sample
        if isinstance(step, (int, np.integer)):
            self.sample_rate = sample_rate / step
            n_sample = ih_n_sample / step
        else:
            self.sample_rate = 1. / step
            # (stop - start) * sample_rate, integration.py:130-136.
            n_sample = (n_in / sample_rate - start_time) * self.sample_rate
        self.start_time = start_time
        self.mean_offset_size = n_sample / ih_n_sample
        self.n_out = int(n_sample + 0.5 * self.mean_offset_size)
        assert self.n_out >= 1, \
            "time per frame larger than total time in stream"

    def offsets(self, samples):
        """Upstream offsets of output bin edges (integration.py:185-186)."""
        return np.around(np.asarray(samples) / self.mean_offset_size
                         + self.ih_start).astype(int)


def integrate(x, offsets, upstream_frame=None):
    """Sum x[offsets[k]:offsets[k+1]] for each k (integration.py:273-303).

    Summation runs per upstream frame like the reference (the _FakeOutput
    callback is called once per upstream frame slice).  Returns (sum, count).
    """
    offsets = np.asarray(offsets)
    nbin = len(offsets) - 1
    data = np.zeros((nbin,) + x.shape[1:], x.dtype)
    count = np.zeros((nbin,) + (1,) * (x.ndim - 1), int)
    rel = offsets - offsets[0]
    base = offsets[0]
    total = rel[-1]
    if upstream_frame is None:
        upstream_frame = max(total, 1)
    pos = 0
    while pos < total:
        # Upstream frames are aligned in the upstream stream.
        stop = min(total,
                   ((base + pos) // upstream_frame + 1) * upstream_frame
                   - base)
        chunk = x[base + pos:base + stop]
        b0 = np.searchsorted(rel[1:], pos, side='right')
        b1 = np.searchsorted(rel[:-1], stop, side='left')
        indices = rel[b0:b1 + 1] - pos
        indices[0] = 0
        indices[-1] = stop - pos
        data[b0:b1] += np.add.reduceat(chunk, indices[:-1])
        count[b0:b1] += np.diff(indices).reshape((-1,) + (1,) * (x.ndim - 1))
        pos = stop
    return data, count


def fold(x, offsets, n_phase, phase_of_sample, searchsorted_side='left'):
    """Fold (integration.py:380-395).

    ``phase_of_sample`` maps absolute upstream sample indices to phase in
    cycles (float64).  Time bin of a sample: for one time bin, 0; otherwise
    searchsorted(offsets[1:], i) with side='left' (the reference's quirk,
    integration.py:386, applied per output frame; here for the given
    offsets as one frame).  Returns (sum, count).
    """
    offsets = np.asarray(offsets)
    nbin = len(offsets) - 1
    data = np.zeros((nbin, n_phase) + x.shape[1:], x.dtype)
    count = np.zeros((nbin, n_phase) + (1,) * (x.ndim - 1), int)
    items = np.arange(offsets[0], offsets[-1])
    if nbin == 1:
        tbin = np.zeros(len(items), int)
    else:
        tbin = np.searchsorted(offsets[1:], items, side=searchsorted_side)
    phases = phase_of_sample(items)
    pbin = ((phases % 1.) * n_phase).astype(int)
    np.add.at(data, (tbin, pbin), x[offsets[0]:offsets[-1]])
    np.add.at(count, (tbin, pbin), 1)
    return data, count


# --------------------------------------------------------------------------
# Packed payloads.  PARITY UNPINNED: the decoding
# lives in the third-party `baseband` package (VDIFPayload._decoders, reused
# by io/hdf5/payload.py:165-166), which is not vendored with the reference and
# is absent here.  This restates its documented VDIF convention: bps-bit
# codes, the first value in the least significant bits of each byte, mapped
# through a table of levels.
def decode_payload(words, bps, levels, n=None):
    """Values of the ``bps``-bit codes in the byte stream ``words``."""
    b = np.ascontiguousarray(words).reshape(-1).view(np.uint8)
    shifts = np.arange(0, 8, bps)
    codes = ((b[:, np.newaxis] >> shifts) & ((1 << bps) - 1)).reshape(-1)
    values = np.asarray(levels)[codes]
    return values if n is None else values[:n]
