"""Parity of the C-ABI kernels (include/bbt_b200.h) against the oracle.

Every test runs on the host-thread emulation of the kernels (CPU, small
sizes) and, marked ``gpu``, through the CUDA library on the B200.

Tolerances (BASELINE.json north_star): complex64 voltages max-abs error
<= 1e-5 x signal RMS; detected / integrated power <= 1e-5 relative;
indices, counts and fold bins bit-exact.
"""
import ctypes

import numpy as np
import pytest

import bbt_oracle as orc

VTOL = 1e-5   # x RMS, voltages
PTOL = 1e-5   # relative, powers


def rms(a):
    return float(np.sqrt(np.mean(np.abs(a) ** 2)))


def assert_voltage(got, want, tol=VTOL):
    assert got.shape == want.shape
    err = float(np.max(np.abs(got - want))) if want.size else 0.
    assert err <= tol * rms(want), (err, rms(want), err / rms(want))


def assert_power(got, want, tol=PTOL):
    assert got.shape == want.shape
    scale = float(np.max(np.abs(want)))
    err = float(np.max(np.abs(got - want)))
    assert err <= tol * scale, (err, scale, err / scale)


def cnoise(rng, shape):
    return (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype('c8')


def make_fft_plan(lib, n, outer, inner, kind, direction, scale):
    plan = ctypes.c_void_p()
    lib.check(lib.bbt_fft_plan_create(ctypes.byref(plan), n, outer, inner,
                                      kind, direction, scale))
    return plan


def run_fft(b, x, n, outer, inner, kind, direction, scale, out_shape,
            out_dtype):
    lib = b.lib
    plan = make_fft_plan(lib, n, outer, inner, kind, direction, scale)
    try:
        d_in = b.to_dev(x)
        d_out = b.empty(out_shape, out_dtype)
        wb = lib.bbt_fft_plan_work_bytes(plan)
        work = b.empty((max(wb, 8) // 8,), 'c8')
        lib.check(lib.bbt_fft_exec(plan, b.ptr(d_in), b.ptr(d_out),
                                   b.ptr(work), b.stream))
        b.sync()
        return b.to_host(d_out)
    finally:
        lib.bbt_fft_plan_destroy(plan)


@pytest.mark.parametrize('log2n', list(range(1, 15)))
def test_fft_c2c_contiguous(backend, log2n):
    rng = np.random.default_rng(log2n)
    n = 1 << log2n
    outer = 3 if log2n > 10 else 37
    x = cnoise(rng, (outer, n))
    got = run_fft(backend, x, n, outer, 1, 0, 0, 1., x.shape, 'c8')
    assert_voltage(got, np.fft.fft(x, axis=1))
    got = run_fft(backend, x, n, outer, 1, 0, 1, 1. / n, x.shape, 'c8')
    assert_voltage(got, np.fft.ifft(x, axis=1))


@pytest.mark.parametrize('log2n,inner', [(1, 5), (4, 3), (6, 16), (8, 7),
                                         (10, 2), (12, 4), (13, 3)])
def test_fft_c2c_strided(backend, log2n, inner):
    rng = np.random.default_rng(100 + log2n)
    n = 1 << log2n
    outer = 2
    x = cnoise(rng, (outer, n, inner))
    got = run_fft(backend, x, n, outer, inner, 0, 0, 1., x.shape, 'c8')
    assert_voltage(got, np.fft.fft(x, axis=1))
    s = 1. / np.sqrt(n)
    got = run_fft(backend, x, n, outer, inner, 0, 1, s, x.shape, 'c8')
    assert_voltage(got, np.fft.ifft(x, axis=1, norm='ortho'))


@pytest.mark.parametrize('log2n,inner', [(3, 1), (7, 1), (10, 3), (11, 1)])
def test_fft_real(backend, log2n, inner):
    rng = np.random.default_rng(200 + log2n)
    n = 1 << log2n
    outer = 5
    x = rng.normal(size=(outer, n, inner)).astype('f4')
    want = np.fft.rfft(x, axis=1).astype('c8')
    got = run_fft(backend, x, n, outer, inner, 1, 0, 1., want.shape, 'c8')
    assert_voltage(got, want)
    back = run_fft(backend, want, n, outer, inner, 2, 1, 1. / n, x.shape,
                   'f4')
    assert_voltage(back, np.fft.irfft(want, n=n, axis=1).astype('f4'))


@pytest.mark.parametrize('log2n', [15, 16, 18])
def test_fft_large(backend, log2n):
    if log2n > 15 and not backend.big:
        pytest.skip('too slow on host threads')
    rng = np.random.default_rng(300 + log2n)
    n = 1 << log2n
    outer = 2
    x = cnoise(rng, (outer, n))
    got = run_fft(backend, x, n, outer, 1, 0, 0, 1., x.shape, 'c8')
    assert_voltage(got, np.fft.fft(x, axis=1))
    got = run_fft(backend, x, n, outer, 1, 0, 1, 1. / n, x.shape, 'c8')
    assert_voltage(got, np.fft.ifft(x, axis=1))


def test_fft_errors(backend):
    lib = backend.lib
    plan = ctypes.c_void_p()
    assert lib.bbt_fft_plan_create(ctypes.byref(plan), 1, 1, 1, 0, 0,
                                   1.) == -2
    with pytest.raises(NotImplementedError):
        lib.check(-2)
    assert lib.bbt_fft_plan_create(ctypes.byref(plan), 16, 1, 0, 0, 0,
                                   1.) == -1


@pytest.mark.parametrize('n,inner', [(7919, 1), (100, 3), (3, 2), (1000, 1),
                                     (135, 4), (19324, 1)])
def test_fft_any_length(backend, n, inner):
    """Lengths that are not powers of two (Bluestein on the device): the
    prime length of the reference's plugin test
    (fourier/tests/test_fourier.py:49,89), its default 2-3-5-7-smooth frame
    lengths (19324-sample frames, tests/test_dispersion.py:64-69), with
    strides, forward and inverse, complex and real."""
    if n > 8000 and not backend.big and inner > 1:
        pytest.skip('too slow on host threads')
    rng = np.random.default_rng(n)
    outer = 2
    x = cnoise(rng, (outer, n, inner))
    got = run_fft(backend, x, n, outer, inner, 0, 0, 1., x.shape, 'c8')
    assert_voltage(got, np.fft.fft(x, axis=1))
    got = run_fft(backend, x, n, outer, inner, 0, 1, 1. / n, x.shape, 'c8')
    assert_voltage(got, np.fft.ifft(x, axis=1))
    r = rng.normal(size=(outer, n, inner)).astype('f4')
    want = np.fft.rfft(r, axis=1).astype('c8')
    got = run_fft(backend, r, n, outer, inner, 1, 0, 1., want.shape, 'c8')
    assert_voltage(got, want)
    back = run_fft(backend, want, n, outer, inner, 2, 1, 1. / n, r.shape,
                   'f4')
    assert_voltage(back, np.fft.irfft(want, n=n, axis=1).astype('f4'))


@pytest.mark.parametrize('log2n,inner', [(15, 2), (16, 3), (17, 1)])
def test_fft_large_strided_and_real(backend, log2n, inner):
    """Above the single-kernel length: the four-step transform along a strided
    axis, and real transforms through the complex one."""
    if not backend.big and (log2n > 15 or inner > 2):
        pytest.skip('too slow on host threads')
    rng = np.random.default_rng(400 + log2n)
    n, outer = 1 << log2n, 2
    x = cnoise(rng, (outer, n, inner))
    got = run_fft(backend, x, n, outer, inner, 0, 0, 1., x.shape, 'c8')
    assert_voltage(got, np.fft.fft(x, axis=1))
    got = run_fft(backend, x, n, outer, inner, 0, 1, 1. / n, x.shape, 'c8')
    assert_voltage(got, np.fft.ifft(x, axis=1))
    r = rng.normal(size=(outer, n, inner)).astype('f4')
    want = np.fft.rfft(r, axis=1).astype('c8')
    got = run_fft(backend, r, n, outer, inner, 1, 0, 1., want.shape, 'c8')
    assert_voltage(got, want)
    back = run_fft(backend, want, n, outer, inner, 2, 1, 1. / n, r.shape,
                   'f4')
    assert_voltage(back, np.fft.irfft(want, n=n, axis=1).astype('f4'))


# ------------------------------------------------------------ dedispersion
class Dd:
    """A dedispersion plan plus the oracle's plan for the same stream."""

    def __init__(self, b, n_in, sample_shape, rate_mhz, freq_mhz, sideband, dm,
                 samples_per_frame=None, fref=None, log2n1=0):
        self.b = b
        self.sample_shape = tuple(sample_shape)
        S = int(np.prod(sample_shape, dtype=int))
        self.S = S
        self.op = orc.DispersePlan(
            dm, freq_mhz, sideband, rate_mhz, True, n_in, 1,
            sample_shape=sample_shape, reference_frequency_mhz=fref,
            samples_per_frame=samples_per_frame, fast_len=orc.next_pow2)
        op = self.op
        freq = np.broadcast_to(np.asarray(freq_mhz, float), sample_shape)
        sb = np.broadcast_to(np.where(np.asarray(sideband) > 0, 1, -1),
                             sample_shape)
        ref = np.broadcast_to(np.asarray(op.reference_frequency_mhz, float),
                              sample_shape)
        keys = list(zip(freq.ravel().tolist(), ref.ravel().tolist(),
                        sb.ravel().tolist()))
        uniq = sorted(set(keys))
        smap = np.array([uniq.index(k) for k in keys], np.int32)
        self.n_chirp = len(uniq)
        f = np.array([u[0] for u in uniq], float)
        r = np.array([u[1] for u in uniq], float)
        s = np.array([u[2] for u in uniq], np.int8)
        self.plan = ctypes.c_void_p()
        lib = b.lib
        lib.check(lib.bbt_dedisperse_plan_create(
            ctypes.byref(self.plan), op.N, S, op.pad_start,
            op.samples_per_frame, self.n_chirp,
            smap.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            f.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            r.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            s.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)),
            float(dm), float(rate_mhz), float(op.sample_offset), log2n1))
        self.smap = smap

    def chirp(self):
        out = np.empty((self.n_chirp, self.op.N), 'c8')
        self.b.lib.check(self.b.lib.bbt_dedisperse_plan_get_response(
            self.plan, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def run(self, x):
        """Whole stream, as PaddedTaskBase frames it (base.py:775-795)."""
        b, lib, op = self.b, self.b.lib, self.op
        n_in = x.shape[0]
        S, N, spf = self.S, op.N, op.samples_per_frame
        n_out = op.n_out
        n_full = (n_in - N) // spf + 1          # frames that start in range
        d_in = b.to_dev(x.reshape(n_in, S))
        d_out = b.zeros((n_out, S), 'c8')
        wb = lib.bbt_dedisperse_work_bytes(self.plan, n_full)
        work = b.empty((max(wb, 8) // 8,), 'c8')
        lib.check(lib.bbt_dedisperse_exec(
            self.plan, b.ptr(d_in), spf * S, n_full, 0, b.ptr(d_out),
            spf * S, b.ptr(work), b.stream))
        done = n_full * spf
        if done < n_out:
            # Last, partial frame: re-anchored to the end of the input.
            skip = n_full * spf - (n_in - N)
            off_in = (n_in - N) * S * 8
            off_out = done * S * 8
            lib.check(lib.bbt_dedisperse_exec(
                self.plan, ctypes.c_void_p(b.ptr(d_in).value + off_in), 0, 1,
                skip, ctypes.c_void_p(b.ptr(d_out).value + off_out), 0,
                b.ptr(work), b.stream))
        b.sync()
        return b.to_host(d_out).reshape((n_out,) + self.sample_shape)

    def close(self):
        self.b.lib.bbt_dedisperse_plan_destroy(self.plan)


DD_CASES = [
    # n_in, sample_shape, rate, freq, sideband, dm, spf, log2n1
    dict(n_in=5000, sample_shape=(), rate_mhz=1., freq_mhz=300.,
         sideband=1, dm=3., spf=1024 - 28),
    dict(n_in=3000, sample_shape=(2,), rate_mhz=1., freq_mhz=300.,
         sideband=np.array([1, -1]), dm=-5., spf=512 - 46),
    dict(n_in=9000, sample_shape=(3, 2), rate_mhz=2.,
         freq_mhz=np.array([[400.], [402.], [404.]]), sideband=1, dm=0.4,
         spf=None),
]


@pytest.mark.parametrize('case', range(len(DD_CASES)))
def test_dedisperse_small(backend, case):
    c = dict(DD_CASES[case])
    spf = c.pop('spf')
    rng = np.random.default_rng(400 + case)
    dd = Dd(backend, samples_per_frame=spf, **c)
    try:
        assert dd.op.N <= 8192
        # Chirp as Disperse.phase_factor (float64 -> complex64).
        want = dd.op.phase_factor('c8')
        want = np.broadcast_to(want, (dd.op.N,) + dd.sample_shape).reshape(
            dd.op.N, -1)
        got = dd.chirp()
        for s in range(dd.S):
            assert np.max(np.abs(got[dd.smap[s]] - want[:, s])) < 2e-6
        x = cnoise(rng, (c['n_in'],) + dd.sample_shape)
        assert_voltage(dd.run(x), orc.disperse(x, dd.op))
    finally:
        dd.close()


INTER, PLANAR, HALF, FULLROW = 512, 256, 4096 | 8192, 16384   # plan hints


@pytest.mark.parametrize('log2n,S,log2n1', [
    (14, 1, 0), (14, 2, 0), (15, 3, 5), (16, 2, 3), (14, 2, 4 | INTER),
    (13, 16, INTER), (15, 3, 5 | INTER | FULLROW), (15, 5, 5 | PLANAR),
    (15, 2, 5 | HALF), (14, 16, 4 | INTER | HALF), (14, 7, 3 | INTER),
    (20, 2, 0), (20, 16, 0), (20, 16, INTER | FULLROW), (20, 16, 6 | PLANAR),
    (22, 2, 10), (24, 2, 0), (21, 3, 0),
    # planar rows of 2^11 .. 2^14 points (warp-local sub-transforms)
    (13, 2, 2 | PLANAR), (14, 1, 2 | PLANAR), (15, 1, 2 | PLANAR),
    (15, 1, 1 | PLANAR), (16, 3, 2 | PLANAR),
    # many interleaved series, not a multiple of eight: padded work buffer
    (13, 18, 3 | INTER), (14, 21, 4 | INTER), (13, 27, 2 | INTER | HALF),
    # the same in 256-thread CTAs, two per SM (rows of up to 8192 points)
    (16, 1, 3 | PLANAR | 8192), (16, 2, 3 | PLANAR | 8192),
    (15, 2, 3 | PLANAR | 8192), (24, 2, 11 | PLANAR | 8192)])
def test_dedisperse_large(backend, log2n, S, log2n1):
    if log2n > 16 and not backend.big:
        pytest.skip('too slow on host threads')
    if backend.name == 'emu' and (log2n > 14 and S > 3 or log2n * S > 250):
        pytest.skip('too slow on host threads')
    rng = np.random.default_rng(500 + log2n)
    N = 1 << log2n
    rate = 16.
    # Choose the DM so that the padding is about N/5.
    f0 = 800.
    k = 1. / 2.41e-4
    width = (1. / (f0 - rate / 2) ** 2 - 1. / (f0 + rate / 2) ** 2) * k
    dm = (N / 5) / (rate * 1e6) / width
    n_in = 2 * N + N // 3 if log2n < 24 else N + N // 2
    shape = (S,) if S > 1 else ()
    sb = np.where(np.arange(S) % 2, -1, 1) if S > 1 else 1
    probe = orc.DispersePlan(dm, f0, sb, rate, True, n_in, 1, shape,
                             fast_len=orc.next_pow2, samples_per_frame=1)
    spf = N - probe.pad_start - probe.pad_end
    dd = Dd(backend, n_in, shape, rate, f0, sb, dm, samples_per_frame=spf,
            log2n1=log2n1)
    try:
        assert dd.op.N == N
        x = cnoise(rng, (n_in,) + shape)
        got = dd.run(x)
        want = orc.disperse(x, dd.op)
        assert_voltage(got, want)
    finally:
        dd.close()


@pytest.mark.parametrize('log2n,log2n1,knobs', [
    (15, 2, dict(row_landp=1)), (15, 1, dict(row_landp=1)),
    (16, 2, dict(row_landp=1)), (15, 1, dict(row_landp=2)),
    (16, 2, dict(row_landp=2)), (15, 2, dict(row_landp=2)),
    (15, 1, dict(row2=0, row_e16=1)), (16, 2, dict(row2=0, row_e16=1))])
def test_dedisperse_row_variants(backend, log2n, log2n1, knobs):
    """Tuning knobs of the row pass compute what the default does.
    row_landp: tiles land at the pitch of the exchange matrix (1: no barrier
    between the landing zone's reads and the first exchange; 2, rows of 16384
    points: half of the next row lands in a side buffer a row ahead).
    row_e16 (with row2=0): rows of 16384 points with 16 values per thread in
    1024-thread CTAs."""
    lib = backend.lib
    rng = np.random.default_rng(900 + log2n + log2n1)
    N = 1 << log2n
    rate, f0, k = 16., 800., 1. / 2.41e-4
    width = (1. / (f0 - rate / 2) ** 2 - 1. / (f0 + rate / 2) ** 2) * k
    dm = (N / 5) / (rate * 1e6) / width
    n_in = 3 * N
    probe = orc.DispersePlan(dm, f0, 1, rate, True, n_in, 1, (),
                             fast_len=orc.next_pow2, samples_per_frame=1)
    spf = N - probe.pad_start - probe.pad_end
    defaults = dict(row_landp=0, row2=1, row_e16=0)
    for key, value in knobs.items():
        lib.check(lib.bbt_tune_set(key.encode(), value))
    try:
        dd = Dd(backend, n_in, (), rate, f0, 1, dm, samples_per_frame=spf,
                log2n1=log2n1 | PLANAR)
        try:
            x = cnoise(rng, (n_in,))
            assert_voltage(dd.run(x), orc.disperse(x, dd.op))
        finally:
            dd.close()
    finally:
        for key in knobs:
            lib.check(lib.bbt_tune_set(key.encode(), defaults[key]))


def test_dedisperse_set_response(backend):
    """An arbitrary response through the same plan (Convolve-style)."""
    rng = np.random.default_rng(77)
    N = 1 << 14
    n_in = 2 * N
    dd = Dd(backend, n_in, (), 1., 300., 1, 1., samples_per_frame=N - 600)
    try:
        assert dd.op.N == N
        resp = cnoise(rng, (1, N))
        backend.lib.check(backend.lib.bbt_dedisperse_plan_set_response(
            dd.plan, resp.ctypes.data_as(ctypes.c_void_p)))
        np.testing.assert_array_equal(dd.chirp(), resp)
        x = cnoise(rng, (n_in,))
        want = orc.disperse(x, dd.op, phase_factor=resp[0])
        assert_voltage(dd.run(x), want)
    finally:
        dd.close()


@pytest.mark.parametrize('n_frames,S,cut', [
    (1, 1, 300), (4, 2, 300), (5, 3, 300), (3, 1, 301), (4, 3, 301)])
def test_real_frames_in_pairs(backend, n_frames, S, cut):
    """A real-valued stream through the complex plan two frames at a time
    (bbt_pair_frames_exec / bbt_unpair_frames_exec): with a real response --
    the Hermitian extension of the rfft phase factor -- the real and imaginary
    parts of a complex frame are convolved separately, so each real frame
    comes out as irfft(rfft(x) * factor) (dispersion.py:135-139 with
    fourier/numpy.py:41-49).  Odd frame counts leave the last imaginary part
    empty."""
    b, lib = backend, backend.lib
    rng = np.random.default_rng(100 * n_frames + S)
    # (An odd number of values per frame takes the kernels' scalar path.)
    N, pad_start, spf = 2048, 100, 2048 - cut
    n_in = (n_frames - 1) * spf + N
    x = rng.standard_normal((n_in, S)).astype('f4')
    half = np.exp(2j * np.pi * rng.uniform(size=(1, N // 2 + 1)))
    half[0, 0] = 1.
    half[0, -1] = -1.
    full = np.empty((1, N), 'c8')
    full[:, :N // 2 + 1] = half
    full[:, N // 2 + 1:] = half[:, N // 2 - 1:0:-1].conj()
    smap = np.zeros(S, np.int32)
    one = np.array([300.])
    plan = ctypes.c_void_p()
    dbl = ctypes.POINTER(ctypes.c_double)
    lib.check(lib.bbt_dedisperse_plan_create(
        ctypes.byref(plan), N, S, pad_start, spf, 1,
        smap.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
        one.ctypes.data_as(dbl), one.ctypes.data_as(dbl),
        np.array([1], np.int8).ctypes.data_as(ctypes.POINTER(ctypes.c_int8)),
        1., 1., 0., 0))
    try:
        lib.check(lib.bbt_dedisperse_plan_set_response(
            plan, full.ctypes.data_as(ctypes.c_void_p)))
        n_pairs = (n_frames + 1) // 2
        d_in = b.to_dev(x)
        z = b.empty((n_pairs * N, S), 'c8')
        lib.check(lib.bbt_pair_frames_exec(b.ptr(d_in), b.ptr(z), n_in, spf,
                                           N, S, n_frames, b.stream))
        b.sync()
        zh = b.to_host(z).reshape(n_pairs, N, S)
        for p in range(n_pairs):
            np.testing.assert_array_equal(
                zh[p].real, x[2 * p * spf:2 * p * spf + N])
            second = np.zeros((N, S), 'f4')
            if 2 * p + 1 < n_frames:
                second = x[(2 * p + 1) * spf:(2 * p + 1) * spf + N]
            np.testing.assert_array_equal(zh[p].imag, second)
        w = b.empty((n_pairs * spf, S), 'c8')
        wb = lib.bbt_dedisperse_work_bytes(plan, n_pairs)
        work = b.empty((max(wb, 8) // 8,), 'c8')
        lib.check(lib.bbt_dedisperse_exec(plan, b.ptr(z), N * S, n_pairs, 0,
                                          b.ptr(w), spf * S, b.ptr(work),
                                          b.stream))
        y = b.zeros((n_frames * spf, S), 'f4')
        lib.check(lib.bbt_unpair_frames_exec(b.ptr(w), b.ptr(y), spf, S,
                                             n_frames, b.stream))
        b.sync()
        got = b.to_host(y)
        for f in range(n_frames):
            frame = x[f * spf:f * spf + N].astype('f8')
            want = np.fft.irfft(np.fft.rfft(frame, axis=0) * half[0][:, None],
                                n=N, axis=0)[pad_start:pad_start + spf]
            assert_voltage(got[f * spf:(f + 1) * spf], want)
    finally:
        lib.bbt_dedisperse_plan_destroy(plan)


# ---------------------------------------------------------------- detection
@pytest.mark.parametrize('a,b', [(1000, 1), (37, 5), (1, 300)])
def test_power(backend, a, b):
    rng = np.random.default_rng(a)
    x = cnoise(rng, (a, 2, b))
    d_in = backend.to_dev(x)
    d_out = backend.empty((a, 4, b), 'f4')
    backend.lib.check(backend.lib.bbt_power_exec(
        backend.ptr(d_in), backend.ptr(d_out), a, b, backend.stream))
    backend.sync()
    assert_power(backend.to_host(d_out), orc.power(x, axis=1))


def test_square(backend):
    rng = np.random.default_rng(5)
    x = cnoise(rng, (777,))
    d_in = backend.to_dev(x)
    d_out = backend.empty((777,), 'f4')
    backend.lib.check(backend.lib.bbt_square_exec(
        backend.ptr(d_in), backend.ptr(d_out), 777, 1, backend.stream))
    backend.sync()
    assert_power(backend.to_host(d_out), orc.square(x))
    xr = x.real.copy()
    d_in = backend.to_dev(xr)
    backend.lib.check(backend.lib.bbt_square_exec(
        backend.ptr(d_in), backend.ptr(d_out), 777, 0, backend.stream))
    backend.sync()
    assert_power(backend.to_host(d_out), orc.square(xr))


@pytest.mark.parametrize('n,m,n_spec', [(16, 1, 40), (64, 3, 9), (1024, 1, 5),
                                        (1024, 8, 3), (4096, 1, 2)])
def test_channelize_power(backend, n, m, n_spec):
    rng = np.random.default_rng(n + m)
    x = cnoise(rng, (n_spec * n, m, 2))
    d_in = backend.to_dev(x)
    d_out = backend.empty((n_spec, n, m, 4), 'f4')
    backend.lib.check(backend.lib.bbt_channelize_power_exec(
        backend.ptr(d_in), backend.ptr(d_out), n, m, n_spec, backend.stream))
    backend.sync()
    want = orc.power(orc.channelize(x, n), axis=-1)
    assert_power(backend.to_host(d_out), want)


@pytest.mark.parametrize('n,m,n_spec,ratio', [(16, 2, 200, 7.8125),
                                              (64, 1, 50, 2.26),
                                              (1024, 8, 17, 4.),
                                              # narrow samples: bulk-copy path
                                              (1024, 1, 23, 5.5),
                                              (1024, 2, 19, 4.),
                                              (1024, 4, 9, 3.),
                                              (512, 1, 40, 13.)])
def test_channelize_power_integrate(backend, n, m, n_spec, ratio):
    rng = np.random.default_rng(n + m + 1)
    x = cnoise(rng, (n_spec * n, m, 2))
    n_bins = int(n_spec / ratio)
    offsets = np.around(np.arange(n_bins + 1) * ratio).astype(np.int64)
    assert offsets[-1] <= n_spec
    spectra = orc.power(orc.channelize(x, n), axis=-1)
    want, wcount = orc.integrate(spectra, offsets)
    d_in = backend.to_dev(x)
    d_off = backend.to_dev(offsets)
    d_sum = backend.zeros((n_bins, n, m, 4), 'f4')
    d_cnt = backend.zeros((n_bins,), 'i8')
    # Two calls, splitting the spectra mid-bin, as successive frames would.
    split = n_spec // 3
    for j0, j1 in ((0, split), (split, n_spec)):
        ptr = ctypes.c_void_p(backend.ptr(d_in).value + j0 * n * m * 2 * 8)
        backend.lib.check(backend.lib.bbt_channelize_power_integrate_exec(
            ptr, n, m, j1 - j0, j0, backend.ptr(d_off), 0, n_bins,
            backend.ptr(d_sum), backend.ptr(d_cnt), 0, backend.stream))
    backend.sync()
    np.testing.assert_array_equal(backend.to_host(d_cnt), wcount.ravel())
    assert_power(backend.to_host(d_sum), want)


@pytest.mark.parametrize('n,inner,ratio', [(1000, 1, 2.26), (333, 12, 10.),
                                           (64, 4096, 7.8125)])
def test_integrate(backend, n, inner, ratio):
    rng = np.random.default_rng(n)
    x = rng.normal(size=(n, inner)).astype('f4') ** 2
    n_bins = int(n / ratio)
    offsets = np.around(np.arange(n_bins + 1) * ratio).astype(np.int64) + 1
    offsets = offsets[offsets <= n]
    n_bins = len(offsets) - 1
    want, wcount = orc.integrate(x, offsets)
    d_in = backend.to_dev(x)
    d_off = backend.to_dev(offsets)
    d_sum = backend.zeros((n_bins, inner), 'f4')
    d_cnt = backend.zeros((n_bins,), 'i8')
    split = n // 2
    for i0, i1 in ((0, split), (split, n)):
        ptr = ctypes.c_void_p(backend.ptr(d_in).value + i0 * inner * 4)
        backend.lib.check(backend.lib.bbt_integrate_exec(
            ptr, i1 - i0, inner, i0, backend.ptr(d_off), 0, n_bins,
            backend.ptr(d_sum), backend.ptr(d_cnt), 0, backend.stream))
    backend.sync()
    np.testing.assert_array_equal(backend.to_host(d_cnt), wcount.ravel())
    assert_power(backend.to_host(d_sum), want)


def poly_phase(coef, i, i_ref, rate):
    """Oracle phase: Horner in float64, same operation order as the kernel."""
    dt = (i.astype(np.float64) - i_ref) / rate
    ph = np.full(dt.shape, coef[-1])
    for c in coef[-2::-1]:
        ph = ph * dt + c
    return ph


@pytest.mark.parametrize('power', [0, 1])
@pytest.mark.parametrize('n_phase,n_tbin', [(50, 1), (512, 3)])
def test_fold(backend, power, n_phase, n_tbin):
    rng = np.random.default_rng(n_phase + power)
    n = 20000
    rate = 10000.
    coef = np.array([0.25, 29.946923, -3.77535e-10 / 2.])
    inner = 4
    if power:
        x = cnoise(rng, (n, 1, 2))
        xin = orc.power(x, axis=-1).reshape(n, 4)
    else:
        xin = (rng.normal(size=(n, inner)) ** 2).astype('f4')
        x = xin
    offsets = np.linspace(7, n - 11, n_tbin + 1).round().astype(np.int64)
    i_ref = -12345.25
    want, wcount = orc.fold(
        xin, offsets, n_phase,
        lambda i: poly_phase(coef, i, i_ref, rate), 'left')
    # The reference's searchsorted(side='left') convention, on the host.
    lo = offsets[:-1].copy()
    hi = offsets[1:].copy()
    if n_tbin > 1:
        # sample i goes to bin searchsorted(offsets[1:], i, 'left'): sample
        # equal to an inner edge stays in the earlier bin.
        lo[1:] += 1
        hi[:-1] += 1
        hi[-1] = offsets[-1]
    d_in = backend.to_dev(x)
    d_lo, d_hi = backend.to_dev(lo), backend.to_dev(hi)
    d_sum = backend.zeros((n_tbin, n_phase, inner), 'f4')
    d_cnt = backend.zeros((n_tbin, n_phase), 'i8')
    item = 16 if power else inner * 4
    split = 9000
    for i0, i1 in ((0, split), (split, n)):
        ptr = ctypes.c_void_p(backend.ptr(d_in).value + i0 * item)
        backend.lib.check(backend.lib.bbt_fold_exec(
            ptr, power, i1 - i0, inner, i0, i0, backend.ptr(d_lo),
            backend.ptr(d_hi), 0, n_tbin, None,
            coef.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(coef),
            i_ref, rate, n_phase, backend.ptr(d_sum), backend.ptr(d_cnt),
            backend.stream))
    backend.sync()
    got_cnt = backend.to_host(d_cnt)
    if n_tbin > 1:
        # oracle.fold with side='left' drops nothing but assigns the sample
        # at offsets[-1]... compare only through the lo/hi convention above.
        pass
    np.testing.assert_array_equal(got_cnt, wcount.reshape(n_tbin, n_phase))
    assert_power(backend.to_host(d_sum), want)


@pytest.mark.parametrize('rate,coef,i_first,i_block', [
    (512e6, [0.1, 641.234567, -1.5e-3], 0, 0),
    (1e6 / 3 + 0.123, [-5.3, -77.7, 0.01, 1e-4, -2e-6, 3e-8], 10 ** 12, 0),
    (float(np.nextafter(8e6, 0)), [1e5 + 0.7, 200.25], 123456789, 0),
    (33554432., [0., 1000.], 0, 0),
    # A block cut out of a longer stream: samples are numbered from the
    # block's start for the time bins and on the whole stream's grid for
    # the phases.
    (512e6, [0.1, 641.234567, -1.5e-3], 7 * 12812320 + 1889551, 7 * 12812320),
])
def test_fold_phase_bins_exact(backend, rate, coef, i_first, i_block):
    """Phase bins are bit-exact with the float64 evaluation of the oracle for
    awkward rates, long polynomials, negative phases and large sample
    indices (the kernel divides through a reciprocal and pads the Horner
    recurrence; neither may change a bit)."""
    n = 3000000 if backend.big else 40000
    n_phase = 128
    coef = np.array(coef)
    i_ref = i_first + 1000.5
    rng = np.random.default_rng(17)
    x = rng.normal(size=(n, 4)).astype('f4')
    i = i_first + np.arange(n, dtype=np.int64)
    pbin = ((poly_phase(coef, i, i_ref, rate) % 1.) * n_phase).astype(int)
    want_cnt = np.bincount(pbin, minlength=n_phase)
    want = np.zeros((n_phase, 4))
    np.add.at(want, pbin, x.astype('f8'))
    lo = np.array([i_first - i_block], dtype=np.int64)
    hi = np.array([i_first - i_block + n], dtype=np.int64)
    d_in = backend.to_dev(x)
    d_lo, d_hi = backend.to_dev(lo), backend.to_dev(hi)
    d_sum = backend.zeros((1, n_phase, 4), 'f4')
    d_cnt = backend.zeros((1, n_phase), 'i8')
    backend.lib.check(backend.lib.bbt_fold_exec(
        backend.ptr(d_in), 0, n, 4, i_first - i_block, i_first,
        backend.ptr(d_lo), backend.ptr(d_hi), 0, 1, None,
        coef.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(coef),
        i_ref, rate, n_phase, backend.ptr(d_sum), backend.ptr(d_cnt),
        backend.stream))
    backend.sync()
    np.testing.assert_array_equal(backend.to_host(d_cnt)[0], want_cnt)
    got = backend.to_host(d_sum)[0]
    scale = np.sqrt(np.maximum(want_cnt, 1))[:, None]
    assert np.abs(got - want).max() <= 1e-5 * scale.max() * 4


@pytest.mark.parametrize('bps', [1, 2, 4, 8])
@pytest.mark.parametrize('n', [1, 13, 4096, 100003])
def test_decode(backend, bps, n):
    """Packed payload decode; exact table look-ups."""
    rng = np.random.default_rng(bps * 1000 + n)
    words = rng.integers(0, 256, (n * bps + 7) // 8, dtype=np.uint8)
    levels = rng.normal(size=1 << bps).astype('f4')
    d_w = backend.to_dev(words)
    d_l = backend.to_dev(levels)
    d_out = backend.empty((n,), 'f4')
    backend.lib.check(backend.lib.bbt_decode_exec(
        backend.ptr(d_w), backend.ptr(d_out), backend.ptr(d_l), n, bps,
        backend.stream))
    backend.sync()
    np.testing.assert_array_equal(backend.to_host(d_out),
                                  orc.decode_payload(words, bps, levels, n))
    with pytest.raises(NotImplementedError):
        backend.lib.check(backend.lib.bbt_decode_exec(
            backend.ptr(d_w), backend.ptr(d_out), backend.ptr(d_l), n, 3,
            backend.stream))
