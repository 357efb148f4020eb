"""The Task / FFTMaker API on top of the kernels, against the oracle.

Runs on the host-thread emulation of the kernels (CPU tensors stand in for
device buffers; test infrastructure only) and, marked ``gpu``, on the B200.
The cases follow the reference's own tests where their inputs are synthetic
(tests/test_dispersion.py, test_integration.py, fourier/tests/test_fourier.py).
"""
import numpy as np
import pytest

import bbt_oracle as orc

from test_kernels import assert_voltage, assert_power, cnoise, rms


@pytest.fixture
def bt(backend):
    """The package, wired to the backend under test."""
    import torch
    import baseband_tasks_b200 as pkg
    from baseband_tasks_b200 import _cabi
    saved = (_cabi._LIB, _cabi._DEVICE)
    if backend.name == 'emu':
        _cabi._LIB = backend.lib
        _cabi._DEVICE = torch.device('cpu')
    else:
        _cabi._LIB = backend.lib
        _cabi._DEVICE = backend.device
    yield pkg
    _cabi._LIB, _cabi._DEVICE = saved


START = None


def start_time(bt):
    return bt.Time(1289567655, 0.)   # 2010-11-12T13:14:15 in unix seconds


# ------------------------------------------------------------------ fourier
def test_fft_maker_registry(bt):
    from baseband_tasks_b200.fourier import (FFT_MAKER_CLASSES, fft_maker,
                                             CudaFFTMaker, FFTMakerBase)
    assert FFT_MAKER_CLASSES['cuda'] is CudaFFTMaker
    assert isinstance(fft_maker.get(), CudaFFTMaker)
    default = fft_maker.system_default
    other = CudaFFTMaker()
    with fft_maker.set(other):
        assert fft_maker.get() is other
        with fft_maker.set('cuda'):
            assert fft_maker.get() is not other
        assert fft_maker.get() is other
    assert fft_maker.get() is default
    with pytest.raises(KeyError):
        fft_maker.set('numpy')
    with pytest.raises(TypeError):
        fft_maker.set(other, threads=2)
    with pytest.raises(ValueError):
        class CudaFFTMaker(FFTMakerBase):  # noqa: F811  key already taken
            pass
    assert CudaFFTMaker.next_fast_len(130) == 256
    assert CudaFFTMaker.next_fast_len(256) == 256
    # Any length is taken (the reference's plugin test runs every maker on a
    # prime length, fourier/tests/test_fourier.py:49,89).
    x = cnoise(np.random.default_rng(7919), (7919,))
    fft = fft_maker((7919,), 'c8')
    assert_voltage(fft(x), np.fft.fft(x))
    assert_voltage(fft.inverse()(fft(x)), x)


def test_fft_object(bt):
    """After fourier/tests/test_fourier.py:89-166, in single precision."""
    from baseband_tasks_b200.fourier import fft_maker
    n = 4096
    t = np.arange(n)
    y = np.exp(2j * np.pi * 200. * t / n).astype('c8')
    fft = fft_maker((n,), 'c8', sample_rate=1e3)
    assert fft.direction == 'forward' and fft.axis == 0 and not fft.ortho
    assert fft.time_shape == (n,) and fft.frequency_shape == (n,)
    Y = fft(y)
    assert isinstance(Y, np.ndarray) and Y.dtype == np.complex64
    # A pure tone puts all power in one bin (n times the RMS of the others'
    # rounding noise), so the error is judged against the peak.
    want = np.fft.fft(y)
    assert np.abs(Y - want).max() <= 1e-5 * np.abs(want).max()
    assert np.argmax(np.abs(Y)) == 200
    np.testing.assert_allclose(fft.frequency[200], 200. * 1e3 / n)
    ifft = fft.inverse()
    assert ifft.direction == 'backward' and ifft == fft.inverse()
    assert ifft != fft
    assert_voltage(ifft(Y), y)
    # real, ortho, starting from the inverse
    x = np.random.default_rng(1).normal(size=(8, n // 4, 3)).astype('f4')
    irfft = fft_maker(x.shape, 'f4', direction='backward', axis=1, ortho=True)
    rfft = irfft.inverse()
    assert rfft.frequency_shape == (8, n // 8 + 1, 3)
    assert rfft.frequency_dtype == np.dtype('c8')
    X = rfft(x)
    assert_voltage(X, np.fft.rfft(x, axis=1, norm='ortho').astype('c8'))
    assert_voltage(irfft(X), x)
    assert rfft.frequency.shape == (n // 8 + 1, 1)
    # Parseval
    np.testing.assert_allclose(np.sum(x.astype('f8') ** 2), (
        np.sum(np.abs(X[:, 0]) ** 2) + 2 * np.sum(np.abs(X[:, 1:-1]) ** 2)
        + np.sum(np.abs(X[:, -1]) ** 2)), rtol=1e-5)
    with pytest.raises(ValueError):
        fft(y[:10])
    # device tensors stay on the device
    d = __import__('baseband_tasks_b200')._buffers.as_device(y)
    D = fft(d)
    assert not isinstance(D, np.ndarray)
    assert np.abs(D.cpu().numpy() - want).max() <= 1e-5 * np.abs(want).max()
    # noise-like data: 1e-5 of the RMS
    z = cnoise(np.random.default_rng(2), (n,))
    assert_voltage(fft(z), np.fft.fft(z))


def test_fft_large_strided(bt):
    from baseband_tasks_b200.fourier import fft_maker
    x = cnoise(np.random.default_rng(3), (1 << 14, 2))
    fft = fft_maker(x.shape, 'c8', axis=0)
    assert_voltage(fft(x), np.fft.fft(x, axis=0))
    # Above the single-kernel length along a strided axis, complex and real,
    # and a frame length of the reference's default (2-3-5-7 smooth) framing.
    x = cnoise(np.random.default_rng(4), (1 << 15, 2))
    assert_voltage(fft_maker(x.shape, 'c8', axis=0)(x), np.fft.fft(x, axis=0))
    r = np.random.default_rng(5).normal(size=(1 << 15, 2)).astype('f4')
    rfft = fft_maker(r.shape, 'f4', axis=0)
    R = rfft(r)
    assert R.shape == ((1 << 14) + 1, 2)
    assert_voltage(R, np.fft.rfft(r, axis=0).astype('c8'))
    assert_voltage(rfft.inverse()(R), r)
    s = cnoise(np.random.default_rng(6), (3, 19324))
    assert_voltage(fft_maker(s.shape, 'c8', axis=1)(s), np.fft.fft(s, axis=1))


# --------------------------------------------------------------------- dm
def test_dm(bt):
    """tests/test_dm.py:33-73 of the reference."""
    dm = bt.DispersionMeasure(29.1168)
    assert abs(bt.DispersionMeasure(1.).time_delay(1e6) - 1. / 2.41e-4) < 1e-9
    f, fref = np.array([320e6, 330e6]), 325e6
    d = dm.dispersion_delay_constant * 29.1168
    np.testing.assert_allclose(
        dm.time_delay(f, fref), d * (1 / 320.**2 - 1 / 325.**2) *
        np.array([1., 0]) + d * (1 / 330.**2 - 1 / 325.**2) *
        np.array([0, 1.]), rtol=1e-13)
    np.testing.assert_allclose(dm.phase_delay(f, fref),
                               orc.dm_phase_delay(29.1168, f / 1e6, 325.),
                               rtol=1e-13)
    np.testing.assert_allclose(dm.phase_factor(f, fref),
                               orc.dm_phase_factor(29.1168, f / 1e6, 325.),
                               rtol=1e-6)
    assert float(-dm) == -29.1168


# ------------------------------------------------------------ dedispersion
def noise_source(bt, n, sample_shape=(), rate=1e6, spf=1000, seed=1234,
                 **kwargs):
    return bt.NoiseGenerator((n,) + sample_shape, start_time(bt), rate,
                             samples_per_frame=spf, dtype='c8', seed=seed,
                             **kwargs)


@pytest.mark.parametrize('shape,sideband,spf', [
    ((), 1, 1024 - 28), ((2,), np.array([1, -1]), None)])
def test_dedisperse_stream(bt, shape, sideband, spf):
    n = 5000
    dm = 3.
    src = noise_source(bt, n, shape, frequency=300e6, sideband=sideband)
    dd = bt.Dedisperse(src, dm, samples_per_frame=spf)
    x = orc.noise_stream(1234, n, 1000, shape)
    op = orc.DispersePlan(-dm, 300., sideband, 1., True, n, 1000, shape,
                          samples_per_frame=spf, fast_len=orc.next_pow2)
    assert dd.samples_per_frame == op.samples_per_frame
    assert dd._ih_samples_per_frame == op.N
    assert dd.shape == (op.n_out,) + shape
    assert dd.dm == dm
    assert abs((dd.start_time - src.start_time) - op.start_offset) < 1e-12
    assert dd.sample_rate == src.sample_rate
    want = orc.disperse(x, op)
    got = dd.read()
    assert got.dtype == np.complex64 and got.shape == want.shape
    assert_voltage(got, want)
    # Arbitrary reads: across frames, and the re-anchored last frame.
    dd.seek(op.samples_per_frame - 10)
    assert_voltage(dd.read(25), want[op.samples_per_frame - 10:
                                     op.samples_per_frame + 15])
    dd.seek(-7, 2)
    assert dd.tell() == op.n_out - 7
    assert_voltage(dd.read(7), want[-7:])
    with pytest.raises(EOFError):
        dd.read(1)
    # phase factor as the reference defines it.
    pf = dd.phase_factor
    assert pf.shape == (op.N,) + shape
    assert np.max(np.abs(pf - np.broadcast_to(op.phase_factor('c8'),
                                              pf.shape))) < 2e-6
    # task on one frame.
    frame = dd.task(x[:op.N])
    assert_voltage(frame, want[:op.samples_per_frame])
    dd.close()
    with pytest.raises(ValueError):
        dd.read(1)


def test_disperse_dedisperse_roundtrip(bt):
    """tests/test_dispersion.py:103-124: giant pulse, disperse then
    dedisperse gives it back."""
    n = 40000
    data = np.zeros((n, 2), 'c8')
    data[16000] = 1.
    src = bt.ArrayStream(data, start_time(bt), 128e3, samples_per_frame=1000,
                         frequency=300e6, sideband=np.array([1, -1]))
    dm = 1000. * 0.05 / 0.039342251 * (128e3 / 128e3) * 0.2
    disp = bt.Disperse(src, dm, samples_per_frame=8192 - 4096)
    dedisp = bt.Dedisperse(disp, dm, samples_per_frame=8192 - 4096)
    assert dedisp.dm == dm and disp.dm == dm
    out = dedisp.read()
    off = round((dedisp.start_time - src.start_time) * 128e3)
    want = data[off:off + out.shape[0]]
    assert np.max(np.abs(out - want)) < 1e-2
    # All the power of the dispersed pulse is spread out.
    disp.seek(0)
    d = disp.read()
    assert np.max(np.abs(d)) < 0.2
    np.testing.assert_allclose(np.sum(np.abs(d) ** 2, axis=0), 1., rtol=1e-3)


def test_dedisperse_errors(bt):
    src = noise_source(bt, 3000)
    with pytest.raises(TypeError):
        bt.Dedisperse(src, 1.)       # no frequency
    with pytest.raises(ValueError):
        noise_source(bt, 10, frequency=300e6)   # frequency without sideband


# ------------------------------------------------------------------- chain
def make_chain(bt, x, rate, freq, dm, n_chan, spf, device_input):
    data = x
    if device_input:
        from baseband_tasks_b200 import _buffers
        data = _buffers.as_device(x)
    src = bt.ArrayStream(data, start_time(bt), rate, samples_per_frame=2048,
                         frequency=freq, sideband=1,
                         polarization=np.array(['X', 'Y']))
    dd = bt.Dedisperse(src, dm, samples_per_frame=spf)
    ch = bt.Channelize(dd, n_chan)
    pw = bt.Power(ch)
    return src, dd, ch, pw


@pytest.mark.parametrize('device_input', [False, True])
def test_chain_dedisperse_channelize_power_integrate(bt, device_input):
    rng = np.random.default_rng(11)
    from baseband_tasks_b200 import _cabi
    # (Four frames; three where the kernels run on host threads.)
    n = 12000 if _cabi._DEVICE.type == 'cpu' else 20000
    rate, dm, n_chan = 1e6, 3., 64
    freq = np.array([[300e6], [301e6], [302e6]])
    x = cnoise(rng, (n, 3, 2))
    src, dd, ch, pw = make_chain(bt, x, rate, freq, dm, n_chan, 4096 - 600,
                                 device_input)
    op = orc.DispersePlan(-dm, freq / 1e6, 1, 1., True, n, 2048, (3, 2),
                          samples_per_frame=4096 - 600,
                          fast_len=orc.next_pow2)
    y = orc.disperse(x, op)
    spectra = orc.channelize(y, n_chan)
    power = orc.power(spectra, axis=-1)
    assert ch.shape == spectra.shape and pw.shape == power.shape
    assert ch.sample_rate == rate / n_chan
    np.testing.assert_array_equal(
        ch.frequency, orc.channelize_frequency(freq, 1, n_chan, rate, False,
                                               2))
    assert list(pw.polarization) == ['XX', 'YY', 'XY', 'YX']
    assert_voltage(ch.read(), spectra)
    assert_power(pw.read(), power)
    # Integrate over 2.26 spectra per bin: uneven bins, exact counts.
    step = 2.26 * n_chan / rate
    ip = orc.IntegratePlan(power.shape[0], rate / n_chan, step)
    offsets = ip.offsets(np.arange(ip.n_out + 1))
    want, wcount = orc.integrate(power, offsets)
    it = bt.Integrate(pw, step, average=False)
    assert it._fused == 'chanpow'
    assert it.shape == (ip.n_out,) + power.shape[1:]
    res = it.read()
    np.testing.assert_array_equal(res['count'][:, 0, 0, 0], wcount.ravel())
    assert_power(res['data'], want)
    avg = bt.Integrate(pw, step).read()
    assert_power(avg, want / wcount)
    # The same with the averages formed inside the kernels.
    from baseband_tasks_b200 import integration
    saved = integration._AVERAGE_IN_KERNEL_MIN
    integration._AVERAGE_IN_KERNEL_MIN = 0
    try:
        assert_power(bt.Integrate(pw, step).read(), want / wcount)
        unfused = bt.Integrate(pw, step)
        unfused._fused, unfused._src, unfused._src_ratio = None, pw, 1
        assert_power(unfused.read(), want / wcount)
    finally:
        integration._AVERAGE_IN_KERNEL_MIN = saved
    # Unfused path gives the same.
    it2 = bt.Integrate(pw, step, average=False)
    it2._fused, it2._src, it2._src_ratio = None, pw, 1
    res2 = it2.read()
    np.testing.assert_array_equal(res2['count'], res['count'])
    assert_power(res2['data'], want)
    # Reading in pieces and from an offset.
    it.seek(3)
    part = it.read(4)
    np.testing.assert_array_equal(part['count'], res['count'][3:7])
    assert_power(part['data'], want[3:7])


def test_square_and_power_axes(bt):
    rng = np.random.default_rng(12)
    x = cnoise(rng, (1000, 2, 3))
    src = bt.ArrayStream(x, start_time(bt), 1e3, samples_per_frame=100,
                         polarization=np.array([['L'], ['R']]))
    sq = bt.Square(src)
    assert sq.polarization.tolist() == [['LL'], ['RR']]
    assert sq.dtype == np.float32
    assert_power(sq.read(), orc.square(x))
    pw = bt.Power(src)
    assert pw.shape == (1000, 4, 3)
    assert pw.polarization.tolist() == [['LL'], ['RR'], ['LR'], ['RL']]
    pw.seek(95)
    assert_power(pw.read(10), orc.power(x, axis=1)[95:105])
    with pytest.raises(ValueError):
        bt.Power(src, polarization=['LL', 'RR', 'LR'])
    real = bt.ArrayStream(x.real.copy(), start_time(bt), 1e3,
                          polarization=np.array([['L'], ['R']]))
    with pytest.raises(ValueError):
        bt.Power(real)
    assert_power(bt.Square(real).read(), x.real ** 2)


def test_dechannelize_roundtrip(bt):
    rng = np.random.default_rng(13)
    x = cnoise(rng, (64 * 33, 2))
    src = bt.ArrayStream(x, start_time(bt), 1e6, samples_per_frame=64,
                         frequency=300e6, sideband=np.array([1, -1]))
    ch = bt.Channelize(src, 64, samples_per_frame=16)
    assert ch.shape == (33 // 16 * 16, 64, 2)
    dc = ch.inverse(ch)
    assert dc.shape == (32 * 64, 2)
    assert dc.sample_rate == src.sample_rate
    assert_voltage(dc.read(), x[:32 * 64])
    np.testing.assert_array_equal(dc.frequency, src.frequency)


def test_channelize_real_stream(bt):
    """Real-valued streams (rfft/irfft, fourier/numpy.py:41-49): n//2+1
    channels; Dechannelize needs the explicit n and the real dtype
    (channelize.py:128-166)."""
    rng = np.random.default_rng(21)
    n = 128
    x = rng.normal(size=(n * 20, 3)).astype('f4')
    src = bt.ArrayStream(x, start_time(bt), 1e6, samples_per_frame=n,
                         frequency=400e6, sideband=1)
    ch = bt.Channelize(src, n, samples_per_frame=5)
    assert ch.shape == (20, n // 2 + 1, 3) and ch.dtype == np.complex64
    want = orc.channelize(x, n)
    assert_voltage(ch.read(), want.astype('c8'))
    np.testing.assert_allclose(
        np.asarray(ch.frequency).reshape(-1)[:n // 2 + 1],
        400e6 + np.fft.rfftfreq(n, 1e-6))
    with pytest.raises(ValueError):
        bt.Dechannelize(ch, dtype='f4')        # n is needed for real data
    dc = bt.Dechannelize(ch, n, dtype='f4')
    assert dc.shape == x.shape and dc.dtype == np.float32
    assert dc.sample_rate == src.sample_rate
    assert_voltage(dc.read(), x)
    assert dc.read(0).shape == (0, 3)


# ---------------------------------------------------------------- integrate
def fake_pulsar(bt, dtype='f4'):
    """tests/test_integration.py:17-43: 16000x2 @10 kHz, a pulse every 125."""
    data = np.full((16000, 2), 0.125, dtype)
    data[::125] = 10.
    return data, bt.ArrayStream(data, start_time(bt), 1e4,
                                samples_per_frame=200)


@pytest.mark.parametrize('spf', [1, 4, 10])
def test_integrate_integer_step(bt, spf):
    data, src = fake_pulsar(bt)
    it = bt.Integrate(src, 125, samples_per_frame=spf)
    assert it.shape == (128, 2) and it.sample_rate == 1e4 / 125
    assert it.start_time == src.start_time
    out = it.read()
    want = data.reshape(-1, 125, 2).mean(1)
    np.testing.assert_allclose(out, want, rtol=1e-6)
    it.seek(100)
    assert abs((it.time - src.start_time) - 100 * 125 / 1e4) < 1e-12
    np.testing.assert_allclose(it.read(3), want[100:103], rtol=1e-6)


def test_integrate_time_step_counts(bt):
    """tests/test_integration.py:212-254: 2.26-sample bins."""
    data, src = fake_pulsar(bt)
    step = 2.26 / 1e4
    it = bt.Integrate(src, step, average=False)
    assert it.dtype.names == ('data', 'count')
    out = it.read(8)
    np.testing.assert_array_equal(out['count'][:, 0],
                                  [2, 3, 2, 2, 2, 3, 2, 2])
    ip = orc.IntegratePlan(16000, 1e4, step)
    assert it.shape[0] == ip.n_out
    offsets = ip.offsets(np.arange(ip.n_out + 1))
    want, wcount = orc.integrate(data, offsets)
    it.seek(0)
    full = it.read()
    np.testing.assert_array_equal(full['count'],
                                  np.broadcast_to(wcount, full.shape))
    np.testing.assert_allclose(full['data'], want, rtol=1e-6)
    # start at an offset and at a time
    it2 = bt.Integrate(src, 10, start=25)
    assert it2.shape[0] == int((16000 - 25) / 10 + 0.5 / 10)
    np.testing.assert_allclose(it2.read(2),
                               data[25:45].reshape(2, 10, 2).mean(1),
                               rtol=1e-6)
    with pytest.raises(ValueError):
        bt.Integrate(src, 10, start=-1)


def test_integrate_whole_stream_complex(bt):
    rng = np.random.default_rng(5)
    x = cnoise(rng, (3000, 3))
    src = bt.ArrayStream(x, start_time(bt), 1e3, samples_per_frame=100)
    it = bt.Integrate(src)
    assert it.shape == (1, 3) and it.dtype == np.complex64
    assert_voltage(it.read(), x.mean(0, keepdims=True), tol=1e-4)


# --------------------------------------------------------------------- fold
@pytest.mark.parametrize('use_poly', [True, False])
@pytest.mark.parametrize('step,spf', [(None, 1), (0.2, 1), (0.2, 3)])
def test_fold(bt, use_poly, step, spf):
    data, src = fake_pulsar(bt)
    n_phase = 50
    f0 = 1e4 / 125 * 1.0000123
    t0 = src.start_time + 0.0123
    poly = bt.PolynomialPhase([0.1, f0, -1e-3], t0)
    rate = 1e4
    i_ref = poly.i_ref(src.start_time, rate)
    if use_poly:
        phase = poly
    else:
        # Any callable; evaluate with the same index-based convention so
        # the oracle below applies to both.
        def phase(t):
            idx = np.round((t - src.start_time) * rate)
            return poly.of_index(idx, i_ref, rate)
    fold = bt.Fold(src, n_phase, phase, step, samples_per_frame=spf,
                   average=False)
    if step is None:
        assert fold.shape == (1, n_phase, 2)
        offsets = np.array([0, 16000])
        frames = [offsets]
    else:
        ip = orc.IntegratePlan(16000, rate, step)
        assert fold.shape == (ip.n_out, n_phase, 2)
        offsets = ip.offsets(np.arange(ip.n_out + 1))
        frames = [offsets[i:i + spf + 1]
                  for i in range(0, ip.n_out, spf)]
    want = []
    wcount = []
    for fo in frames:
        d, c = orc.fold(data, fo, n_phase,
                        lambda i: poly.of_index(i, i_ref, rate), 'left')
        want.append(d)
        wcount.append(c)
    want = np.concatenate(want)
    wcount = np.concatenate(wcount)
    out = fold.read()
    np.testing.assert_array_equal(out['count'],
                                  np.broadcast_to(wcount, out.shape))
    np.testing.assert_allclose(out['data'], want, rtol=1e-5)
    assert out['count'].sum() == (offsets[-1] - offsets[0]) * 2
    avg = bt.Fold(src, n_phase, phase, step, samples_per_frame=spf).read()
    with np.errstate(invalid='ignore', divide='ignore'):
        np.testing.assert_allclose(avg, want / wcount, rtol=1e-5)


def test_fold_power_fused(bt):
    rng = np.random.default_rng(21)
    x = cnoise(rng, (6000, 3, 2))
    src = bt.ArrayStream(x, start_time(bt), 1e4, samples_per_frame=500,
                         polarization=np.array(['X', 'Y']))
    pw = bt.Power(src)
    poly = bt.PolynomialPhase([0., 29.946923, -3.77535e-10 / 2],
                              src.start_time)
    fold = bt.Fold(pw, 32, poly, average=False)
    assert fold._fused == 'power'
    out = fold.read()
    power = orc.power(x, axis=-1)
    want, wcount = orc.fold(power, np.array([0, 6000]), 32,
                            lambda i: poly.of_index(i, 0., 1e4))
    np.testing.assert_array_equal(out['count'],
                                  np.broadcast_to(wcount, out.shape))
    assert_power(out['data'], want)


# --------------------------------------------------------------- framework
def test_base_stream_semantics(bt):
    """read/seek/tell/time, slicing and errors (tests/test_base.py)."""
    x = cnoise(np.random.default_rng(2), (1000, 4))
    src = bt.ArrayStream(x, start_time(bt), 1e3, samples_per_frame=100,
                         frequency=np.array([300e6, 300e6, 316e6, 316e6]),
                         sideband=np.array([-1, 1, -1, 1]))
    assert src.sample_shape == (4,) and src.size == 4000 and src.ndim == 2
    assert src.complex_data and src.dtype == np.complex64
    assert src.sideband.dtype == np.int8
    assert src.stop_time - src.start_time == 1.
    assert src.seek(10) == 10 and src.seek(5, 1) == 15
    assert src.seek(-5, 'end') == 995
    assert src.seek(src.start_time + 0.5) == 500
    assert src.seek(0.25) == 250      # time offset in seconds
    assert src.tell(unit='time') - src.start_time == 0.25
    with pytest.raises(ValueError):
        src.seek(0, 3)
    sa = bt.SetAttribute(src, frequency=np.array([1e6, 2e6, 3e6, 4e6]),
                         sideband=1)
    assert sa.frequency.tolist() == [1e6, 2e6, 3e6, 4e6]
    assert np.all(sa.sideband == 1)
    sa.seek(20)
    np.testing.assert_array_equal(sa.read(5), x[20:25])
    task = bt.Task(src, lambda data: data * 2, samples_per_frame=50)
    task.seek(120)
    np.testing.assert_array_equal(task.read(100), x[120:220] * 2)
    sl = src[100:200, 1]
    assert sl.shape == (100,) and sl.frequency == 300e6
    np.testing.assert_array_equal(sl.read(), x[100:200, 1])
    np.testing.assert_array_equal(np.array(src[10:12]), x[10:12])
    with pytest.raises(AssertionError):
        src.read(out=np.empty((3, 5), 'c8'))
    with pytest.raises(TypeError):
        bt.ArrayStream(x, start_time(bt), 1e3, bla=1)
    assert 'ArrayStream' in repr(src) and 'ih' in repr(sa)


def test_padded_task_base_framing(bt):
    """tests/test_base.py:491-546: PaddedTaskBase sizes."""
    from baseband_tasks_b200.base import PaddedTaskBase
    x = np.arange(40000.).astype('f4')
    src = bt.ArrayStream(x, start_time(bt), 1e3, samples_per_frame=20000)

    class SquareHat(PaddedTaskBase):
        def __init__(self, ih, n, **kwargs):
            self._n = n
            super().__init__(ih, pad_start=n - 1, **kwargs)

        def task(self, data):
            size = data.shape[0] - self._n + 1
            return sum(data[i:i + size] for i in range(self._n))

    sh = SquareHat(src, 3)
    assert sh._ih_samples_per_frame == 20000 and sh.samples_per_frame == 19998
    assert sh.shape == (39998,)
    assert abs((sh.start_time - src.start_time) - 2 / 1e3) < 1e-12
    want = x[:-2] + x[1:-1] + x[2:]
    np.testing.assert_array_equal(sh.read(10), want[:10])
    sh.seek(-10, 2)
    np.testing.assert_array_equal(sh.read(10), want[-10:])
    nfl = orc.next_fast_len
    assert SquareHat(src, 3, samples_per_frame=128,
                     next_fast_len=nfl).samples_per_frame == 133
    assert SquareHat(src, 5, samples_per_frame=128,
                     next_fast_len=nfl)._ih_samples_per_frame == 135
    with pytest.raises(ValueError):
        SquareHat(src, 0)
    with pytest.warns(UserWarning, match='inefficient'):
        SquareHat(src, 10, samples_per_frame=8)


def test_noise_generator_reproducible(bt):
    """tests/test_generators.py:253-316."""
    nh = noise_source(bt, 10000, (2,), spf=1000, seed=12345)
    a = nh.read(2500)
    nh.seek(1500)
    b = nh.read(1000)
    np.testing.assert_array_equal(a[1500:], b)
    np.testing.assert_array_equal(
        a, orc.noise_stream(12345, 2500, 1000, (2,)))
    nh.seek(0)
    full = nh.read()
    assert abs(full.real.std() - 1.) < 0.02 and abs(full.mean()) < 0.02


# ---------------------------------------------------------------------- pfb
def test_sinc_hamming_guppi(bt):
    """tests/test_pfb.py:26-35: the GUPPI coefficients."""
    import os
    path = os.path.join(os.path.dirname(__file__), 'golden',
                        'guppi_pfb_coeffs.txt')
    coeffs = np.loadtxt(path).reshape(8, -1).T.reshape(12, 64)
    from baseband_tasks_b200.pfb import sinc_hamming
    np.testing.assert_allclose(sinc_hamming(12, 64, sinc_scale=0.95), coeffs)


@pytest.mark.parametrize('dtype,shape', [('f4', ()), ('f4', (2,)),
                                         ('f4', (3,)), ('f4', (2, 2)),
                                         ('c8', (3,))])
def test_polyphase_filter_bank(bt, dtype, shape):
    """tests/test_pfb.py:54-102 (single precision): the filter bank equals
    rfft((h * d[i:i+4]).sum(0)), at the start and at an offset."""
    from baseband_tasks_b200.pfb import sinc_hamming
    n, n_tap = 64, 4
    n_in = 40 * n
    rng = np.random.default_rng(12345)
    x = rng.normal(size=(n_in,) + shape)
    if dtype == 'c8':
        x = x + 1j * rng.normal(size=(n_in,) + shape)
    x = x.astype(dtype)
    response = sinc_hamming(n_tap, n)
    src = bt.ArrayStream(x, start_time(bt), 1e6, samples_per_frame=128,
                         frequency=300e6, sideband=1)
    for cls in (bt.PolyphaseFilterBankSamples, bt.PolyphaseFilterBank):
        pfb = cls(src, response)
        want = orc.pfb(x.astype('f8' if dtype == 'f4' else 'c16'), response,
                       ih_samples_per_frame=128)
        n_chan = n // 2 + 1 if dtype == 'f4' else n
        assert pfb.shape == want.shape == (want.shape[0], n_chan) + shape
        assert pfb.dtype == np.complex64
        assert pfb.sample_rate == 1e6 / n
        assert abs((pfb.start_time - src.start_time)
                   - (n_tap - 1) * n / 2 / 1e6) < 1e-12
        got = pfb.read()
        assert_voltage(got, want.astype('c8'))
        # Direct definition at an offset.
        h = response.reshape(response.shape + (1,) * len(shape))
        j = 17
        d = x[j * n:(j + n_tap) * n].reshape((n_tap, n) + shape)
        direct = (np.fft.rfft if dtype == 'f4' else np.fft.fft)(
            (h * d).sum(0), axis=0)
        pfb.seek(j)
        assert_voltage(pfb.read(1)[0], direct.astype('c8'))
        np.testing.assert_array_equal(
            pfb.frequency, orc.channelize_frequency(
                300e6, 1, n, 1e6, dtype == 'f4', len(shape)))
    pfb2 = bt.PolyphaseFilterBank(src, response, samples_per_frame=5)
    assert pfb2.samples_per_frame == 5
    assert_voltage(pfb2.read(), want[:pfb2.shape[0]].astype('c8'))
    if dtype == 'f4' and shape:
        # Raw 8-bit samples stay int8 on the device.
        xi = np.clip(np.round(x * 30), -127, 127).astype('i1')
        si = bt.ArrayStream(xi, start_time(bt), 1e6, samples_per_frame=128,
                            frequency=300e6, sideband=1)
        pi = bt.PolyphaseFilterBank(si, response)
        wi = orc.pfb(xi.astype('f8'), response, ih_samples_per_frame=128)
        assert pi.dtype == np.complex64 and pi.shape == wi.shape
        assert_voltage(pi.read(), wi.astype('c8'))


def test_integrate_over_phase(bt):
    """tests/test_integration.py:407-470: integrate in steps of pulse phase
    (offsets found by inverting the phase callable, integration.py:188-228)."""
    data, src = fake_pulsar(bt)
    f0 = 1e4 / 125            # one cycle per 125 samples
    t_ref = src.start_time

    def phase(t):
        return f0 * np.asarray(t - t_ref, dtype=float)

    n_phase = 5
    it = bt.Integrate(src, 1. / n_phase, phase, average=False)
    assert it.shape == (128 * n_phase, 2)
    out = it.read()
    # Every phase bin gets 25 samples; the pulse is in the first.
    np.testing.assert_array_equal(out['count'], 25)
    want = data.reshape(-1, 25, 2).sum(1)
    np.testing.assert_allclose(out['data'], want, rtol=1e-6)
    it.seek(7)
    assert abs((it.time - src.start_time) - 7 * 25 / 1e4) < 1e-9
    np.testing.assert_allclose(it.read(3)['data'], want[7:10], rtol=1e-6)


# ------------------------------------------------- real streams, convolution
def test_disperse_real_stream(bt):
    """tests/test_dispersion.py:206-240: a real-valued two-sideband stream
    (rfft / irfft in the reference)."""
    n = 30000
    rng = np.random.default_rng(31)
    x = rng.normal(size=(n, 2)).astype('f4')
    rate = 256e3
    freq = np.array([299.936e6, 300.064e6])
    sideband = np.array([1, -1])
    dm = 1000. * 0.05 / 0.039342251 * 0.01
    src = bt.ArrayStream(x, start_time(bt), rate, samples_per_frame=1000,
                         frequency=freq, sideband=sideband)
    disp = bt.Disperse(src, dm)
    op = orc.DispersePlan(dm, freq / 1e6, sideband, rate / 1e6, False, n,
                          1000, (2,), fast_len=orc.next_pow2)
    assert (disp._pad_start, disp._pad_end, disp._ih_samples_per_frame) == (
        op.pad_start, op.pad_end, op.N)
    assert disp.dtype == np.float32 and disp.shape == (op.n_out, 2)
    want = orc.disperse(x, op)
    got = disp.read()
    assert got.dtype == np.float32
    assert_voltage(got, want)
    pf = disp.phase_factor
    assert pf.shape == (op.N // 2 + 1, 2)
    assert np.max(np.abs(pf - op.phase_factor('f4'))) < 2e-6
    # and back again
    dedisp = bt.Dedisperse(disp, dm)
    back = dedisp.read()
    off = int(round((dedisp.start_time - src.start_time) * rate))
    # Overlap-save truncates the chirp's response at the padding, so the
    # round trip of a noise stream is close, not exact (the reference's own
    # round-trip test uses atol=1e-2 on a unit pulse).
    assert rms(back - x[off:off + back.shape[0]]) < 0.1 * rms(x)


@pytest.mark.parametrize('dtype', ['c8', 'f4'])
def test_convolve(bt, dtype):
    """tests/test_convolution.py:14-39 on a synthetic stream."""
    from test_kernels import rms as _rms  # noqa: F401
    rng = np.random.default_rng(41)
    x = rng.normal(size=(16000, 2))
    if dtype == 'c8':
        x = x + 1j * rng.normal(size=x.shape)
    x = x.astype(dtype)
    src = bt.ArrayStream(x, start_time(bt), 1e4, samples_per_frame=1000)
    ct = bt.Convolve(src, np.ones(3), samples_per_frame=1024 - 2)
    expected = x[:-2] + x[1:-1] + x[2:]
    data1 = ct.read()
    assert ct.tell() == ct.shape[0] == x.shape[0] - 2
    assert abs((ct.start_time - src.start_time) - 2 / 1e4) < 1e-9
    assert data1.dtype == np.dtype(dtype)
    assert np.allclose(expected, data1, atol=1e-4)
    ct.seek(-3, 2)
    assert np.allclose(expected[-3:], ct.read(), atol=1e-4)
    # Per-series responses and an offset.
    response = rng.normal(size=(5, 2))
    ct2 = bt.Convolve(src, response, offset=2)
    want = np.stack([np.convolve(x[:, i], response[:, i], mode='valid')
                     for i in range(2)], axis=1)
    assert abs((ct2.start_time - src.start_time) - 2 / 1e4) < 1e-9
    got = ct2.read()
    assert got.shape == want.shape
    assert np.allclose(got, want, atol=2e-4)


def test_disperse_samples(bt):
    """tests/test_dispersion.py:309-358: integer shifts of a giant pulse in a
    real two-sideband stream land exactly where the delays say, and
    DedisperseSamples undoes DisperseSamples exactly."""
    rate, n, gp_sample = 256e3, 328000, 128000
    x = np.zeros((n, 2), 'f4')
    x[gp_sample] = 1.
    freq = np.array([299.936e6, 300.064e6])
    sideband = np.array([1, -1])
    dm = 1000. * 0.05 / 0.039342251
    src = bt.ArrayStream(x, start_time(bt), rate, samples_per_frame=1000,
                         frequency=freq, sideband=sideband)
    for fref in (None, 300e6, 300.064e6, 200e6):
        disperse = bt.DisperseSamples(src, dm, reference_frequency=fref)
        center = (disperse.frequency
                  + disperse.sideband * disperse.sample_rate / 2.)
        delay = bt.DispersionMeasure(dm).time_delay(
            center, disperse.reference_frequency)
        t_gp = gp_sample / rate + delay         # seconds since stream start
        disperse.seek(src.start_time + t_gp.min())
        around = disperse.read(gp_sample)
        diff = int(np.round((delay.max() - delay.min()) * rate))
        expected = np.zeros_like(around)
        expected[0, t_gp.argmin()] = 1.
        expected[diff, t_gp.argmax()] = 1.
        np.testing.assert_array_equal(around, expected)
    disperse = bt.DisperseSamples(src, dm, reference_frequency=300e6)
    dedisperse = bt.DedisperseSamples(disperse, dm, reference_frequency=300e6)
    assert dedisperse.dm == dm and dedisperse._dm == -dm
    dedisperse.seek(src.start_time + gp_sample / rate)
    dedisperse.seek(-1024, 1)
    gp_dd = dedisperse.read(2048)
    np.testing.assert_array_equal(gp_dd, x[gp_sample - 1024:gp_sample + 1024])
    # complex samples, arbitrary shifts
    z = cnoise(np.random.default_rng(9), (5000, 3))
    zs = bt.ArrayStream(z, start_time(bt), 1e3, samples_per_frame=700)
    sh = bt.ShiftSamples(zs, np.array([3, -2, 0]), samples_per_frame=512)
    assert sh.shape == (4995, 3)
    assert abs((sh.start_time - zs.start_time) - 3 / 1e3) < 1e-12
    out = sh.read()
    for i, s in enumerate([3, -2, 0]):
        np.testing.assert_array_equal(out[:, i], z[3 - s:3 - s + 4995, i])


def test_pulse_stack(bt):
    """tests/test_integration.py:430-470: one profile per pulse."""
    data, src = fake_pulsar(bt)
    f0 = 1e4 / 125
    t_ref = src.start_time

    def phase(t):
        return f0 * np.asarray(t - t_ref, dtype=float)

    ps = bt.PulseStack(src, 5, phase, average=False)
    assert ps.shape == (128, 5, 2)
    assert abs(float(ps.sample_rate) - 1.) < 1e-12      # per cycle
    out = ps.read()
    np.testing.assert_array_equal(out['count'], 25)
    want = data.reshape(128, 5, 25, 2).sum(2)
    np.testing.assert_allclose(out['data'], want, rtol=1e-6)
    avg = bt.PulseStack(src, 5, phase, samples_per_frame=4)
    avg.seek(100)
    np.testing.assert_allclose(avg.read(8), want[100:108] / 25, rtol=1e-6)
    assert abs((avg.time - src.start_time) - 108 * 125 / 1e4) < 1e-9


# ------------------------------------------------------------- edge cases
def test_ragged_and_odd_shapes(bt):
    """Sample shapes that are not powers of two, reads into ``out``, streams
    too short for a frame, and an integration start given as a time."""
    rng = np.random.default_rng(51)
    n = 3 * 4096
    x = cnoise(rng, (n, 5))
    freq = 300e6 + 1e6 * np.arange(5)
    src = bt.ArrayStream(x, start_time(bt), 1e6, samples_per_frame=512,
                         frequency=freq, sideband=np.array([1, -1, 1, 1, -1]))
    dd = bt.Dedisperse(src, 1., samples_per_frame=2048 - 400)
    op = orc.DispersePlan(-1., freq / 1e6, np.array([1, -1, 1, 1, -1]), 1.,
                          True, n, 512, (5,), samples_per_frame=2048 - 400,
                          fast_len=orc.next_pow2)
    assert dd._ih_samples_per_frame == op.N
    assert dd.samples_per_frame == op.samples_per_frame
    want = orc.disperse(x, op)
    out = np.empty((1000, 5), 'c8')
    dd.seek(777)
    assert dd.read(out=out) is out
    assert_voltage(out, want[777:1777])
    assert dd.tell() == 1777
    sq = bt.Square(dd)
    it = bt.Integrate(sq, 10, start=sq.start_time + 100.4e-6)   # a time
    assert it._fused is None
    got = it.read(5)
    p = orc.square(want)
    np.testing.assert_allclose(
        got, p[100:150].reshape(5, 10, 5).mean(1), rtol=1e-5)
    assert abs((it.start_time - sq.start_time) - 100.4e-6) < 1e-12
    short = bt.ArrayStream(x[:1000], start_time(bt), 1e6, frequency=300e6,
                           sideband=1)
    with pytest.raises(AssertionError):
        bt.Dedisperse(short, 1., samples_per_frame=2048 - 400)
    with pytest.raises(AssertionError):
        dd.read(out=np.empty((10, 4), 'c8'))
    # Closing frees the plan; the stream cannot be read any more.
    dd.close()
    with pytest.raises(ValueError):
        dd.read(1)


@pytest.mark.parametrize('bps,complex_data', [(8, True), (2, False),
                                              (4, True), (1, False)])
def test_payload_stream(bt, bps, complex_data):
    """Packed payloads are decoded on the device and feed the chain."""
    rng = np.random.default_rng(bps)
    n, shape = 4096, (3,)
    x = rng.normal(size=(n,) + shape + ((2,) if complex_data else ()))
    x = (x * {1: 1., 2: 1., 4: 1., 8: 30.}[bps]).astype('f4')
    words = bt.encode_payload(x, bps)
    levels = bt.payload_levels(bps)
    want = orc.decode_payload(words, bps, levels).reshape(x.shape)
    if complex_data:
        want = want[..., 0] + 1j * want[..., 1]
    # encode picks the nearest level
    assert np.all(np.abs(want.view('f4').reshape(x.shape) - x)
                  <= np.abs(levels[:, None] - x.reshape(-1)).min(0)
                  .reshape(x.shape) + 1e-6)
    fh = bt.PayloadStream(words, bps, shape, start_time(bt), 1e6,
                          samples_per_frame=512, complex_data=complex_data,
                          frequency=300e6, sideband=1)
    assert fh.shape == (n,) + shape
    assert fh.dtype == (np.complex64 if complex_data else np.float32)
    got = fh.read()
    np.testing.assert_array_equal(got, want.astype(fh.dtype))
    fh.seek(1000)
    np.testing.assert_array_equal(fh.read(77), want[1000:1077])
    fh.seek(1001)     # 1-bit: not on a byte, read through frames
    np.testing.assert_array_equal(fh.read(600), want[1001:1601])
    if bps == 1:
        odd = bt.PayloadStream(words, bps, shape, start_time(bt), 1e6,
                               samples_per_frame=5)
        odd.seek(7)
        with pytest.raises(ValueError):
            odd.read(8)
    # and through a task
    fh.seek(0)
    sq = bt.Square(fh)
    assert_power(sq.read(), orc.square(want.astype(fh.dtype)))
    with pytest.raises(NotImplementedError):
        bt.PayloadStream(words, 3, shape, start_time(bt), 1e6)


# ------------------------------------------------- round-1 advisor findings
def test_dedisperse_complex128_stream(bt):
    """A complex128 stream: aligned, unaligned and multi-block reads all give
    the complex64 result (the kernels write complex64; the caller's complex128
    buffer is filled through a temporary)."""
    rng = np.random.default_rng(40)
    x = cnoise(rng, (5 * 4096, 2))
    kw = dict(frequency=300e6, sideband=1)
    s16 = bt.ArrayStream(x.astype('c16'), start_time(bt), 1e6,
                         samples_per_frame=1000, **kw)
    s8 = bt.ArrayStream(x, start_time(bt), 1e6, samples_per_frame=1000, **kw)
    spf = 4096 - 923
    d16 = bt.Dedisperse(s16, 3., samples_per_frame=spf)
    d8 = bt.Dedisperse(s8, 3., samples_per_frame=spf)
    want = d8.read()
    got = d16.read()
    assert got.dtype == np.complex128
    np.testing.assert_array_equal(got.astype('c8'), want)
    d16.seek(37)
    part = d16.read(2 * spf + 11)       # unaligned start, several frames
    np.testing.assert_array_equal(part.astype('c8'), want[37:37 + 2 * spf + 11])
    op = orc.DispersePlan(-3., 300., 1, 1., True, x.shape[0], 1000, (2,),
                          samples_per_frame=spf, fast_len=orc.next_pow2)
    assert_voltage(got.astype('c8'), orc.disperse(x, op))


def test_dedisperse_many_short_frames(bt, monkeypatch):
    """More frames in one read than a single launch takes (65535)."""
    n_frames = 70000
    N, spf = 16, 14                 # pads 1 + 1
    n = (n_frames - 1) * spf + N
    rng = np.random.default_rng(41)
    x = cnoise(rng, (n,))
    src = bt.ArrayStream(x, start_time(bt), 1e6, samples_per_frame=n,
                         frequency=300e6, sideband=1)
    dd = bt.Dedisperse(src, 0.0045, samples_per_frame=spf)
    assert (dd._ih_samples_per_frame, dd._pad_start, dd._pad_end) == (N, 1, 1)
    got = dd.read()
    assert got.shape == (n_frames * spf,)
    op = orc.DispersePlan(-0.0045, 300., 1, 1., True, n, n, (),
                          samples_per_frame=spf, fast_len=orc.next_pow2)
    assert (op.pad_start, op.pad_end, op.N) == (1, 1, N)
    assert_voltage(got, orc.disperse(x, op))


def test_integrate_long_channelizer_unfused(bt):
    """Channelizers longer than the fused kernel takes are read unfused."""
    n_chan = 32768
    rng = np.random.default_rng(42)
    x = cnoise(rng, (3 * n_chan, 2))
    src = bt.ArrayStream(x, start_time(bt), 1e6, samples_per_frame=n_chan,
                         polarization=np.array(['X', 'Y']))
    it = bt.Integrate(bt.Power(bt.Channelize(src, n_chan)), 3, average=False)
    assert it._fused is None
    short = bt.Integrate(bt.Power(bt.Channelize(src, 1024)), 3)
    assert short._fused == 'chanpow'
    got = it.read()
    power = orc.power(orc.channelize(x, n_chan), axis=-1)
    want, count = orc.integrate(power, np.array([0, 3]))
    assert np.all(got['count'] == 3)
    assert_power(got['data'], want)


@pytest.mark.parametrize('shape', [(2,), (3, 2), (9, 2)])
def test_power_fused_into_dedispersion(bt, shape):
    """Power straight after Dedisperse is formed by the last dedispersion
    pass (bbt_dedisperse_power_exec): same numbers as the two tasks apart
    (functions.py:138-142 after dispersion.py:135-139), for whole reads,
    reads across frame boundaries and the partial last frame."""
    rate, dm = 1e6, 2.
    n_chan = int(np.prod(shape[:-1], dtype=int))
    freq = (300e6 + 1e6 * np.arange(n_chan)).reshape(shape[:-1] + (1,))
    # (Host-thread emulation: the many-series cases on shorter frames, which
    # go through the same three passes.)
    from baseband_tasks_b200 import _cabi
    emulated = _cabi._DEVICE.type == 'cpu'
    nf = 8192 if emulated and n_chan > 1 else 32768
    spf = nf - 1000
    n = 3 * nf + 5000
    x = cnoise(np.random.default_rng(5), (n,) + shape)
    kw = dict(frequency=freq if n_chan > 1 else 300e6, sideband=1,
              polarization=np.array(['X', 'Y']))
    src = bt.ArrayStream(x, start_time(bt), rate, samples_per_frame=4096, **kw)
    dd = bt.Dedisperse(src, dm, samples_per_frame=spf)
    nf = dd._ih_samples_per_frame    # (more with the delay between channels)
    fused = bt.Power(dd)
    assert fused._fused
    src2 = bt.ArrayStream(x, start_time(bt), rate, samples_per_frame=4096, **kw)
    dd2 = bt.Dedisperse(src2, dm, samples_per_frame=spf)
    v = dd2.read()
    want = orc.power(v.astype(np.complex128), axis=v.ndim - 1)
    assert fused.shape == want.shape
    got = fused.read()
    assert got.dtype == np.float32
    assert_power(got, want)
    fused.seek(spf - 768)       # across a frame boundary
    assert_power(fused.read(2000), want[spf - 768:spf + 1232])
    # The voltages are still what they were.
    dd.seek(100)
    assert_voltage(dd.read(50), v[100:150])
    # A read that goes in several blocks (products written straight into
    # slices of the result).
    from baseband_tasks_b200 import base
    saved = base.BLOCK_BYTES
    base.BLOCK_BYTES = (nf + nf // 4) * 16 * (fused.shape[1]
                                              if fused.ndim > 2 else 1)
    try:
        fused.seek(0)
        assert_power(fused.read(), want)
    finally:
        base.BLOCK_BYTES = saved
