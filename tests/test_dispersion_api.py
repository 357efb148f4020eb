"""Disperse/Dedisperse on the reference's giant-pulse stream
(tests/test_dispersion.py:25-190): 164000 x 2 complex samples at 128 kHz, a
single unit sample at 64000, upper and lower sideband around 300 MHz, and a
dispersion measure that delays by 0.05 s over the band."""
import numpy as np
import pytest

from test_tasks import bt, start_time  # noqa: F401  (fixture)

RATE = 128e3
GP_SAMPLE = 64000
DM = 1000. * 0.05 / 0.039342251
REFERENCE_FREQUENCIES = (None, 300e6, 300.064e6, 299.936e6, 300.128e6,
                         300.123456789e6, 299.872e6)


def giant_pulse(bt):
    def make(sh):
        data = np.empty((sh.samples_per_frame,) + sh.shape[1:], sh.dtype)
        do_gp = sh.tell() + np.arange(sh.samples_per_frame) == GP_SAMPLE
        data[...] = do_gp[:, np.newaxis]
        return data
    return bt.StreamGenerator(make, (164000, 2), start_time(bt), RATE,
                              samples_per_frame=1000, dtype=np.complex64,
                              frequency=300e6, sideband=np.array((1, -1)))


def test_time_delay(bt):
    dm = bt.DispersionMeasure(DM)
    delay = dm.time_delay(300e6 - RATE / 2, 300e6 + RATE / 2)
    assert abs(float(delay) - 0.05) < 1e-9


@pytest.mark.parametrize('reference_frequency', REFERENCE_FREQUENCIES)
def test_disperse(bt, reference_frequency):
    gp = giant_pulse(bt)
    dm = bt.DispersionMeasure(DM)
    disperse = bt.Disperse(gp, dm, reference_frequency=reference_frequency)
    # start time moves by the delay of the lowest frequency
    offset = disperse.start_time - gp.start_time
    expected = float(dm.time_delay(299.936e6, disperse.reference_frequency))
    assert abs(offset - expected) < 1. / RATE
    # the pulse is smeared over 0.05 s around its arrival time at the
    # reference frequency
    delay = float(dm.time_delay(300e6, disperse.reference_frequency))
    disperse.seek(gp.start_time + GP_SAMPLE / RATE + delay)
    disperse.seek(-GP_SAMPLE // 2, 1)
    around = disperse.read(GP_SAMPLE)
    p = (np.abs(around) ** 2).reshape(-1, 10, GP_SAMPLE // 20 // 10, 2).sum(2)
    assert np.all(p[:9].sum(1) < 0.005)
    assert np.all(p[11:].sum(1) < 0.005)
    assert np.all(p[9:11].sum() > 0.99)
    assert np.all(p[9:11] > 0.047)


@pytest.mark.parametrize('reference_frequency', [None, 300.064e6, 299.872e6])
@pytest.mark.parametrize('spf,atol', [(None, 1e-2), (50000, 1e-4)])
def test_disperse_roundtrip(bt, reference_frequency, spf, atol):
    gp = giant_pulse(bt)
    gp.seek(gp.start_time + 0.5)
    gp.seek(-1024, 1)
    want = gp.read(2048)
    disperse = bt.Disperse(gp, DM, reference_frequency=reference_frequency,
                           samples_per_frame=spf)
    dedisperse = bt.Dedisperse(disperse, DM,
                               reference_frequency=reference_frequency,
                               samples_per_frame=spf)
    dedisperse.seek(gp.start_time + GP_SAMPLE / RATE)
    dedisperse.seek(-1024, 1)
    got = dedisperse.read(2048)
    assert np.all(np.abs(got - want) < atol)
    assert dedisperse.dm == DM and disperse.dm == DM


def test_disperse_roundtrip_to_mean_frequency(bt):
    """Dedispersing to the mean frequency leaves a net time shift."""
    gp = giant_pulse(bt)
    dm = bt.DispersionMeasure(DM)
    disperse = bt.Disperse(gp, dm, reference_frequency=300.064e6,
                           samples_per_frame=50000)
    delay = float(dm.time_delay(300e6, disperse.reference_frequency))
    dedisperse = bt.Dedisperse(disperse, dm, samples_per_frame=50000)
    dedisperse.seek(gp.start_time + GP_SAMPLE / RATE + delay)
    dedisperse.seek(-1024, 1)
    dd_gp = dedisperse.read(2048)
    p = np.abs(dd_gp) ** 2
    assert np.all(p[1023:1026].sum(0) > 0.9)


def test_disperse_negative_dm_and_close(bt):
    gp = giant_pulse(bt)
    disperse = bt.Disperse(gp, -DM)
    disperse.seek(gp.start_time + GP_SAMPLE / RATE)
    disperse.seek(-GP_SAMPLE // 2, 1)
    around = disperse.read(GP_SAMPLE)
    p = (np.abs(around) ** 2).reshape(-1, 10, GP_SAMPLE // 10 // 20, 2).sum(2)
    assert np.all(p[:9].sum(1) < 0.01)
    assert np.all(p[11:].sum(1) < 0.01)
    assert np.all(p[9:11].sum() > 0.99)
    assert np.all(p[9:11] > 0.047)
    disperse.close()
    with pytest.raises(ValueError):
        disperse.read(1)


@pytest.mark.parametrize('reference_frequency', [None, 300.064e6, 299.872e6])
def test_reference_framing(bt, reference_frequency):
    """With ``fft_maker.set('cuda', fast_len='reference')`` frames are padded
    like the reference's numpy maker pads them, so the default framing is the
    reference's own: samples_per_frame 19324 or 19200 (its known answer,
    tests/test_dispersion.py:64-69), and the samples agree with the oracle
    on that framing."""
    import bbt_oracle as orc
    from test_kernels import assert_voltage
    from baseband_tasks_b200 import _cabi
    from baseband_tasks_b200.fourier import fft_maker
    from baseband_tasks_b200.fourier.cuda import smooth_fast_len
    if _cabi._DEVICE.type == 'cpu' and reference_frequency is not None:
        pytest.skip('one framing is enough on host threads (a minute each)')
    for n in (1, 7, 8, 130, 4095, 19324 + 6400, 24000, 100003):
        assert smooth_fast_len(n) == orc.next_fast_len(n)
    assert smooth_fast_len(130) == 135          # tests/test_base.py:522-537
    gp = giant_pulse(bt)
    with fft_maker.set('cuda', fast_len='reference'):
        dd = bt.Disperse(gp, DM, reference_frequency=reference_frequency)
    assert dd.samples_per_frame in (19324, 19200)
    n = dd._ih_samples_per_frame
    assert n & (n - 1) and n == orc.next_fast_len(n)
    x = np.zeros((164000, 2), 'c8')
    x[GP_SAMPLE] = 1.
    fref = None if reference_frequency is None else reference_frequency / 1e6
    op = orc.DispersePlan(DM, 300., np.array([1, -1]), RATE / 1e6, True,
                          164000, 1000, (2,), reference_frequency_mhz=fref)
    assert (op.N, op.samples_per_frame) == (n, dd.samples_per_frame)
    want = orc.disperse(x, op)
    got = dd.read()
    assert got.shape == want.shape
    # A single unit sample: judged against the peak of the smeared pulse.
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    # Real-valued stream on the same framing (rfft / irfft route).
    rng = np.random.default_rng(12)
    r = rng.normal(size=(60000,)).astype('f4')
    src = bt.ArrayStream(r, start_time(bt), RATE, samples_per_frame=1000,
                         frequency=300e6, sideband=1)
    with fft_maker.set('cuda', fast_len='reference'):
        dr = bt.Disperse(src, DM / 8)
    opr = orc.DispersePlan(DM / 8, 300., 1, RATE / 1e6, False, 60000, 1000, ())
    assert (opr.N, opr.samples_per_frame) == (dr._ih_samples_per_frame,
                                              dr.samples_per_frame)
    assert_voltage(dr.read(), orc.disperse(r, opr).astype('f4'))
