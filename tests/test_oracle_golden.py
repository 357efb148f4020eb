"""Pin the numpy oracle against the reference's own known-answer tests.

Each test names the reference test (under /root/reference/baseband_tasks) whose
assertions it replays on the same synthetic inputs.
"""
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

import bbt_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


# ---- tests/test_dm.py:33-73
def test_dm_calculation():
    assert orc.dm_time_delay(1., 1.) == 1. / 2.41e-4
    assert orc.dm_phase_delay(1., 1.) == 1. / 2.41e-4 * 1e6
    freqs = np.array([369.66462, 373.56482, 319.541562, 297.2516, 321.053234])
    ref = 321.582761
    dm = 29.1168
    d = dm / 2.41e-4
    assert_allclose(orc.dm_time_delay(dm, freqs, ref),
                    d * (1 / freqs**2 - 1 / ref**2), rtol=1e-13)
    assert_allclose(orc.dm_time_delay(dm, freqs), d / freqs**2, rtol=1e-13)
    assert_allclose(orc.dm_phase_delay(dm, freqs, ref),
                    d * 1e6 * freqs * (1 / ref - 1 / freqs)**2, rtol=1e-13)
    assert_allclose(orc.dm_phase_delay(dm, freqs), d * 1e6 / freqs,
                    rtol=1e-13)
    assert_allclose(orc.dm_phase_factor(dm, freqs, ref),
                    np.exp(2j * np.pi * d * 1e6 * freqs
                           * (1 / ref - 1 / freqs)**2), rtol=1e-6)


# ---- tests/test_base.py:522-537 and fourier/numpy.py:99-126
@pytest.mark.parametrize('n,expected', [
    (1, 1), (7, 7), (8, 8), (11, 12), (13, 14), (130, 135), (133, 135),
    (20000, 20000), (20001, 20160), (1025, 1029), (83349, 83349),
    (25725, 25725)])
def test_next_fast_len(n, expected):
    assert orc.next_fast_len(n) == expected


def test_next_fast_len_brute():
    smooth = sorted(2**a * 3**b * 5**c * 7**d
                    for a in range(12) for b in range(8)
                    for c in range(6) for d in range(5))
    for n in list(range(1, 600)) + [4097, 19324 + 1000, 99999]:
        if n <= 7:
            assert orc.next_fast_len(n) == n
        else:
            assert orc.next_fast_len(n) == next(s for s in smooth if s >= n)


# ---- tests/test_dispersion.py:14-124
REFERENCE_FREQUENCIES = (None, 300., 300.0123456789, 300.064, 299.936,
                         300.128, 300.123456789, 299.872)
GP_SAMPLE = 64000
GP_SHAPE = (164000, 2)
GP_DM = 1000. * 0.05 / 0.039342251
GP_RATE_MHZ = 0.128
GP_SIDEBAND = np.array((1, -1))


def giant_pulse(dtype=np.complex64):
    data = np.zeros(GP_SHAPE, dtype)
    data[GP_SAMPLE] = 1.
    return data


def gp_plan(ref, dm=GP_DM, spf=None, n_in=GP_SHAPE[0], ih_spf=1000):
    return orc.DispersePlan(dm, 300., GP_SIDEBAND, GP_RATE_MHZ, True,
                            n_in, ih_spf, sample_shape=(2,),
                            reference_frequency_mhz=ref,
                            samples_per_frame=spf)


def test_time_delay():
    delay = orc.dm_time_delay(GP_DM, 300. - GP_RATE_MHZ / 2,
                              300. + GP_RATE_MHZ / 2)
    assert abs(delay - 0.05) < 1e-9


@pytest.mark.parametrize('ref', REFERENCE_FREQUENCIES)
def test_disperse_samples_per_frame(ref):
    plan = gp_plan(ref)
    assert plan.samples_per_frame in (19324, 19200)


def test_disperse_sample_offsets():
    # SURVEY.md section 9.4 (survey-time replay).
    got = [gp_plan(ref).sample_offset for ref in REFERENCE_FREQUENCIES]
    assert got == [0, 0, 0, 0, 0, 3196, 2970, -3203]


@pytest.mark.parametrize('ref', REFERENCE_FREQUENCIES)
def test_disperse_time_offset(ref):
    plan = gp_plan(ref)
    expected = orc.dm_time_delay(GP_DM, 299.936, plan.reference_frequency_mhz)
    assert abs(plan.start_offset - expected) < 1. / (GP_RATE_MHZ * 1e6)


@pytest.mark.parametrize('ref', REFERENCE_FREQUENCIES)
def test_disperse(ref):
    plan = gp_plan(ref)
    out = orc.disperse(giant_pulse(), plan)
    assert out.shape == (GP_SHAPE[0] - plan.pad_start - plan.pad_end, 2)
    rate = GP_RATE_MHZ * 1e6
    t_gp = GP_SAMPLE / rate + orc.dm_time_delay(
        GP_DM, 300., plan.reference_frequency_mhz)
    offset = int(np.round((t_gp - plan.start_offset) * rate))
    offset -= GP_SAMPLE // 2
    around_gp = out[offset:offset + GP_SAMPLE]
    p = (np.abs(around_gp) ** 2).reshape(
        -1, 10, GP_SAMPLE // 20 // 10, 2).sum(2)
    assert np.all(p[:9].sum(1) < 0.005)
    assert np.all(p[11:].sum(1) < 0.005)
    assert np.all(p[9:11].sum() > 0.99)
    assert np.all(p[9:11] > 0.047)


@pytest.mark.parametrize('ref', REFERENCE_FREQUENCIES[:4])
@pytest.mark.parametrize('spf,atol', [(None, 1e-2), (50000, 1e-4)])
def test_disperse_roundtrip1(ref, spf, atol):
    gp = giant_pulse()
    plan = gp_plan(ref, spf=spf)
    dispersed = orc.disperse(gp, plan)
    plan2 = gp_plan(ref, dm=-GP_DM, spf=spf, n_in=dispersed.shape[0],
                    ih_spf=plan.samples_per_frame)
    dedispersed = orc.disperse(dispersed, plan2)
    rate = GP_RATE_MHZ * 1e6
    total_offset = plan.start_offset + plan2.start_offset
    pos = int(np.round(GP_SAMPLE - total_offset * rate))
    gp_dd = dedispersed[pos - 1024:pos + 1024]
    assert np.all(np.abs(gp_dd - gp[GP_SAMPLE - 1024:GP_SAMPLE + 1024])
                  < atol)


def test_disperse_negative_dm():
    plan = gp_plan(None, dm=-GP_DM)
    out = orc.disperse(giant_pulse(), plan)
    rate = GP_RATE_MHZ * 1e6
    offset = int(np.round((GP_SAMPLE / rate - plan.start_offset) * rate))
    offset -= GP_SAMPLE // 2
    p = (np.abs(out[offset:offset + GP_SAMPLE]) ** 2).reshape(
        -1, 10, GP_SAMPLE // 10 // 20, 2).sum(2)
    assert np.all(p[:9].sum(1) < 0.01)
    assert np.all(p[11:].sum(1) < 0.01)
    assert np.all(p[9:11].sum() > 0.99)
    assert np.all(p[9:11] > 0.047)


# ---- tests/test_pfb.py:26-102
def test_sinc_hamming_guppi():
    a = np.loadtxt(os.path.join(GOLDEN, 'guppi_pfb_coeffs.txt'))
    guppi = a.reshape(8, -1).T.reshape(12, 64)
    assert_allclose(orc.sinc_hamming(12, 64, sinc_scale=0.95), guppi)


@pytest.mark.parametrize('offset', (0, 1000))
@pytest.mark.parametrize('dtype', ('f8', 'c16'))
def test_pfb_understanding(offset, dtype):
    h = orc.sinc_hamming(4, 2048)
    n_in = 2500 * 2048
    # Reference framing: pad = 3*2048, ih spf 128 -> N = 4*pad.
    big_n, spf, n_out = orc.pfb_framing(n_in, 128, h)
    assert (big_n, spf) == (4 * 3 * 2048, 3 * 3 * 2048)
    # Read a region covering output spectra [offset, offset+2) only.
    first_frame = offset // (spf // 2048)
    x0 = first_frame * spf
    x = orc.noise_stream(12345, big_n + spf, 128, (), dtype, start=x0)
    d = orc.noise_stream(12345, 5 * 2048, 128, (), dtype,
                         start=offset * 2048).reshape(-1, 2048)
    rfft = np.fft.rfft if dtype == 'f8' else np.fft.fft
    ft2 = rfft((h * d[:4]).sum(0))
    ft1 = rfft((h * d[:4]).ravel())[::4]
    assert_allclose(ft1, ft2)
    o = offset - first_frame * (spf // 2048)
    for fourier in (False, True):
        ft = orc.pfb(x, h, ih_samples_per_frame=128, fourier=fourier)
        assert_allclose(ft[o], ft2)
        assert_allclose(ft[o + 1], rfft((h * d[1:]).sum(0)))


# ---- tests/test_integration.py
def fake_pulsar():
    idx = np.arange(16000)
    data = np.where(idx % 125 == 0, 10., 0.125)
    return np.repeat(data[:, None], 2, axis=1)


def test_integrate_all():
    raw_power = fake_pulsar() ** 2
    plan = orc.IntegratePlan(16000, 1e4)
    assert plan.n_out == 1
    data, count = orc.integrate(raw_power, plan.offsets([0, 1]), 200)
    assert np.all(count == 16000)
    assert np.allclose(data / count, raw_power.mean(0))


@pytest.mark.parametrize('seek', [121, -10])
@pytest.mark.parametrize('n', (1, 3))
def test_integrate_n(n, seek):
    raw_power = fake_pulsar() ** 2
    n_sample = 16000 // n
    seek = seek if seek > 0 else n_sample + seek
    ref = raw_power[seek * n:(seek + 10) * n].reshape(-1, n, 2).sum(1)
    plan = orc.IntegratePlan(16000, 1e4, n)
    assert plan.n_out == n_sample
    offsets = plan.offsets(np.arange(seek, seek + 11))
    data, count = orc.integrate(raw_power, offsets, 200)
    assert np.allclose(data, ref)
    assert np.all(count == n)
    # Same via a time step and start time (test_integration.py:137-148).
    plan = orc.IntegratePlan(16000, 1e4, n / 1e4, start=seek * n / 1e4)
    data, count = orc.integrate(raw_power, plan.offsets(np.arange(11)), 200)
    assert np.allclose(data, ref)
    assert np.all(count == n)


def test_integrate_time_non_integer_ratio():
    expected_count = [2, 3, 2, 2, 2, 3, 2, 2]
    raw_power = fake_pulsar() ** 2
    step = 2.26 / 1e4
    ref = np.add.reduceat(raw_power[:18],
                          np.add.accumulate([0] + expected_count[:-1]))
    plan = orc.IntegratePlan(16000, 1e4, step)
    data, count = orc.integrate(raw_power, plan.offsets(np.arange(9)), 200)
    assert np.allclose(data, ref)
    assert np.all(count.ravel() == expected_count)
    for k in (1, 3):
        plan2 = orc.IntegratePlan(16000, 1e4, step, start=k * step)
        d2, c2 = orc.integrate(raw_power, plan2.offsets(np.arange(9 - k)),
                               200)
        assert np.all(d2 == data[k:]) and np.all(c2 == count[k:])


def test_integrate_errors():
    with pytest.raises(ValueError):
        orc.IntegratePlan(16000, 1e4, start=-1.)
    with pytest.raises(ValueError):
        orc.IntegratePlan(16000, 1e4, start=3.)
    with pytest.raises(AssertionError):
        orc.IntegratePlan(16000, 1e4, step=3600.)


def fake_phase(i):
    f0 = 1.0 / (125 / 1e4)
    return f0 * (np.asarray(i) / 1e4)


def test_fold_step_shorter_than_period():
    raw = fake_pulsar()
    plan = orc.IntegratePlan(16000, 1e4, 10e-3)
    counts = []
    datas = []
    for k in range(3):
        d, c = orc.fold(raw, plan.offsets([k, k + 1]), 50, fake_phase)
        counts.append(c[0, :, 0])
        datas.append(d[0])
    fr_count = np.array(counts)
    fr_data = np.array(datas)
    assert np.all(fr_count.sum(1) == 100)
    assert np.all((fr_count[0, :40] == 3) | (fr_count[0, :40] == 2))
    assert np.all(fr_count[0, 41:] == 0)
    assert np.all(fr_count[1, :30] != 0)
    assert np.all(fr_count[1, 40:] != 0)
    assert np.all(fr_count[1, 31:39] == 0)
    assert np.all(fr_data[:, (0, 1, -1)].sum(1) > 10)
    assert np.all(fr_data[:, 2:49] <= 0.125 * 3)


def test_fold_whole_file():
    raw = fake_pulsar()
    phase = fake_phase(np.arange(16000))
    i_phase = ((phase % 1.) * 50).astype(int)
    expected = np.bincount(i_phase, raw[:, 0]) / np.bincount(i_phase)
    plan = orc.IntegratePlan(16000, 1e4)
    d, c = orc.fold(raw, plan.offsets([0, 1]), 50, fake_phase)
    fr = d / c
    assert np.all(fr[:, 2:-1] == 0.125)
    assert np.all(fr[0, :, 0] == expected)
    assert c.sum() == 16000


def test_fold_multi_bin_quirk():
    # integration.py:386: with several time bins per output frame, a sample
    # exactly on a bin edge goes to the earlier bin (side='left').
    raw = fake_pulsar()
    plan = orc.IntegratePlan(16000, 1e4, 26e-3)
    offsets = plan.offsets(np.arange(11))
    d, c = orc.fold(raw, offsets, 50, fake_phase)
    widths = np.diff(offsets)
    got = c[:, :, 0].sum(1)
    assert got[0] == widths[0] + 1 and got[-1] == widths[-1] - 1
    assert np.all(got[1:-1] == widths[1:-1])


# ---- tests/test_generators.py:253-316
def test_noise_reproducible():
    shape = (4, 2)
    full = orc.noise_stream(1234567, 10000, 10, shape, 'c16')
    assert full.shape == (10000, 4, 2) and full.dtype == np.dtype('c16')
    assert np.all(orc.noise_stream(1234567, 3, 10, shape, 'c16', start=9)
                  == full[9:12])
    assert np.all(orc.noise_stream(1234567, 1000, 10, shape, 'c16',
                                   start=9000) == full[9000:])
    assert abs(full.mean()) < 10. / full.size ** 0.5
    assert abs(full.std() - np.sqrt(2.)) < 14. / full.size ** 0.5
    one = orc.noise_stream(1234567, 5, 1, shape, 'c16')
    assert not np.any(one[0] == one[3]) and not np.any(one[2] == one[4])


# ---- fourier/tests/test_fourier.py:77-166 (the numpy maker boundary)
def test_fft_conventions():
    n = 7919
    x = (np.exp(2j * np.pi * 720 * np.arange(n) / n)).astype('c16')
    ft = orc.fft(x)
    assert_allclose(ft, np.fft.fft(x), atol=1e-13, rtol=1e-6)
    assert np.argmax(np.abs(ft)) == 720
    assert_allclose(orc.ifft(ft, 'c16'), x, atol=1e-13, rtol=1e-6)
    xr = np.cos(2 * np.pi * 10 * np.arange(1000) / 1000)
    ftr = orc.fft(xr, ortho=True)
    assert ftr.shape == (501,) and ftr.dtype == np.dtype('c16')
    assert_allclose(orc.ifft(ftr, 'f8', n=1000, ortho=True), xr, atol=1e-13)
    assert_allclose((np.abs(orc.fft(x, ortho=True))**2).sum(),
                    (np.abs(x)**2).sum())
    x32 = x.astype('c8')
    assert orc.fft(x32).dtype == np.dtype('c8')
    assert orc.freq_dtype('f4') == np.dtype('c8')
    f = orc.fft_frequency(8, 8., False, trailing=2)
    assert f.shape == (8, 1, 1) and f[5, 0, 0] == -3.
