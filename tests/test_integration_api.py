"""Integrate/Fold API semantics after the reference's
tests/test_integration.py:60-330 (same fake pulsar: 16000 x 2 samples at
10 kHz, a pulse every 125 samples)."""
import numpy as np
import pytest

from test_tasks import bt, start_time, fake_pulsar  # noqa: F401  (fixture)


def test_integrate_all_and_part(bt):
    data, src = fake_pulsar(bt)
    power = data ** 2
    st = bt.Square(src)
    ip = bt.Integrate(st)                        # everything
    assert ip.shape == (1, 2)
    assert ip.start_time == src.start_time
    out = ip.read()
    np.testing.assert_allclose(out, power.mean(0, keepdims=True), rtol=1e-5)
    # not averaged: structured samples with data and count
    ip = bt.Integrate(st, average=False)
    res = ip.read()
    assert res.dtype.names == ('data', 'count')
    assert res['data'].dtype == st.dtype and res['data'].shape == (1, 2)
    np.testing.assert_allclose(res['data'], power.sum(0, keepdims=True),
                               rtol=1e-5)
    assert np.all(res['count'] == 16000)
    assert 'average=False' in repr(ip)
    # a part, from a start offset to the end
    ip = bt.Integrate(st, start=1000, average=False)
    assert abs((ip.start_time - src.start_time) - 0.1) < 1e-9
    res = ip.read()
    np.testing.assert_allclose(res['data'], power[1000:].sum(0, keepdims=True),
                               rtol=1e-5)
    assert np.all(res['count'] == 15000)


@pytest.mark.parametrize('n', [1, 3])
@pytest.mark.parametrize('spf', [1, 4, 10])
def test_integrate_n(bt, n, spf):
    data, src = fake_pulsar(bt)
    power = data ** 2
    st = bt.Square(src)
    n_sample = 16000 // n
    ip = bt.Integrate(st, n, average=False, samples_per_frame=spf)
    assert ip.shape[0] == n_sample
    assert ip.sample_rate == src.sample_rate / n
    assert f'step={n}' in repr(ip)
    for seek in (121, n_sample - 10):
        ip.seek(seek)
        assert abs((ip.time - src.start_time) - seek * n / 1e4) < 1e-9
        res = ip.read(10)
        assert ip.tell() == seek + 10
        want = power[seek * n:(seek + 10) * n].reshape(-1, n, 2).sum(1)
        np.testing.assert_allclose(res['data'], want, rtol=1e-5)
        assert np.all(res['count'] == n)
    # the same through a time step, a start offset and a start time
    want = power[151 * n:161 * n].reshape(-1, n, 2).sum(1)
    ip = bt.Integrate(st, n / 1e4, average=False, samples_per_frame=spf)
    assert ip.sample_rate == src.sample_rate / n
    ip.seek(151)
    np.testing.assert_allclose(ip.read(10)['data'], want, rtol=1e-5)
    ip = bt.Integrate(st, n / 1e4, start=151 * n, average=False,
                      samples_per_frame=spf)
    assert abs((ip.start_time - src.start_time) - 151 * n / 1e4) < 1e-9
    np.testing.assert_allclose(ip.read(10)['data'], want, rtol=1e-5)
    st.seek(151 * n)
    ip = bt.Integrate(st, n / 1e4, start=st.time, average=False,
                      samples_per_frame=spf)
    res = ip.read(10)
    assert ip.tell() == 10
    np.testing.assert_allclose(res['data'], want, rtol=1e-5)
    assert np.all(res['count'] == n)


@pytest.mark.parametrize('spf', [1, 4, 10])
def test_integrate_non_integer_ratio(bt, spf):
    """2.26 samples per bin: counts 2, 3, 2, 2, 2, 3, 2, 2."""
    data, src = fake_pulsar(bt)
    power = data ** 2
    expected = [2, 3, 2, 2, 2, 3, 2, 2]
    step = 2.26 / 1e4
    st = bt.Square(src)
    ip = bt.Integrate(st, step, average=False, samples_per_frame=spf)
    assert abs(ip.sample_rate - 1. / step) < 1e-6
    res = ip.read(8)
    want = np.add.reduceat(power[:18],
                           np.add.accumulate([0] + expected[:-1]))
    np.testing.assert_allclose(res['data'], want, rtol=1e-5)
    assert np.all(res['count'].reshape(8, -1).T == expected)
    for k, m in ((1, 7), (3, 5)):
        t = src.start_time + k * step
        ip2 = bt.Integrate(st, step, start=t, average=False,
                           samples_per_frame=spf)
        assert abs(ip2.start_time - t) < 1e-9
        res2 = ip2.read(m)
        np.testing.assert_allclose(res2['data'], res['data'][k:], rtol=1e-6)
        np.testing.assert_array_equal(res2['count'], res['count'][k:])


def test_integrate_times_wrong(bt):
    data, src = fake_pulsar(bt)
    with pytest.raises(ValueError):
        bt.Integrate(src, start=src.start_time - 1.)
    with pytest.raises(ValueError):
        bt.Integrate(src, start=src.start_time + 3.)
    with pytest.raises(AssertionError):
        bt.Integrate(src, step=3600.)


def test_fold_steps(bt):
    """test_integration.py:262-330: step shorter and longer than the period."""
    data, src = fake_pulsar(bt)
    n_phase = 50

    def phase(t):
        return (t - src.start_time) * 80.

    fh = bt.Fold(src, n_phase, phase, 10e-3, samples_per_frame=1,
                 average=False)
    fr = fh.read(3)
    cnt = fr['count'].reshape(3, n_phase, -1)[..., 0]
    dat = fr['data']
    assert np.all(cnt.sum(1) == 100)
    assert np.all((cnt[0, :40] == 3) | (cnt[0, :40] == 2))
    assert np.all(cnt[0, 41:] == 0)
    assert np.all(cnt[1, :30] != 0) and np.all(cnt[1, 40:] != 0)
    assert np.all(cnt[1, 31:39] == 0)
    assert np.all(dat[:, (0, 1, -1)].sum(1) > 10)
    assert np.all(dat[:, 2:49] <= 0.125 * 3)
    fh = bt.Fold(src, n_phase, phase, 30e-3, samples_per_frame=1,
                 average=False)
    fr = fh.read(10)
    cnt = fr['count'].reshape(10, n_phase, -1)[..., 0]
    assert np.all(cnt.sum(1) == 300)
    dat = fr['data']
    on = (0, 1, -1)
    pulse_power = dat[:, on].sum(1) / cnt[:, on].sum(1)[:, None]
    assert np.all(np.abs(pulse_power - 10. / 7.5 - 0.125) < 0.5)
    np.testing.assert_allclose(dat[:, 2:-1] / cnt[:, 2:-1, None], 0.125,
                               rtol=1e-5)
    # with an offset start time
    fh2 = bt.Fold(src, n_phase, phase, 30e-3, start=src.start_time + 30e-3,
                  samples_per_frame=1, average=False)
    fr2 = fh2.read(9)
    np.testing.assert_array_equal(fr2['count'], fr['count'][1:])
    np.testing.assert_allclose(fr2['data'], fr['data'][1:], rtol=1e-6)
    # averaged (test_folding_with_averaging, test_non_integer_sample_rate_ratio)
    fa = bt.Fold(src, n_phase, phase, 26e-3, samples_per_frame=20)
    out = fa.read(10)
    assert out.shape == (10, n_phase, 2)
    np.testing.assert_allclose(out[:, 2:-1], 0.125, rtol=1e-6)
    fb = bt.Fold(src, n_phase, phase, 1. / 3.)
    out = fb.read()
    assert out.shape[0] == 4
    np.testing.assert_allclose(out[:, 2:-1], 0.125, rtol=1e-6)


def test_fold_whole_and_part(bt):
    """test_integration.py:355-392: fold everything, or from a start time."""
    data, src = fake_pulsar(bt)
    n_phase = 50

    def phase(t):
        return (t - src.start_time) * 80.

    i = np.arange(16000)
    i_phase = ((i / 1e4 * 80. * n_phase) % n_phase).astype(int)
    expected = (np.bincount(i_phase, data[:, 0].astype('f8'))
                / np.bincount(i_phase))
    fh = bt.Fold(src, n_phase, phase)
    assert abs(fh.stop_time - src.stop_time) < 1e-9
    fr = fh.read(1)
    np.testing.assert_allclose(fr[:, 2:-1], 0.125, rtol=1e-6)
    np.testing.assert_allclose(fr[0, :, 0], expected, rtol=1e-5)
    start = src.start_time + 1.
    fh = bt.Fold(src, n_phase, phase, average=False, start=start)
    assert abs(fh.start_time - start) < 1e-9
    assert abs(fh.stop_time - src.stop_time) < 1e-9
    fr = fh.read(1)
    assert np.all(fr['count'].sum((0, 1)) == 6000)
    average = fr['data'][0] / fr['count'][0].reshape(n_phase, -1)
    np.testing.assert_allclose(average[2:-1], 0.125, rtol=1e-6)
    with pytest.raises(ValueError):
        bt.Fold(src, 8, phase, start=src.start_time - 1.)
    with pytest.raises(ValueError):
        bt.Fold(src, 8, phase, start=src.start_time + 3.)
    with pytest.raises(AssertionError):
        bt.Fold(src, 8, phase, step=3600.)


@pytest.mark.parametrize('spf', [1, 160])
def test_integrate_phase_steps(bt, spf):
    """test_integration.py:406-425: 25 phase steps per cycle, 5 samples each."""
    data, src = fake_pulsar(bt)

    def phase(t):
        return (t - src.start_time) * 80.

    ref = data.reshape(-1, 5, 2).mean(1)
    fh = bt.Integrate(src, 1. / 25, phase, samples_per_frame=spf)
    assert fh.start_time == src.start_time
    assert abs(fh.stop_time - src.stop_time) < 1e-9
    assert fh.samples_per_frame == spf
    np.testing.assert_allclose(fh.read(20), ref[:20], rtol=1e-6)
    fh.seek(250)
    np.testing.assert_allclose(fh.read(75), ref[250:325], rtol=1e-6)
    if spf > 1:
        np.testing.assert_allclose(fh.read(), ref[325:], rtol=1e-6)


@pytest.mark.parametrize('spf', [1, 16])
def test_pulse_stack_basics(bt, spf):
    """test_integration.py:432-470."""
    data, src = fake_pulsar(bt)

    def phase(t):
        return (t - src.start_time) * 80.

    ref = data.reshape(-1, 25, 5, 2).mean(2)
    fh = bt.PulseStack(src, 25, phase, samples_per_frame=spf)
    assert fh.start_time == src.start_time
    assert abs(fh.stop_time - src.stop_time) < 1e-9
    assert fh.samples_per_frame == spf
    fh.seek(5)
    assert abs((fh.time - src.start_time) - 5 / 80.) < 1e-9
    fh.seek(0)
    np.testing.assert_allclose(fh.read(2), ref[:2], rtol=1e-6)
    fh.seek(10)
    np.testing.assert_allclose(fh.read(3), ref[10:13], rtol=1e-6)
    np.testing.assert_allclose(fh.read(), ref[13:], rtol=1e-6)
    # a slice of the input
    ref2 = data[-360:-110].reshape(-1, 25, 5, 2).mean(2)
    fh = bt.PulseStack(src[-360:-10], 25, phase, samples_per_frame=spf)
    assert fh.shape == ref2.shape
    np.testing.assert_allclose(fh.read(), ref2, rtol=1e-6)


def test_pulse_stack_offset_slice_integrate(bt):
    """test_integration.py:465-520."""
    data, src = fake_pulsar(bt)

    def phase(t):
        return (t - src.start_time) * 80.

    ref = data[124:-1].reshape(-1, 25, 5, 2).mean(2)
    fh = bt.PulseStack(src, 25, phase, start=124)
    assert abs((fh.start_time - src.start_time) - 124 / 1e4) < 1e-9
    assert abs((fh.stop_time - src.stop_time) + 1 / 1e4) < 1e-9
    np.testing.assert_allclose(fh.read(2), ref[:2], rtol=1e-6)
    fh.seek(10)
    assert abs((fh.time - src.start_time) - 124 / 1e4 - 10 / 80.) < 1e-9
    np.testing.assert_allclose(fh.read(), ref[10:], rtol=1e-6)
    for item in (slice(10, 100), slice(-10, None), slice(None, 10),
                 slice(None), (slice(10, 100), 0)):
        sliced = fh[item]
        first = item[0] if isinstance(item, tuple) else item
        start, stop, _ = first.indices(fh.shape[0])
        t_start = 124 / 1e4 + start / 80.
        assert abs((sliced.start_time - src.start_time) - t_start) < 1e-9
        assert abs((sliced.stop_time - src.start_time)
                   - 124 / 1e4 - stop / 80.) < 1e-9
        sliced.seek(5)
        assert abs((sliced.time - sliced.start_time) - 5 / 80.) < 1e-9
        sliced.seek(0)
        want = ref[item]
        assert sliced.shape == want.shape
        np.testing.assert_allclose(np.asarray(sliced.read()), want, rtol=1e-6)
    # integrating a stack
    fh = bt.PulseStack(src, 25, phase)
    stack = fh.read(3)
    ih = bt.Integrate(fh, 3)
    np.testing.assert_allclose(ih.read(1)[0], stack.mean(0), rtol=1e-6)
    assert ih.tell() == 1
    assert abs((ih.time - src.start_time) - 3 / 80.) < 1e-9
