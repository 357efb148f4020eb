"""Task API semantics after the reference's tests/test_base.py and
tests/test_generators.py, on synthetic streams (the reference uses a VDIF
sample file read with `baseband`, which is not available here)."""
import operator

import numpy as np
import pytest

from test_tasks import bt, start_time  # noqa: F401  (fixture)
from test_kernels import cnoise


def make_stream(bt, n=4000, shape=(8,), rate=32e6, spf=400, **kwargs):
    x = cnoise(np.random.default_rng(7), (n,) + shape).real.copy()
    return x, bt.ArrayStream(x, start_time(bt), rate, samples_per_frame=spf,
                             **kwargs)


def test_set_attribute(bt):
    """test_base.py:75-165."""
    x, fh = make_stream(bt)
    frequency = 311.25e6 + (np.arange(8.) // 2) * 16e6
    sideband = np.tile([-1, +1], 4)
    sa = bt.SetAttribute(fh, frequency=frequency, sideband=sideband)
    assert np.all(sa.frequency == frequency)
    assert np.all(sa.sideband == sideband)
    for attr in ('sample_rate', 'samples_per_frame', 'shape', 'dtype'):
        assert getattr(sa, attr) == getattr(fh, attr)
    assert sa.start_time == fh.start_time
    frequency[...] = 0          # no aliasing of the caller's arrays
    sideband[...] = 0
    assert np.all(sa.frequency != 0) and np.all(np.abs(sa.sideband) == 1)
    np.testing.assert_array_equal(sa.read(), x)
    with pytest.raises(AttributeError):
        sa.polarization
    # start time
    t1 = fh.start_time + 0.1
    sb = bt.SetAttribute(fh, start_time=t1)
    assert sb.start_time == t1
    sb.seek(10)
    np.testing.assert_array_equal(sb.read(10), x[10:20])
    sb.seek(10 / 32e6)          # a time offset in seconds
    np.testing.assert_array_equal(sb.read(10), x[10:20])
    sb.seek(sb.start_time + 10 / 32e6)
    np.testing.assert_array_equal(sb.read(10), x[10:20])
    for attr in ('frequency', 'sideband', 'polarization'):
        with pytest.raises(AttributeError):
            getattr(sb, attr)
    # frequency and sideband go together; unknown attributes fail
    with pytest.raises(ValueError):
        bt.SetAttribute(fh, frequency=frequency)
    with pytest.raises(ValueError):
        bt.SetAttribute(fh, sideband=sideband)
    with pytest.raises(TypeError):
        bt.SetAttribute(fh, bogus=1)
    # samples_per_frame and dtype overrides
    sc = bt.SetAttribute(fh, samples_per_frame=57, dtype='f8')
    assert sc.samples_per_frame == 57 and sc.dtype == np.dtype('f8')
    sc.seek(60)
    got = sc.read(100)
    assert got.dtype == np.dtype('f8')
    np.testing.assert_array_equal(got, x[60:160])


def test_task_base_subclass(bt):
    """test_base.py:16-33,180-260: a reshaping TaskBase subclass."""
    class ReshapeTime(bt.TaskBase):
        def __init__(self, ih, n, samples_per_frame=1, **kwargs):
            self._n = n = operator.index(n)
            super().__init__(ih, shape=(-1, n) + ih.shape[1:],
                             sample_rate=ih.sample_rate / n,
                             samples_per_frame=samples_per_frame, **kwargs)

        def task(self, data):
            return data.reshape((-1,) + self.sample_shape)

    x, fh = make_stream(bt)
    for n, spf in ((16, 1), (16, 7), (125, 3)):
        rt = ReshapeTime(fh, n, samples_per_frame=spf)
        n_out = (4000 // (n * spf)) * spf
        assert rt.shape == (n_out, n, 8)
        assert rt.sample_rate == 32e6 / n
        assert rt.samples_per_frame == spf
        assert abs((rt.stop_time - rt.start_time) - n_out * n / 32e6) < 1e-12
        want = x[:n_out * n].reshape(n_out, n, 8)
        np.testing.assert_array_equal(np.asarray(rt.read()), want)
        rt.seek(-3, 2)
        assert rt.tell() == n_out - 3
        np.testing.assert_array_equal(np.asarray(rt.read()), want[-3:])
        with pytest.raises(EOFError):
            rt.read(1)
        rt.seek(1)
        assert abs((rt.time - rt.start_time) - n / 32e6) < 1e-12
        with pytest.raises(ValueError):
            rt.seek(0, 5)
        rt.close()
        with pytest.raises(ValueError):
            rt.read(1)
    # frequency/sideband propagate, and can be overridden
    f = 300e6 + 16e6 * np.arange(8)
    fs = bt.SetAttribute(fh, frequency=f, sideband=1)
    rt = ReshapeTime(fs, 16)
    np.testing.assert_array_equal(rt.frequency, f)
    rt2 = ReshapeTime(fs, 16, frequency=f + 1., sideband=-1)
    np.testing.assert_array_equal(rt2.frequency, f + 1.)
    assert np.all(rt2.sideband == -1)


def test_function_and_method_tasks(bt):
    """test_base.py:262-330: `Task` with a function or a method."""
    x, fh = make_stream(bt)

    def zero_every_8th_sample(data):
        data = np.array(data)
        data[::8] = 0.
        return data

    def zero_every_8th_complex(fh, data):
        data = np.array(data)
        data[::8] = 0.
        return data

    for task, spf in ((zero_every_8th_sample, 8), (zero_every_8th_complex, 16)):
        ft = bt.Task(fh, task, samples_per_frame=spf)
        assert ft.shape == fh.shape and ft.samples_per_frame == spf
        got = np.asarray(ft.read())
        want = x.copy()
        want[::8] = 0.
        np.testing.assert_array_equal(got, want)
        assert task.__name__ in repr(ft)

    def double_rate(data):
        return np.repeat(np.asarray(data), 2, axis=0)

    up = bt.Task(fh, double_rate, samples_per_frame=16,
                 sample_rate=fh.sample_rate * 2)
    assert up.shape[0] == 2 * fh.shape[0]
    np.testing.assert_array_equal(np.asarray(up.read(32)),
                                  np.repeat(x[:16], 2, axis=0))
    with pytest.raises(Exception):
        bt.Task(fh, lambda a, b, c: a)     # not a function of 1 or 2 arguments


def test_sample_slices_and_arrays(bt):
    """test_base.py:395-470: time and sample slicing, array conversion."""
    x, fh = make_stream(bt, frequency=300e6 + 16e6 * np.arange(8), sideband=1)
    sl = fh[100:300]
    assert sl.shape == (200, 8)
    assert abs((sl.start_time - fh.start_time) - 100 / 32e6) < 1e-12
    np.testing.assert_array_equal(np.asarray(sl.read()), x[100:300])
    sl2 = fh[10:, 3]
    assert sl2.shape == (3990,)
    np.testing.assert_array_equal(np.asarray(sl2.read(5)), x[10:15, 3])
    assert np.all(sl2.frequency == 300e6 + 48e6)
    sl3 = fh[:, 2:6]
    assert sl3.shape == (4000, 4)
    np.testing.assert_array_equal(sl3.frequency,
                                  300e6 + 16e6 * np.arange(2, 6))
    with pytest.raises(IndexError):
        fh[5000:]
    a = np.asanyarray(fh[:50])
    np.testing.assert_array_equal(a, x[:50])


def test_generators(bt):
    """test_generators.py:20-250."""
    t0 = start_time(bt)

    def alternate(sh):
        return np.full((sh.samples_per_frame,) + sh.sample_shape,
                       sh.tell() // sh.samples_per_frame % 2 * 2 - 1, sh.dtype)

    sg = bt.StreamGenerator(alternate, (20, 4, 2), t0, 1e3,
                            samples_per_frame=2, dtype='f4',
                            frequency=np.array([[320e6], [350e6], [380e6],
                                                [410e6]]),
                            sideband=np.array([-1, 1]),
                            polarization=np.array(['X', 'Y']))
    assert sg.shape == (20, 4, 2) and sg.size == 160 and sg.ndim == 3
    assert sg.tell() == 0 and sg.tell(unit='time') == sg.time == t0
    assert abs((sg.stop_time - t0) - 0.02) < 1e-12
    data = sg.read()
    assert data.shape == (20, 4, 2)
    assert np.all(data[0::4] == -1) and np.all(data[2::4] == 1)
    assert sg.frequency.shape == (4, 1) and sg.sideband.shape == (2,)
    sg.seek(-3, 2)
    assert sg.tell() == 17
    np.testing.assert_array_equal(sg.read(), data[-3:])
    assert 'StreamGenerator' in repr(sg)
    with pytest.raises(ValueError):   # sideband must broadcast to the samples
        bt.StreamGenerator(alternate, (20, 4, 2), t0, 1e3,
                           sideband=np.ones((3, 3), dtype='i1'))
    plain = bt.StreamGenerator(alternate, (20, 4, 2), t0, 1e3)
    for attr in ('frequency', 'sideband', 'polarization'):
        with pytest.raises(AttributeError):
            getattr(plain, attr)
    with pytest.raises(EOFError):
        plain.seek(-10, 2)
        plain.read(20)
    # an empty stream filled by a task: a tone (test_generators.py:160-200)
    eh = bt.EmptyStreamGenerator((1000,), t0, 1e3, samples_per_frame=100,
                                 dtype='c8')

    def set_tone(ih, data):
        phi = 2 * np.pi * 0.01 * (ih.tell() + np.arange(len(data)))
        return (np.cos(phi) + 1j * np.sin(phi)).astype('c8')

    tone = bt.Task(eh, set_tone)
    got = np.asarray(tone.read())
    phi = 2 * np.pi * 0.01 * np.arange(1000)
    np.testing.assert_allclose(got, np.exp(1j * phi), atol=1e-5)
    tone.seek(500)
    np.testing.assert_allclose(np.asarray(tone.read(3)),
                               np.exp(1j * phi[500:503]), atol=1e-5)
    # noise is not repeated between frames but reproducible on re-reading
    ng = bt.NoiseGenerator((400, 2), t0, 1e3, samples_per_frame=100,
                           dtype='f4', seed=5)
    d1 = ng.read()
    assert not np.any(d1[:100] == d1[100:200])
    ng.seek(100)
    np.testing.assert_array_equal(ng.read(100), d1[100:200])
