"""Channelize/Dechannelize semantics after the reference's
tests/test_channelize.py:14-250, on a synthetic 8-thread real stream shaped
like its VDIF sample (40000 x 8 float32 at 32 MHz) and a complex one."""
import numpy as np
import pytest

from test_tasks import bt, start_time  # noqa: F401  (fixture)
from test_kernels import assert_voltage, cnoise

N = 1024


def vdif_like(bt, complex_data=False, **kwargs):
    rng = np.random.default_rng(42)
    x = cnoise(rng, (40000, 8))
    if not complex_data:
        x = x.real.copy()
    fh = bt.ArrayStream(x, start_time(bt), 32e6, samples_per_frame=20000,
                        **kwargs)
    return x, fh


@pytest.mark.parametrize('complex_data', [False, True])
def test_channelize_task(bt, complex_data):
    x, fh = vdif_like(bt, complex_data)
    part = x[:N * (40000 // N)].reshape(-1, N, 8)
    ref = (np.fft.fft if complex_data else np.fft.rfft)(part, axis=1)
    ct = bt.Channelize(fh, N)
    data1 = ct.read()
    assert ct.tell() == ct.shape[0] == 39
    assert abs((ct.time - ct.start_time) - 39 * N / 32e6) < 1e-9
    assert data1.dtype == np.complex64
    assert_voltage(data1, ref.astype('c8'))
    ct.seek(-3, 2)
    assert ct.tell() == ct.shape[0] - 3
    data2 = ct.read()
    assert data2.shape[0] == 3
    np.testing.assert_array_equal(data2, data1[-3:])
    ct.seek(-2, 2)
    with pytest.raises(EOFError):
        ct.read(10)
    cr = repr(ct)
    assert cr.startswith('Channelize(ih') and f'n={N}' in cr
    with pytest.raises(AttributeError):
        ct.frequency
    with pytest.raises(AttributeError):
        ct.sideband
    ct.close()
    assert ct.closed
    with pytest.raises(ValueError):
        ct.read(1)
    with pytest.raises(AttributeError):
        ct.ih
    with pytest.raises(AssertionError):      # more channels than samples
        bt.Channelize(fh, 65536)
    with pytest.raises(AssertionError):
        bt.Channelize(fh, 400001)
    # Any number of channels: 1000 is not a power of two (chirp-z transform).
    c1000 = bt.Channelize(fh, 1000)
    ref1000 = (np.fft.fft if complex_data else np.fft.rfft)(
        x[:40000].reshape(-1, 1000, 8), axis=1)
    assert_voltage(c1000.read(), ref1000.astype('c8'))


@pytest.mark.parametrize('spf', [1, 16, 33])
def test_channelize_samples_per_frame(bt, spf):
    x, fh = vdif_like(bt)
    ref = np.fft.rfft(x[:39 * N].reshape(-1, N, 8), axis=1).astype('c8')
    ct = bt.Channelize(fh, N, samples_per_frame=spf)
    data1 = ct.read()
    assert len(data1) % spf == 0 and len(data1) // spf == 39 // spf
    assert_voltage(data1, ref[:len(data1)])
    ct.seek(-3, 2)
    data2 = ct.read()
    np.testing.assert_array_equal(data2, data1[-3:])


@pytest.mark.parametrize('complex_data', [False, True])
def test_frequency_and_dechannelize(bt, complex_data):
    sideband = np.tile([-1, 1], 4)
    frequency = 311.25e6 + (np.arange(8.) // 2) * 16e6
    x, fh = vdif_like(bt, complex_data, frequency=frequency, sideband=sideband)
    ct = bt.Channelize(fh, N)
    freqs = (np.fft.fftfreq if complex_data else np.fft.rfftfreq)(N, 1 / 32e6)
    np.testing.assert_array_equal(ct.sideband, sideband)
    np.testing.assert_allclose(ct.frequency,
                               frequency + sideband * freqs[:, np.newaxis],
                               rtol=1e-14)
    dt = (bt.Dechannelize(ct) if complex_data
          else bt.Dechannelize(ct, N, dtype=fh.dtype))
    nrec = (40000 // N) * N
    assert dt.shape == (nrec, 8) and dt.dtype == fh.dtype
    data = dt.read()
    np.testing.assert_allclose(data, x[:nrec], atol=1e-5)
    np.testing.assert_array_equal(dt.frequency, fh.frequency)
    np.testing.assert_array_equal(dt.sideband, fh.sideband)
    dr = repr(dt)
    assert dr.startswith('Dechannelize(ih')
    if complex_data:
        np.testing.assert_array_equal(ct.inverse(ct).read(), data)
    else:
        with pytest.raises(ValueError):
            bt.Dechannelize(ct, dtype=fh.dtype)     # real data need n
