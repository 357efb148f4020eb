"""Polyphase filter bank semantics after the reference's
tests/test_pfb.py:37-103 (the CHIME-like 4-tap, 2048-sample filter on
Philox noise, real and complex), in single precision."""
import numpy as np
import pytest

from test_tasks import bt, start_time  # noqa: F401  (fixture)
from test_kernels import assert_voltage


@pytest.mark.parametrize('offset', [0, 40])
@pytest.mark.parametrize('dtype', ['f4', 'c8'])
def test_understanding(bt, dtype, offset):
    """The filter bank equals multiplying n_tap blocks with the response,
    summing them and transforming; both task variants agree, and timestamps
    are centred on the filter."""
    n = 2048
    nh = bt.NoiseGenerator((100 * n,), start_time(bt), 1e3, seed=12345,
                           samples_per_frame=128, dtype=dtype)
    chime = bt.sinc_hamming(4, n)
    nh.seek(offset * n)
    d = nh.read(5 * n).reshape(-1, n).astype('c16' if dtype == 'c8' else 'f8')
    ft = np.fft.fft if dtype == 'c8' else np.fft.rfft
    want = np.stack([ft((chime * d[:4]).sum(0)), ft((chime * d[1:]).sum(0))])
    # the frequency-selection view of the same thing
    np.testing.assert_allclose(ft((chime * d[:4]).ravel())[::4], want[0],
                               atol=1e-9 * np.abs(want[0]).max())
    pfb = bt.PolyphaseFilterBankSamples(nh, chime)
    # 97 spectra fit; whole frames of them are offered (TaskBase semantics)
    assert pfb.shape[1:] == (n if dtype == 'c8' else n // 2 + 1,)
    assert 80 <= pfb.shape[0] <= 97
    assert pfb.shape[0] % pfb.samples_per_frame == 0
    assert pfb.dtype == np.complex64
    assert pfb.sample_rate == 1e3 / n
    # 3 blocks of padding, split half and half: 1.5 blocks later
    assert abs((pfb.start_time - nh.start_time) - 1.5 * n / 1e3) < 1e-9
    pfb.seek(offset)
    got = pfb.read(2)
    assert_voltage(got, want.astype('c8'))
    pfb2 = bt.PolyphaseFilterBank(nh, chime)
    assert pfb2.shape == pfb.shape and pfb2.start_time == pfb.start_time
    pfb2.seek(offset)
    assert_voltage(pfb2.read(2), want.astype('c8'))
    r = repr(pfb)
    assert r.startswith('PolyphaseFilterBankSamples(ih')
