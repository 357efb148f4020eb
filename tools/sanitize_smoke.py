"""Small instances of every hot kernel through the Task API, checked against
the oracle; meant to run under compute-sanitizer (one tool per run):

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

(On the pool this round was built on compute-sanitizer was closed by the
operators; the script then still serves as a quick GPU-side parity check.)

Covers the persistent TMA kernels (column passes, the row pass in both
formulations, the bulk-copy channelizer), the single-pass dedispersion, fold,
the filter bank and the payload decoder; results are checked against the
oracle so that a run that "passes" the sanitizer also computed the right thing.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import baseband_tasks_b200 as bt   # noqa: E402
import bbt_oracle as orc           # noqa: E402
from baseband_tasks_b200 import _cabi  # noqa: E402

T0 = bt.Time(1289567655)
rng = np.random.default_rng(5)


def cnoise(shape):
    return (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype('c8')


def check(name, got, want, tol=1e-5):
    rms = np.sqrt(np.mean(np.abs(want) ** 2))
    err = np.abs(got - want).max() / rms
    print(f'{name}: max err / rms = {err:.2e}', flush=True)
    assert err < tol, (name, err)


def chain(n_log2, shape, dm, rate, freq, n_chan, hint=0):
    _cabi.lib().check(_cabi.lib().bbt_tune_set(b'dd_hint', hint))
    N = 1 << n_log2
    # A dispersion measure for which the padding is about a fifth of a frame.
    f_mhz, r_mhz = freq / 1e6, rate / 1e6
    width = (1. / (f_mhz - r_mhz / 2) ** 2
             - 1. / (f_mhz + r_mhz / 2) ** 2) / 2.41e-4
    dm = (N / 5) / rate / width
    probe = orc.DispersePlan(-dm, freq / 1e6, 1, rate / 1e6, True, 3 * N, N,
                             shape, fast_len=orc.next_pow2,
                             samples_per_frame=1)
    spf = N - probe.pad_start - probe.pad_end
    n = 2 * spf + N
    x = cnoise((n,) + shape)
    src = bt.ArrayStream(x, T0, rate, samples_per_frame=N, frequency=freq,
                         sideband=1, polarization=np.array(['X', 'Y']))
    dd = bt.Dedisperse(src, dm, samples_per_frame=spf)
    op = orc.DispersePlan(-dm, freq / 1e6, 1, rate / 1e6, True, n, N, shape,
                          samples_per_frame=spf, fast_len=orc.next_pow2)
    y = orc.disperse(x, op)
    check(f'dedisperse 2^{n_log2} {shape} hint {hint}', dd.read(), y)
    it = bt.Integrate(bt.Power(bt.Channelize(dd, n_chan)), 7, average=False)
    power = orc.power(orc.channelize(y, n_chan), axis=-1)
    want, cnt = orc.integrate(power, np.arange(0, power.shape[0] + 1, 7))
    got = it.read()
    assert np.array_equal(got['count'].reshape(len(got), -1)[:, 0],
                          cnt.ravel()[:len(got)])
    check('  channelize-power-integrate', got['data'], want[:len(got)])
    fold = bt.Fold(bt.Power(dd), 64, bt.PolynomialPhase([0.1, 977.3], T0),
                   average=False)
    f = fold.read()
    assert f['count'].reshape(f.shape[0], f.shape[1], -1)[..., 0].sum() \
        == dd.shape[0]


# Planar three-pass plans: rows of 2^11 (hint: 2^5 columns), 2^13 and 2^14
# points (tensor-map column passes + dd_row2), narrow samples (bulk-copy
# channelizer).
chain(16, (2,), 30., 16e6, 800e6, 256, hint=5 | 256)
chain(16, (2,), 30., 16e6, 800e6, 256, hint=3 | 256)
chain(17, (2,), 60., 16e6, 800e6, 1024, hint=3 | 256)
# Interleaved plan (wide samples), and the single-pass kernel.
chain(15, (4, 2), 15., 16e6, 800e6, 64, hint=5 | 512)
chain(13, (2,), 4., 16e6, 800e6, 64)
_cabi.lib().check(_cabi.lib().bbt_tune_set(b'dd_hint', 0))
# Filter bank on 8-bit samples and the payload decoder.
x8 = np.clip(np.round(rng.normal(size=(64 * 256, 2)) * 20), -127,
             127).astype('f4')
src = bt.ArrayStream(x8, T0, 800e6, samples_per_frame=4096, frequency=800e6,
                     sideband=-1, polarization=np.array(['X', 'Y']))
resp = bt.sinc_hamming(4, 256)
check('pfb', bt.PolyphaseFilterBank(src, resp).read(),
      orc.pfb(x8.astype('f8'), resp, ih_samples_per_frame=4096).astype('c8'))
words = rng.integers(0, 256, 4096, dtype=np.uint8)
ps = bt.PayloadStream(words, 2, (2,), T0, 1e6)
assert np.array_equal(ps.read().ravel(),
                      orc.decode_payload(words, 2, bt.payload_levels(2)))
print('sanitize smoke ok,', _cabi.lib().bbt_launch_count(), 'launches')
