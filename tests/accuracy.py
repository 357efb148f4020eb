"""Numerical accuracy of the GPU FFT and dedispersion against float64 (GPU only).

Usage: [BBT_B200_LIB=...] python tests/accuracy.py
Prints max and RMS errors relative to the RMS of the float64 result, next to
the same figures for numpy's single-precision FFT.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import baseband_tasks_b200 as bt  # noqa: E402
from baseband_tasks_b200.fourier import fft_maker  # noqa: E402
import bbt_oracle as orc  # noqa: E402  (the checker)


def rel(got, want):
    d = np.abs(got - want)
    r = np.sqrt((np.abs(want) ** 2).mean())
    return d.max() / r, np.sqrt((d ** 2).mean()) / r


rng = np.random.default_rng(0)
for log2n in (10, 12, 14, 20, 24):
    n = 1 << log2n
    y = np.exp(2j * np.pi * 200. * np.arange(n) / n).astype('c8')
    x = (rng.normal(size=n) + 1j * rng.normal(size=n)).astype('c8')
    fft = fft_maker((n,), 'c8')
    for name, a in (('tone', y), ('noise', x)):
        want = np.fft.fft(a.astype('c16'))
        print('fft 2^%d %-5s gpu max %.2e rms %.2e | numpy c8 max %.2e rms %.2e'
              % ((log2n, name) + rel(fft(a), want) + rel(np.fft.fft(a), want)),
              flush=True)

rate, freq, dm = 512e6, 8192e6, 1000.
N = 1 << 24
pad_start, pad_end = 1889551, 2075345
spf = N - pad_start - pad_end
x = (rng.normal(size=(N, 2)) + 1j * rng.normal(size=(N, 2))).astype('c8')
src = bt.ArrayStream(x, bt.Time(1289567655), rate, frequency=freq, sideband=1)
got = bt.Dedisperse(src, dm, samples_per_frame=spf).read()
op = orc.DispersePlan(-dm, freq / 1e6, 1, rate / 1e6, True, N, N, (2,),
                      samples_per_frame=spf, fast_len=orc.next_pow2)
want = orc.disperse(x.astype('c16'), op)
want32 = orc.disperse(x, op)
print('dedisperse 2^24 gpu max %.2e rms %.2e | numpy c8 max %.2e rms %.2e'
      % (rel(got, want) + rel(want32, want)), flush=True)
