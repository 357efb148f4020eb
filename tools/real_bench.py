"""Timing of Dedisperse on a real-valued stream (development aid, GPU only):
python tools/real_bench.py [log2n] [series] [frames]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import baseband_tasks_b200 as bt  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 32
N = 1 << log2n
rate, dm = 16e6, 30.
g = torch.Generator(device='cuda').manual_seed(1)
sb = np.array(([1, -1] * S)[:S])
probe = bt.Dedisperse(bt.ArrayStream(
    torch.zeros((2 * N, S), device='cuda'), bt.Time(0), rate,
    frequency=400e6, sideband=sb), dm, samples_per_frame=N // 2)
pad = probe._pad_start + probe._pad_end
del probe
spf = N - pad
n = frames * spf + pad
x = torch.randn((n, S), device='cuda', generator=g)
src = bt.ArrayStream(x, bt.Time(0), rate, frequency=400e6,
                     sideband=sb)
dd = bt.Dedisperse(src, dm, samples_per_frame=spf)
assert dd._ih_samples_per_frame == N, dd._ih_samples_per_frame
from baseband_tasks_b200 import base
base.BLOCK_BYTES = 1 << 40
for _ in range(3):
    dd.seek(0)
    y = dd.read_device()
e0 = torch.cuda.Event(enable_timing=True)
e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    dd.seek(0)
    y = dd.read_device()
e1.record()
e1.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f'real Dedisperse N=2^{log2n} S={S} frames={frames}: {ms:.3f} ms, '
      f'{y.shape[0] * S / ms / 1e6:.1f} G real samples/s')
