"""One dedisperse call per configuration, for ncu captures (development aid).

Usage: python tools/prof_dd.py [C4|C2|...]  -- launches the three passes of
each named configuration once (after plan creation)."""
import ctypes
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.rawbench import lib, dev, ptr, stream, dd_plan  # noqa: E402

CONFIGS = {'C4': (1 << 24, 2, 2, 0), 'C2': (1 << 20, 16, 4, 512),
           'C2P': (1 << 20, 16, 4, 0)}

for name in (sys.argv[1:] or ['C4', 'C2']):
    N, S, frames, hint = CONFIGS[name]
    pad = N // 5
    spf = N - pad
    n_in = spf * (frames - 1) + N
    x = torch.randn(n_in * S, dtype=torch.complex64, device=dev)
    out = torch.empty(spf * frames * S, dtype=torch.complex64, device=dev)
    plan = dd_plan(N, S, pad // 2, spf, hint)
    wb = lib.bbt_dedisperse_work_bytes(plan, frames)
    work = torch.empty(max(wb, 8) // 8, dtype=torch.complex64, device=dev)
    lib.check(lib.bbt_dedisperse_exec(plan, ptr(x), spf * S, frames, 0,
                                      ptr(out), spf * S, ptr(work), stream()))
    torch.cuda.synchronize()
    print(name, 'done', flush=True)
