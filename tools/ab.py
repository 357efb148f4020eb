"""A/B timing of kernel variants through the C ABI (development aid, GPU only).

Usage: python tools/ab.py WORKLOAD [variant ...]
  WORKLOAD  C2 (8ch x 2pol, N=2^20, 16 frames) or C4 (2pol, N=2^24, 8 frames)
  variant   comma-separated tuning knobs, e.g. "row_tma=1,col_tma=0"
            ("base" = library defaults)

For every variant the dedispersion passes and the fused
Channelize(1024) -> Power -> Integrate kernel run on the same synthetic block;
per-kernel times come from the library's own CUDA-event profile.  The
outputs of every variant are compared with those of the first one, so a
variant that computes something else shows up here before the parity tests.
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseband_tasks_b200 import _cabi  # noqa: E402

lib = _cabi.lib()
dev = torch.device('cuda:0')

SHAPES = {
    # N, series, frames, pad_start, pad_end, n_chirp
    'C2': (1 << 20, 16, 16, 74847, 80161, 8),
    'C4': (1 << 24, 2, 8, 1889551, 2075345, 1),
    'C1': (1 << 21, 1, 16, 431817, 458523, 1),
}


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def set_variant(spec):
    if spec == 'base':
        return
    for item in spec.split(','):
        key, value = item.split('=')
        lib.check(lib.bbt_tune_set(key.encode(), int(value)))


def profile(fn, reps):
    lib.bbt_profile_enable(1)
    torch.cuda.synchronize()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    lib.bbt_profile_enable(0)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.check(lib.bbt_profile_report(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, count, total = line.split()
        out[name] = float(total) / int(count)
    return out


def main():
    wl = sys.argv[1]
    variants = sys.argv[2:] or ['base']
    N, S, frames, pad_start, pad_end, n_chirp = SHAPES[wl]
    frames = int(os.environ.get('AB_FRAMES', frames))
    spf = N - pad_start - pad_end
    n_in = spf * (frames - 1) + N
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(n_in * S, 2, device=dev, generator=g)
    x = torch.view_as_complex(x).contiguous()
    out = torch.empty(spf * frames * S, dtype=torch.complex64, device=dev)
    smap = (np.arange(S) // max(1, S // n_chirp)).astype(np.int32)
    freq = 1400. + 8. * np.arange(n_chirp)
    fref = np.full(n_chirp, 1400.)
    sb = np.ones(n_chirp, np.int8)
    n_chan = 1024
    m = S // 2
    ratio = 500. if wl == 'C4' else 7.8125
    n_spec = spf * frames // n_chan
    n_bins = int(n_spec / ratio)
    off = torch.from_numpy(np.around(
        np.arange(n_bins + 1) * ratio).astype(np.int64)).to(dev)
    ref = None
    print(torch.cuda.get_device_name(0), wl, flush=True)
    for spec in variants:
        set_variant(spec)
        plan = ctypes.c_void_p()
        dbl = ctypes.POINTER(ctypes.c_double)
        lib.check(lib.bbt_dedisperse_plan_create(
            ctypes.byref(plan), N, S, pad_start, spf, n_chirp,
            smap.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            freq.ctypes.data_as(dbl), fref.ctypes.data_as(dbl),
            sb.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)),
            -100., 8., 0., 0))
        wb = lib.bbt_dedisperse_work_bytes(plan, frames)
        work = torch.empty(max(wb, 8) // 8, dtype=torch.complex64, device=dev)
        sums = torch.zeros(n_bins * n_chan * max(m, 1) * 4,
                           dtype=torch.float32, device=dev)
        cnt = torch.zeros(n_bins, dtype=torch.int64, device=dev)

        def step():
            lib.check(lib.bbt_dedisperse_exec(
                plan, ptr(x), spf * S, frames, 0, ptr(out), spf * S,
                ptr(work), stream()))
            if m >= 1:
                lib.check(lib.bbt_channelize_power_integrate_exec(
                    ptr(out), n_chan, m, n_spec, 0, ptr(off), 0, n_bins,
                    ptr(sums), ptr(cnt), 0, stream()))
        out.zero_()
        for _ in range(2):
            step()
        sums.zero_()
        cnt.zero_()
        step()
        torch.cuda.synchronize()
        got = (out.clone(), sums.clone(), cnt.clone())
        reps = int(os.environ.get('AB_REPS', 5))
        for _ in range(int(os.environ.get('AB_SOAK', 0))):
            step()
        times = profile(step, reps)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        e1.synchronize()
        total = e0.elapsed_time(e1) / 5
        line = '  '.join(f'{k} {v:.3f}' for k, v in sorted(times.items()))
        msg = ''
        if ref is None:
            ref = got
        else:
            rms = float(ref[0].abs().pow(2).mean().sqrt())
            dv = float((got[0] - ref[0]).abs().max()) / rms
            dp = float((got[1] - ref[1]).abs().max()
                       / ref[1].abs().max().clamp_min(1e-30))
            msg = (f'  | vs first: voltage {dv:.2e} x rms, power {dp:.2e} rel,'
                   f' counts {"equal" if torch.equal(got[2], ref[2]) else "DIFFER"}')
        gs = spf * frames * S / (total * 1e-3) / 1e9
        print(f'{spec:40s} step {total:.3f} ms ({gs:.1f} Gs/s)  {line}{msg}',
              flush=True)
        lib.bbt_dedisperse_plan_destroy(plan)
        del work


if __name__ == '__main__':
    main()
