"""Raw kernel timings through the C ABI (development aid, GPU only).

Usage: python tools/rawbench.py [section ...]   (sections: copy fft dd chan)
Prints one line per measurement: name, ms, algorithmic GB/s.
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


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t, byte_offset=0):
    return ctypes.c_void_p(t.data_ptr() + byte_offset)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def report(name, ms, nbytes):
    best, med = ms
    print(f'{name:60s} best {best:9.3f} ms  med {med:9.3f} ms  '
          f'{nbytes / best / 1e6:8.1f} GB/s', flush=True)


def sec_copy():
    n = 1 << 30   # bytes
    a = torch.empty(n, dtype=torch.uint8, device=dev)
    b = torch.empty(n, dtype=torch.uint8, device=dev)
    report('torch copy 1 GiB', timeit(lambda: b.copy_(a)), 2 * n)
    for rows in (64, 256, 1024):
        for chunk in (32, 64, 128, 256, 512, 1024):
            row_stride = n // rows
            n_tiles = row_stride // chunk
            report(f'strided copy rows={rows} chunk={chunk}B',
                   timeit(lambda: lib.check(lib.bbt_strided_copy_bench(
                       ptr(a), ptr(b), rows, row_stride, chunk, n_tiles,
                       stream()))), 2 * n)


def sec_fft():
    total = 1 << 27   # complex points = 1 GiB
    x = torch.randn(total, dtype=torch.complex64, device=dev)
    y = torch.empty_like(x)
    for log2n in (6, 8, 10, 11, 12, 13, 14):
        n = 1 << log2n
        plan = ctypes.c_void_p()
        lib.check(lib.bbt_fft_plan_create(ctypes.byref(plan), n, total // n,
                                          1, 0, 0, 1.))
        report(f'fft c2c n={n} contiguous',
               timeit(lambda: lib.check(lib.bbt_fft_exec(
                   plan, ptr(x), ptr(y), None, stream()))), 16 * total)
        lib.bbt_fft_plan_destroy(plan)
    for log2n, inner in ((8, 4096), (10, 1024), (6, 16384 * 16)):
        n = 1 << log2n
        outer = total // n // inner
        plan = ctypes.c_void_p()
        lib.check(lib.bbt_fft_plan_create(ctypes.byref(plan), n, outer,
                                          inner, 0, 0, 1.))
        report(f'fft c2c n={n} inner={inner}',
               timeit(lambda: lib.check(lib.bbt_fft_exec(
                   plan, ptr(x), ptr(y), None, stream()))), 16 * total)
        lib.bbt_fft_plan_destroy(plan)
    for log2n in (20, 24):
        n = 1 << log2n
        plan = ctypes.c_void_p()
        lib.check(lib.bbt_fft_plan_create(ctypes.byref(plan), n, total // n,
                                          1, 0, 0, 1.))
        work = torch.empty_like(x)
        report(f'fft c2c n=2^{log2n} four-step',
               timeit(lambda: lib.check(lib.bbt_fft_exec(
                   plan, ptr(x), ptr(y), ptr(work), stream()))), 32 * total)
        lib.bbt_fft_plan_destroy(plan)
    xt = torch.fft.fft(x.view(-1, 1 << 20), dim=1)
    report('torch.fft (cuFFT) n=2^20, context only',
           timeit(lambda: torch.fft.fft(x.view(-1, 1 << 20), dim=1)),
           16 * total)
    del xt


def dd_plan(N, S, pad_start, n_valid, log2n1):
    smap = np.zeros(S, np.int32)
    f = np.array([1400.])
    r = np.array([1400.])
    s = np.array([1], np.int8)
    plan = ctypes.c_void_p()
    lib.check(lib.bbt_dedisperse_plan_create(
        ctypes.byref(plan), N, S, pad_start, n_valid, 1,
        smap.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
        f.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
        r.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
        s.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)),
        -100., 8., 0., log2n1))
    return plan


def sec_dd():
    P, I = 256, 512       # planar / interleaved work-buffer layout
    HC, HR = 4096, 8192   # half-size tiles in column / row passes
    FR = 16384            # full-size row tiles for the interleaved layout
    for (N, S, frames, hints) in (
            (1 << 20, 16, 16, (0, I | FR, P, I | HC)),
            (1 << 24, 2, 8, (0, HC, 11 | P | HR)),
            (1 << 22, 2, 4, (0, HC, 9 | P | HR)),
            (1 << 14, 2050, 2, (0,)),
            (1 << 13, 2050, 4, (0,))):
        pad = N // 5
        spf = N - pad
        n_in = spf * (frames - 1) + N
        x = torch.randn(n_in * S, dtype=torch.complex64, device=dev)
        out = torch.empty(spf * frames * S, dtype=torch.complex64, device=dev)
        for h in hints:
            if N <= 16384 and h:
                continue
            plan = dd_plan(N, S, pad // 2, spf, h)
            wb = lib.bbt_dedisperse_work_bytes(plan, frames)
            work = torch.empty(max(wb, 8) // 8, dtype=torch.complex64,
                               device=dev)
            passes = 3 if (N > 16384 or (S > 1 and N > 1024)) else 1
            report(f'dedisperse N=2^{int(np.log2(N))} S={S} frames={frames} '
                   f'log2n1={h}',
                   timeit(lambda: lib.check(lib.bbt_dedisperse_exec(
                       plan, ptr(x), spf * S, frames, 0, ptr(out), spf * S,
                       ptr(work), stream()))), 16 * passes * N * S * frames)
            lib.bbt_dedisperse_plan_destroy(plan)
            del work


def sec_chan():
    n, m = 1024, 8
    n_spec = 1 << 13
    x = torch.randn(n_spec * n * m * 2, dtype=torch.complex64, device=dev)
    out = torch.empty(n_spec * n * m * 4, dtype=torch.float32, device=dev)
    report('channelize(1024)+power m=8',
           timeit(lambda: lib.check(lib.bbt_channelize_power_exec(
               ptr(x), ptr(out), n, m, n_spec, stream()))),
           x.numel() * 8 + out.numel() * 4)
    for ratio in (7.8125, 500.):
        n_bins = int(n_spec / ratio)
        off = torch.from_numpy(np.around(
            np.arange(n_bins + 1) * ratio).astype(np.int64)).to(dev)
        s = torch.zeros(n_bins * n * m * 4, dtype=torch.float32, device=dev)
        c = torch.zeros(n_bins, dtype=torch.int64, device=dev)
        report(f'channelize(1024)+power+integrate m=8 ratio={ratio}',
               timeit(lambda: lib.check(
                   lib.bbt_channelize_power_integrate_exec(
                       ptr(x), n, m, n_spec, 0, ptr(off), 0, n_bins, ptr(s),
                       ptr(c), 0, stream()))), x.numel() * 8)
    n, m = 1024, 1
    n_spec = 1 << 16
    x = torch.randn(n_spec * n * m * 2, dtype=torch.complex64, device=dev)
    ratio = 500.
    n_bins = int(n_spec / ratio)
    off = torch.from_numpy(np.around(
        np.arange(n_bins + 1) * ratio).astype(np.int64)).to(dev)
    s = torch.zeros(n_bins * n * m * 4, dtype=torch.float32, device=dev)
    c = torch.zeros(n_bins, dtype=torch.int64, device=dev)
    report('channelize(1024)+power+integrate m=1 ratio=500',
           timeit(lambda: lib.check(
               lib.bbt_channelize_power_integrate_exec(
                   ptr(x), n, m, n_spec, 0, ptr(off), 0, n_bins, ptr(s),
                   ptr(c), 0, stream()))), x.numel() * 8)


if __name__ == "__main__":
    secs = sys.argv[1:] or ['copy', 'fft', 'dd', 'chan']
    print(torch.cuda.get_device_name(0), flush=True)
    for s in secs:
        globals()['sec_' + s]()
