"""Benchmark of the Dedisperse -> Channelize -> Power -> Integrate chain.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C2|C4|C5]
                    [--impl b200|reference]

One step = one pass of the chain over one block of F overlap-save frames of
a synthetic NoiseGenerator stream:

  C2 (default; BASELINE.json configs[1]): (T, 8, 2) complex64, 8 channels of
     8 MHz at 1372+8k MHz, Dedisperse(DM=100, N=2^20) -> Channelize(1024) ->
     Power -> Integrate(1 ms).
  C4 (configs[3], the north-star target): (T, 2) complex64, 512 MHz at
     8192 MHz, Dedisperse(DM=1000, N=2^24) -> Channelize(1024) -> Power ->
     Integrate(1 ms).
  C5 (configs[4]): the C4 stream -> Dedisperse -> Power -> Fold(512 bins,
     polynomial phase), the folded profile summed over ranks with NCCL.

``value``: complex source samples (time x channel x polarization) per second
through the public Task API with the input block resident in HBM.  ``e2e``:
the same with the block in pinned host memory, copied to the device inside
the timed region (frame by frame on a copy stream, overlapped with the work
on the frames that have arrived), and the integrated spectra copied back.
Timing: CUDA events, max over ranks.  With N > 1 every rank processes its own
time block of the stream (time-block sharding with an overlap-save halo; no
collective on the Integrate chains), so scaling is weak.

``--impl reference`` times the reference's CPU path -- its numpy arithmetic
restated in oracle/bbt_oracle.py, since the reference itself needs astropy and
baseband, which are not installable here -- on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # rate/Hz, channel frequencies/Hz, DM, log2 N, frames per step
    'C2': dict(rate=8e6, freq=(1372e6 + 8e6 * np.arange(8)).reshape(8, 1),
               sample_shape=(8, 2), dm=100., log2n=20, frames=16,
               n_chan=1024, step=1e-3, seed=1234567 + 2,
               desc='8ch x 2pol x 8 MHz c64 -> Dedisperse(DM=100, N=2^20) -> '
                    'Channelize(1024) -> Power -> Integrate(1 ms)'),
    'C4': dict(rate=512e6, freq=8192e6, sample_shape=(2,), dm=1000.,
               log2n=24, frames=8, n_chan=1024, step=1e-3, seed=1234567 + 4,
               desc='2pol x 512 MHz c64 -> Dedisperse(DM=1000, N=2^24) -> '
                    'Channelize(1024) -> Power -> Integrate(1 ms)'),
    # configs[4]: fold with a polynomial phase; profiles reduced over ranks.
    'C5': dict(rate=512e6, freq=8192e6, sample_shape=(2,), dm=1000.,
               log2n=24, frames=8, n_chan=None, step=None, seed=1234567 + 5,
               fold=dict(n_phase=512, coef=[0.25, 29.946923,
                                            -3.77535e-10 / 2.]),
               desc='2pol x 512 MHz c64 -> Dedisperse(DM=1000, N=2^24) -> '
                    'Power -> Fold(512 bins, polynomial phase), NCCL reduce '
                    'of the profile'),
}


def framing(w):
    """Overlap-save framing, as Disperse.__init__ derives it
    (reference dispersion.py:54-93), for samples_per_frame = N - pad."""
    k = 1. / 2.41e-4
    rate_mhz = w['rate'] / 1e6
    f = np.asarray(w['freq'], float) / 1e6
    lo, hi = f - rate_mhz / 2, f + rate_mhz / 2
    fref = np.mean(lo + hi) / 2.
    d = k * -w['dm']

    def delay(x):
        return d * (1. / x ** 2 - 1. / fref ** 2)
    dmax = max(np.max(delay(lo)), np.max(delay(hi)))
    dmin = min(np.min(delay(lo)), np.min(delay(hi)))
    pad_start = int(np.ceil(dmax * w['rate']))
    pad_end = int(np.ceil(-dmin * w['rate']))
    N = 1 << w['log2n']
    spf = N - pad_start - pad_end
    return N, spf, pad_start, pad_end


def model_bytes(w):
    """Pre-registered algorithmic bytes per source sample."""
    N, spf, _, _ = framing(w)
    eff = spf / N
    if w.get('fold'):
        return 48. / eff + 8.
    per_bin = w['step'] * w['rate'] / w['n_chan']
    return 48. / eff + 8. + 8. / per_bin


class ClocksSampler:
    """nvidia-smi polled every 100 ms in the background; every line is kept
    with the time it arrived, so the samples taken under load can be picked."""

    def __init__(self, device_index):
        import threading
        cmd = ['nvidia-smi', '-i', str(device_index),
               '--query-gpu=clocks.sm,clocks.max.sm,'
               'clocks_event_reasons.hw_slowdown,'
               'clocks_event_reasons.hw_thermal_slowdown,'
               'clocks_event_reasons.sw_thermal_slowdown,'
               'clocks_event_reasons.sw_power_cap',
               '--format=csv,noheader,nounits', '-lms', '100']
        self.samples = []
        try:
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.monotonic(), line))

    def count(self, t0):
        return sum(1 for t, _ in self.samples if t >= t0)

    def summary(self, t0, t1):
        """Median SM clock and throttle reasons of the samples in [t0, t1]."""
        if self.proc is None:
            return None
        self.proc.terminate()
        sm, smax, reasons = [], 0., set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for t, line in list(self.samples):
            if not t0 <= t <= t1:
                continue
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                smax = max(smax, float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return None
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': smax,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------- GPU arm
def make_block(w, rank):
    """One step's input: F frames of the NoiseGenerator stream (host)."""
    import baseband_tasks_b200 as bt
    N, spf, _, _ = framing(w)
    n_in = (w['frames'] - 1) * spf + N
    gen_spf = 1 << 20
    # Rank r owns frames [r F, (r+1) F) of the stream: its block starts at
    # r F spf (a halo of pad samples is shared with the next rank).
    start = rank * w['frames'] * spf
    total = start + n_in
    total = -(-total // gen_spf) * gen_spf
    nh = bt.NoiseGenerator((total,) + w['sample_shape'], bt.Time(1289567655),
                           w['rate'], samples_per_frame=gen_spf,
                           dtype='c8', seed=w['seed'])
    nh.seek(start)
    return nh.read(n_in), start


def build_chain(w, data, start):
    import baseband_tasks_b200 as bt
    N, spf, _, _ = framing(w)
    t0 = bt.Time(1289567655) + start / w['rate']
    pol = np.array(['X', 'Y'])
    src = bt.ArrayStream(data, t0, w['rate'], samples_per_frame=1 << 20,
                         frequency=w['freq'], sideband=1, polarization=pol)
    dd = bt.Dedisperse(src, w['dm'], samples_per_frame=spf)
    assert dd._ih_samples_per_frame == N, (dd._ih_samples_per_frame, N)
    if w.get('fold'):
        # Phase polynomial referred to the start of the whole stream, so all
        # ranks fold on the same ephemeris.
        poly = bt.PolynomialPhase(w['fold']['coef'], bt.Time(1289567655))
        return src, bt.Fold(bt.Power(dd), w['fold']['n_phase'], poly,
                            average=False)
    ch = bt.Channelize(dd, w['n_chan'])
    pw = bt.Power(ch)
    it = bt.Integrate(pw, w['step'])
    return src, it


def run_b200(args):
    import torch
    import torch.distributed as dist
    import baseband_tasks_b200 as bt
    from baseband_tasks_b200 import _cabi

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    lib = _cabi.lib()
    w = WORKLOADS[args.workload]
    N, spf, pad_start, pad_end = framing(w)
    S = int(np.prod(w['sample_shape']))

    host_np, start = make_block(w, rank)
    host = torch.from_numpy(host_np).pin_memory()
    dev_in = host.to('cuda', non_blocking=True)
    torch.cuda.synchronize()
    samples_per_step = w['frames'] * spf * S

    # HBM-resident chain.
    src, chain = build_chain(w, dev_in, start)
    out_host = None

    from baseband_tasks_b200 import parallel
    folding = bool(w.get('fold'))

    def run_chain(c):
        c.seek(0)
        if not folding:
            return c.read_device()
        # Profile sums and counts, summed over ranks (NCCL over NVLink).
        sums, counts = c.read_sums()
        parallel.reduce_sums(sums, counts)
        return sums

    def step_resident():
        return run_chain(chain)

    # End to end: pinned host block -> device -> chain -> host.  The block is
    # copied frame by frame on a copy stream while the chain already works on
    # the frames that have arrived (through the public API: successive reads
    # of the output samples whose input is on the device).
    dev_stage = torch.empty_like(dev_in)
    _, chain_e2e = build_chain(w, dev_stage, start)
    copy_stream = torch.cuda.Stream()
    d2h_stream = torch.cuda.Stream()
    n_fr = w['frames']
    row = host.shape[1:].numel() if host.dim() > 1 else 1
    pieces = [(0, N)] + [(N + (k - 1) * spf, N + k * spf)
                         for k in range(1, n_fr)]
    if folding:
        reads = None
    else:
        # Output samples computable from the first k+1 frames.
        n_out = chain_e2e.shape[0]
        edges = np.asarray(chain_e2e._get_offsets(np.arange(n_out + 1)))
        per_out = w['n_chan']           # upstream samples per spectrum
        reads = []
        done = 0
        for k in range(n_fr):
            avail = (k + 1) * spf
            upto = int(np.searchsorted(edges * per_out, avail, side='right')
                       - 1) if k < n_fr - 1 else n_out
            upto = max(done, min(upto, n_out))
            reads.append((done, upto))
            done = upto

    # The same block as recorded baseband data are stored: 8-bit (re, im)
    # codes, a quarter of the bytes over PCIe, decoded on the device.
    scale8 = 30.
    levels8 = bt.payload_levels(8) / scale8
    host8 = torch.from_numpy(bt.encode_payload(
        host_np, 8, levels8).reshape(host_np.shape[0], -1)).pin_memory()
    dev8 = torch.empty_like(host8, device='cuda')
    d_levels8 = torch.from_numpy(levels8).cuda()
    row_values = 2 * S

    def step_e2e(packed=False):
        nonlocal out_host
        main = torch.cuda.current_stream()
        copy_stream.wait_stream(main)      # previous step is done with the stage
        events = []
        with torch.cuda.stream(copy_stream):
            for a0, a1 in pieces:
                if packed:
                    dev8[a0:a1].copy_(host8[a0:a1], non_blocking=True)
                else:
                    dev_stage[a0:a1].copy_(host[a0:a1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                events.append(ev)

        def arrived(k):
            main.wait_event(events[k])
            if packed:
                a0, a1 = pieces[k]
                lib.check(lib.bbt_decode_exec(
                    ctypes.c_void_p(dev8[a0:a1].data_ptr()),
                    ctypes.c_void_p(dev_stage[a0:a1].data_ptr()),
                    ctypes.c_void_p(d_levels8.data_ptr()),
                    (a1 - a0) * row_values, 8,
                    ctypes.c_void_p(main.cuda_stream)))
        def to_host(part, pos):
            # Copy a finished part back while later frames are still in flight.
            nonlocal out_host
            if out_host is None:
                n_total = chain_e2e.shape[0] if reads is not None \
                    else part.shape[0]
                out_host = torch.empty((n_total,) + tuple(part.shape[1:]),
                                       dtype=part.dtype, pin_memory=True)
            ev = torch.cuda.Event()
            ev.record(main)
            d2h_stream.wait_event(ev)
            with torch.cuda.stream(d2h_stream):
                out_host[pos:pos + part.shape[0]].copy_(part,
                                                        non_blocking=True)
                part.record_stream(d2h_stream)

        if reads is None:
            for k in range(len(pieces)):
                arrived(k)
            res = run_chain(chain_e2e)
            to_host(res, 0)
        else:
            chain_e2e.seek(0)
            res = None
            for k, (b0, b1) in enumerate(reads):
                arrived(k)
                if b1 > b0:
                    res = chain_e2e.read_device(b1 - b0)
                    to_host(res, b0)
        main.wait_stream(d2h_stream)
        return res

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClocksSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        res = step_resident()
    out_bytes = res.numel() * res.element_size()
    l0 = lib.bbt_launch_count()
    t_load = time.monotonic()
    ms = timed(step_resident, args.steps)
    launches = lib.bbt_launch_count() - l0
    # The timed region lasts tens of milliseconds, less than one nvidia-smi
    # poll: every rank keeps the same load on (untimed) for another ~0.5 s so
    # that the clocks are sampled under it.
    n_extra = int(min(5000, np.ceil(500. / max(ms / args.steps, 1e-3))))
    for _ in range(n_extra):
        step_resident()
    barrier()
    clocks = None
    if sampler is not None:
        clocks = sampler.summary(t_load + 0.05, time.monotonic())
        if clocks is not None:
            clocks['window'] = ('timed steps plus %d more identical steps'
                                % n_extra)

    for _ in range(2):
        res_e2e = step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    for _ in range(2):
        step_e2e(packed=True)
    ms_e2e8 = timed(lambda: step_e2e(packed=True), args.steps)

    # Per-kernel durations, CUDA events on the launching stream.
    lib.bbt_profile_enable(1)
    barrier()
    for _ in range(args.steps):
        step_resident()
    barrier()
    lib.bbt_profile_enable(0)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.check(lib.bbt_profile_report(buf, len(buf)))
    kernels = {}
    for line in buf.value.decode().splitlines():
        name, count, total = line.split()
        kernels[name] = (int(count), float(total))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.))
    peak_src = 'measured' if 'hbm_gbs' in peaks else 'fallback'
    # Dominant kernel and its algorithmic bytes per launch.
    points = w['frames'] * N * S              # FFT points per dedisperse pass
    alg = {'dd_col_fwd': 16. * points, 'dd_row': 16. * points,
           'dd_col_inv': 8. * points + 8. * w['frames'] * spf * S,
           'chanpow_integrate': 8. * samples_per_step + out_bytes,
           'fold': 8. * samples_per_step}
    top = max((k for k in kernels if k in alg),
              key=lambda k: kernels[k][1], default=None)
    roofline = None
    if top:
        count, total_ms = kernels[top]
        per_launch_ms = total_ms / count
        achieved = alg[top] / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
                traffic = json.load(f).get(args.workload, {}).get(top)
        except OSError:
            pass
        roofline = {'bound': 'hbm', 'kernel': top, 'achieved': achieved,
                    'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                    'frac': achieved / peak, 'traffic': traffic,
                    'ms_per_launch': per_launch_ms,
                    'algorithmic_bytes_per_launch': alg[top]}
    total_kernel_ms = sum(v[1] for v in kernels.values()) or 1.
    shares = {k: {'launches': v[0], 'ms_per_launch': v[1] / v[0],
                  'share': v[1] / total_kernel_ms}
              for k, v in sorted(kernels.items())}

    value = samples_per_step * args.steps * world / (ms * 1e-3) / 1e9
    e2e = samples_per_step * args.steps * world / (ms_e2e * 1e-3) / 1e9
    mb = model_bytes(w)
    # The CPU baseline is timed on rank 0 of single-GPU runs only.
    cpu = (cpu_baseline(w, budget_s=15.)
           if world == 1 and not args.no_cpu else None)
    line = {
        'metric': ('Dedisperse->Power->Fold complex Gsamples/s' if folding
                   else 'Dedisperse->Channelize->Power->Integrate complex '
                   'Gsamples/s'),
        'value': value, 'unit': 'Gsamples/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c64 (fp32)',
        'data': 'synthetic (NoiseGenerator, Philox)',
        'config': {'workload': f"{args.workload}: {w['desc']}",
                   'frames_per_step': w['frames'], 'fft_length': N,
                   'samples_per_frame': spf, 'series': S,
                   'input_bytes_per_step': int(host_np.nbytes),
                   'l2': 'inputs (%.2f GB per step) larger than L2'
                         % (host_np.nbytes / 1e9),
                   'sharding': 'time blocks with overlap-save halo, one per '
                               'rank; ' + ('NCCL all-reduce of the folded '
                                           'profile' if folding
                                           else 'no collective')},
        'e2e': {'value': e2e, 'unit': 'Gsamples/s',
                'h2d_bytes_per_step': int(host_np.nbytes),
                'd2h_bytes_per_step': int(out_bytes),
                'ms_per_step': ms_e2e / args.steps},
        # Supplementary: the same step fed with the block stored as 8-bit
        # (re, im) codes (as recorded baseband data are) and decoded on the
        # device; `e2e` above is the float32 stream BASELINE.json describes.
        'e2e_packed8': {'value': samples_per_step * args.steps * world
                        / (ms_e2e8 * 1e-3) / 1e9, 'unit': 'Gsamples/s',
                        'h2d_bytes_per_step': int(host8.numel()),
                        'd2h_bytes_per_step': int(out_bytes),
                        'ms_per_step': ms_e2e8 / args.steps},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': roofline,
        'chain_roofline': {
            'model_bytes_per_sample': mb,
            'achieved_gbs': value * mb, 'peak': peak * world,
            'frac': value * mb / (peak * world), 'frac_of_8TBs_nominal':
            value * mb / (8000. * world)},
        'kernels': shares,
        'cpu_baseline': cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------- CPU arm
def _cpu_unit(args):
    """One (frame, channel) unit of the chain with the oracle (numpy)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import bbt_oracle as orc
    wname, chan, seed = args
    w = WORKLOADS[wname]
    N, spf, pad_start, pad_end = framing(w)
    rng = np.random.default_rng(seed)
    npol = w['sample_shape'][-1]
    x = (rng.normal(size=(N, npol)) + 1j * rng.normal(size=(N, npol))
         ).astype('c8')
    f = np.asarray(w['freq'], float).ravel() / 1e6
    lo, hi = f - w['rate'] / 2e6, f + w['rate'] / 2e6
    fref = np.mean(lo + hi) / 2.
    plan = orc.DispersePlan(-w['dm'], f[chan % len(f)], 1, w['rate'] / 1e6,
                            True, N, 1, (npol,), reference_frequency_mhz=fref,
                            samples_per_frame=spf, fast_len=orc.next_pow2)
    plan.pad_start, plan.pad_end = pad_start, pad_end
    plan.samples_per_frame = spf
    t0 = time.perf_counter()
    pf = plan.phase_factor('c8')
    t1 = time.perf_counter()
    y = orc.disperse(x, plan, phase_factor=pf)
    if w.get('fold'):
        power = orc.power(y, axis=-1)
        coef = w['fold']['coef']

        def phase(i):
            dt = i.astype(np.float64) / w['rate']
            ph = np.full(dt.shape, coef[-1])
            for c in coef[-2::-1]:
                ph = ph * dt + c
            return ph
        orc.fold(power, np.array([0, power.shape[0]]), w['fold']['n_phase'],
                 phase)
    else:
        spectra = orc.channelize(y, w['n_chan'])
        power = orc.power(spectra, axis=-1)
        ip = orc.IntegratePlan(power.shape[0], w['rate'] / w['n_chan'],
                               w['step'])
        offsets = ip.offsets(np.arange(ip.n_out + 1))
        orc.integrate(power, offsets)
    t2 = time.perf_counter()
    return spf * npol, t2 - t1, t1 - t0


def cpu_run(w_name, cores, units):
    """Process ``units`` (frame, channel) units on ``cores`` processes."""
    import multiprocessing as mp
    jobs = [(w_name, i, 100 + i) for i in range(units)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_unit(j) for j in jobs]
    else:
        with mp.get_context('fork').Pool(cores) as pool:
            res = pool.map(_cpu_unit, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall, sum(r[1] for r in res)


def cpu_baseline(w, budget_s=15.):
    name = [k for k, v in WORKLOADS.items() if v is w][0]
    # One unit first, to size the sample.
    n, wall, busy = cpu_run(name, 1, 1)
    units = int(max(1, min(16, budget_s / max(busy, 1e-3))))
    if units > 1:
        n2, wall2, busy2 = cpu_run(name, 1, units)
        n, busy = n + n2, busy + busy2
        units += 1
    return {'value': n / busy / 1e9, 'unit': 'Gsamples/s', 'cores': 1,
            'kind': 'port',
            'sample': f'{units} (frame, channel) units of the same workload '
                      '(one N-point frame of one channel, both '
                      'polarizations) through oracle/bbt_oracle.py '
                      '(numpy.fft); chirp construction excluded'}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    w = WORKLOADS[args.workload]
    cores = len(os.sched_getaffinity(0))
    n_series_groups = max(1, int(np.prod(w['sample_shape'][:-1])))
    units = max(cores, n_series_groups)
    for _ in range(min(args.warmup, 1)):
        cpu_run(args.workload, cores, units)
    total, wall = 0, 0.
    steps = args.steps
    for _ in range(steps):
        n, t, _ = cpu_run(args.workload, cores, units)
        total += n
        wall += t
    value = total / wall / 1e9
    N, spf, _, _ = framing(w)
    sample = (f'{units} (frame, channel) units per step on {cores} processes '
              '(numpy.fft restatement of the reference; the reference needs '
              'astropy+baseband, not installable here)')
    line = {
        'impl': 'reference',
        'metric': 'Dedisperse->Channelize->Power->Integrate complex '
                  'Gsamples/s',
        'value': value, 'unit': 'Gsamples/s', 'n_gpus': int(args.gpus),
        'steps': steps, 'warmup': min(args.warmup, 1),
        'ms_per_step': wall / steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c64 (fp32)',
        'data': 'synthetic',
        'config': {'workload': f"{args.workload}: {w['desc']}",
                   'fft_length': N, 'samples_per_frame': spf},
        'cpu_baseline': {'value': value, 'unit': 'Gsamples/s',
                         'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'Gsamples/s',
                'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', default='C2', choices=sorted(WORKLOADS))
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true',
                    help='skip the CPU baseline leg (kernel A/B runs)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
