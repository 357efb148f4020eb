"""Benchmark of the coherent-dedispersion / channelization hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W]
                    [--workload C4|C2|C5|C1|C3] [--impl b200|reference]

One step = one pass of the chain over one block of F overlap-save frames per
GPU of a synthetic NoiseGenerator stream:

  C4 (default; BASELINE.json configs[3], the north-star target): (T, 2)
     complex64, 512 MHz at 8192 MHz, Dedisperse(DM=1000, N=2^24) ->
     Channelize(1024) -> Power -> Integrate(1 ms); 32 frames per GPU and step.
  C2 (configs[1]): (T, 8, 2) complex64, 8 channels of 8 MHz at 1372+8k MHz,
     Dedisperse(DM=100, N=2^20) -> Channelize(1024) -> Power -> Integrate(1 ms).
  C5 (configs[4]): the C4 stream -> Dedisperse -> Power -> Fold(512 bins,
     polynomial phase), the folded profile summed over ranks with NCCL.
  C1 (configs[0]): (T,) complex64 16 MHz at 400 MHz, Dedisperse(DM=26.8,
     N=2^21) only.
  C3 (configs[2]): real 8-bit (T, 2) at 800 MS/s -> PolyphaseFilterBank(4 taps,
     2048) -> per-channel Dedisperse -> Power.

``value``: complex source samples (time x channel x polarization; real samples
for C3) per second through the public Task API with this rank's block of the
stream resident in HBM.  ``e2e``: the same with the block in pinned host
memory, copied to the device inside the timed region (in pieces on a copy
stream, overlapped with the work on the pieces that have arrived), and the
result copied back.  Timing: CUDA events, max over ranks.

With N > 1 the job is ONE stream of N x F frames shared out in time: every
rank builds the chain on its own block of the stream (its frames plus the
overlap-save halo, `parallel.StreamBlock`) and integrates the bins its block
touches; the bins cut by a rank boundary are completed over ranks with one
small NCCL all-reduce per step (`parallel.reduce_edge_bins`; for C5 the whole
folded profile is all-reduced).  Per-GPU work is fixed: weak scaling.  The
line also carries a strong-scaling figure: a fixed stream of 64 frames (device
generated) split over the ranks the same way.

``--impl reference`` times the reference's CPU path -- its numpy arithmetic
restated in oracle/bbt_oracle.py, since the reference itself needs astropy and
baseband, which are not installable here -- on the host cores: a persistent
pool of one process per core, each holding one frame of the same stream (input
and chirp prepared outside the timed region, as the reference caches them).
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

T0 = 1289567655          # 2010-11-12T13:14:15, unix seconds
GEN_SPF = 1 << 20        # samples per frame of the noise generator

WORKLOADS = {
    'C4': dict(rate=512e6, freq=8192e6, sample_shape=(2,), dm=1000.,
               log2n=24, frames=32, n_chan=1024, step=1e-3, seed=1234567 + 4,
               desc='2pol x 512 MHz c64 -> Dedisperse(DM=1000, N=2^24) -> '
                    'Channelize(1024) -> Power -> Integrate(1 ms)'),
    'C2': dict(rate=8e6, freq=(1372e6 + 8e6 * np.arange(8)).reshape(8, 1),
               sample_shape=(8, 2), dm=100., log2n=20, frames=32,
               n_chan=1024, step=1e-3, seed=1234567 + 2,
               desc='8ch x 2pol x 8 MHz c64 -> Dedisperse(DM=100, N=2^20) -> '
                    'Channelize(1024) -> Power -> Integrate(1 ms)'),
    'C5': dict(rate=512e6, freq=8192e6, sample_shape=(2,), dm=1000.,
               log2n=24, frames=32, n_chan=None, step=None, seed=1234567 + 5,
               fold=dict(n_phase=512, coef=[0.25, 29.946923,
                                            -3.77535e-10 / 2.]),
               desc='2pol x 512 MHz c64 -> Dedisperse(DM=1000, N=2^24) -> '
                    'Power -> Fold(512 bins, polynomial phase), NCCL reduce '
                    'of the profile'),
    'C1': dict(rate=16e6, freq=400e6, sample_shape=(), dm=26.8, log2n=21,
               frames=128, n_chan=None, step=None, seed=1234567 + 1,
               only_dedisperse=True,
               desc='1ch x 16 MHz c64 -> Dedisperse(DM=26.8, N=2^21)'),
    'C3': dict(rate=800e6, freq=800e6, sample_shape=(2,), dm=100.,
               seed=1234567 + 3, pfb=dict(n_tap=4, n=2048, n_spec=1 << 17),
               desc='2pol x 800 MS/s real 8-bit -> PolyphaseFilterBank(4 x '
                    '2048) -> per-channel Dedisperse(DM=100) -> Power'),
}
METRIC = {
    'C4': 'Dedisperse->Channelize->Power->Integrate complex Gsamples/s',
    'C2': 'Dedisperse->Channelize->Power->Integrate complex Gsamples/s',
    'C5': 'Dedisperse->Power->Fold complex Gsamples/s',
    'C1': 'Dedisperse complex Gsamples/s',
    'C3': 'PolyphaseFilterBank->Dedisperse->Power real Gsamples/s',
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
    if w.get('pfb'):
        return 13.6
    N, spf, _, _ = framing(w)
    eff = spf / N
    if w.get('only_dedisperse'):
        return 48. / eff
    if w.get('fold'):
        return 48. / eff + 8.
    per_bin = w['step'] * w['rate'] / w['n_chan']
    return 48. / eff + 8. + 8. / per_bin


def config_of(name, w):
    """The ``config`` object, identical for both arms."""
    cfg = {'workload': f"{name}: {w['desc']}"}
    if not w.get('pfb'):
        N, spf, _, _ = framing(w)
        cfg.update(frames_per_gpu_and_step=w['frames'], fft_length=N,
                   samples_per_frame=spf)
    cfg['series'] = int(np.prod(w['sample_shape'], dtype=np.int64))
    cfg['l2'] = 'every kernel streams its whole block (GBs) per step: ' \
                'inputs larger than L2'
    cfg['sharding'] = ('one stream of n_gpus x frames_per_gpu_and_step frames '
                       'shared out in time (overlap-save halo per rank); '
                       'NCCL all-reduce of '
                       + ('the folded profile' if w.get('fold') else
                          'the integration bins cut by rank boundaries'))
    return cfg


class ClocksSampler:
    """nvidia-smi polled every 50 ms in the background; every line is kept
    with the time it arrived, so the samples taken under load can be picked."""

    def __init__(self, device_index):
        import threading
        cmd = ['nvidia-smi', '-i', str(device_index),
               '--query-gpu=clocks.sm,clocks.max.sm,'
               'clocks_event_reasons.hw_slowdown,'
               'clocks_event_reasons.hw_thermal_slowdown,'
               'clocks_event_reasons.sw_thermal_slowdown,'
               'clocks_event_reasons.sw_power_cap',
               '--format=csv,noheader,nounits', '-lms', '50']
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

    def summary(self, windows):
        """Median SM clock and throttle reasons of the samples that fall in
        one of the (t0, t1) windows."""
        if self.proc is None:
            return None
        time.sleep(0.06)
        self.proc.terminate()
        sm, smax, reasons = [], 0., set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for t, line in list(self.samples):
            if not any(a <= t <= b for a, b in windows):
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
                'reasons': sorted(reasons), 'samples': len(sm),
                'window': 'samples taken during the timed regions '
                          '(HBM-resident, end-to-end and strong-scaling '
                          'loops)'}


# --------------------------------------------------------------- input data
def noise_block(w, first, count, out=None):
    """Samples [first, first + count) of the workload's NoiseGenerator stream
    (host; the generator's frames are drawn on a pool of threads)."""
    import baseband_tasks_b200 as bt
    from concurrent.futures import ThreadPoolExecutor
    shape = tuple(w['sample_shape'])
    f0, f1 = first // GEN_SPF, -(-(first + count) // GEN_SPF)
    if out is None:
        out = np.empty((count,) + shape, np.complex64)

    def draw(f):
        nh = bt.NoiseGenerator(((f + 1) * GEN_SPF,) + shape, bt.Time(T0),
                               w['rate'], samples_per_frame=GEN_SPF,
                               dtype='c8', seed=w['seed'])
        nh.seek(f * GEN_SPF)
        frame = nh.read(GEN_SPF)
        a = max(first, f * GEN_SPF)
        b = min(first + count, (f + 1) * GEN_SPF)
        out[a - first:b - first] = frame[a - f * GEN_SPF:b - f * GEN_SPF]

    threads = max(1, min(32, len(os.sched_getaffinity(0))))
    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(draw, range(f0, f1)))
    return out


class Job:
    """One rank's share of a stream of ``world * frames`` frames."""

    def __init__(self, w, rank, world, frames=None):
        from baseband_tasks_b200 import parallel
        self.w = w
        self.N, self.spf, self.pad_start, self.pad_end = framing(w)
        self.pad = self.pad_start + self.pad_end
        self.S = int(np.prod(w['sample_shape'], dtype=np.int64))
        per_rank = w['frames'] if frames is None else None
        self.n_frames = (per_rank * world if frames is None else frames)
        self.n_total = self.n_frames * self.spf + self.pad
        unit = w['n_chan'] or 1
        self.unit = unit
        self.plans = [parallel.block_plan(self.n_frames, self.spf, self.pad,
                                          unit, r, world)
                      for r in range(world)]
        self.first, self.last, self.in0, self.in1 = self.plans[rank]
        self.rank, self.world = rank, world
        # Source samples this rank turns into output per step.
        self.samples = (self.last - self.first) * self.S
        self.total_samples = sum((p[1] - p[0]) * self.S for p in self.plans)

    def chain(self, data):
        """The task chain on this rank's block ``data`` (host or device)."""
        import baseband_tasks_b200 as bt
        from baseband_tasks_b200 import parallel
        w = self.w
        src = parallel.StreamBlock(
            data, self.in0, self.n_total, bt.Time(T0), w['rate'],
            samples_per_frame=GEN_SPF, frequency=w['freq'], sideband=1,
            **({'polarization': np.array(['X', 'Y'])}
               if w['sample_shape'] else {}))
        dd = bt.Dedisperse(src, w['dm'], samples_per_frame=self.spf)
        assert dd._ih_samples_per_frame == self.N, dd._ih_samples_per_frame
        if w.get('only_dedisperse'):
            return dd
        if w.get('fold'):
            poly = bt.PolynomialPhase(w['fold']['coef'], bt.Time(T0))
            return bt.Fold(bt.Power(dd), w['fold']['n_phase'], poly,
                           average=False)
        return bt.Integrate(bt.Power(bt.Channelize(dd, w['n_chan'])),
                            w['step'], average=False)

    def runner(self, chain):
        """``run(lo, hi)``: this rank's result for the part [lo, hi) of its
        output range (in units of ``self.unit`` output samples of the
        dedispersion).  For the whole range the result is complete: bins
        shared with other ranks are reduced over ranks (NCCL).  For a part
        (the end-to-end arm works on the pieces of the block as they arrive)
        partial sums are added into one accumulator, reduced with the last
        part."""
        from baseband_tasks_b200 import parallel
        import torch
        w = self.w
        lo_all, hi_all = self.first // self.unit, self.last // self.unit
        if w.get('only_dedisperse'):
            def run(lo, hi):
                chain.seek(lo)
                return chain.read_device(hi - lo)
            return run
        if w.get('fold'):
            def run(lo, hi):
                chain.seek(0)
                sums, counts = chain.read_sums(within=(lo, hi))
                parallel.reduce_sums(sums, counts)
                return sums
            return run
        edges = np.asarray(chain._get_offsets(np.arange(chain.shape[0] + 1)))
        bins = [parallel.bin_range(edges, p[0] // self.unit, p[1] // self.unit)
                for p in self.plans]
        b_lo, b_hi = bins[self.rank]
        acc = {}

        def run(lo, hi):
            b0, b1 = parallel.bin_range(edges, lo, hi)
            chain.seek(b0)
            sums, counts = chain.read_sums(b1 - b0, within=(lo, hi))
            if (lo, hi) != (lo_all, hi_all):
                if lo == lo_all:
                    acc['s'] = torch.zeros((b_hi - b_lo,) + sums.shape[1:],
                                           dtype=sums.dtype,
                                           device=sums.device)
                    acc['c'] = torch.zeros((b_hi - b_lo,) + counts.shape[1:],
                                           dtype=counts.dtype,
                                           device=counts.device)
                acc['s'][b0 - b_lo:b1 - b_lo] += sums
                acc['c'][b0 - b_lo:b1 - b_lo] += counts
                if hi != hi_all:
                    return None
                sums, counts = acc['s'], acc['c']
            parallel.reduce_edge_bins(sums, counts, bins[self.rank], bins)
            return sums
        return run


def bind_near_gpu(local):
    """Run this process (and allocate its pinned host block) on the NUMA
    node the GPU hangs off: with eight ranks copying 6.6 GB per step each,
    host memory on the wrong node halves the PCIe rate.  Returns the node or
    None when the topology cannot be read."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f'/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node'
        node = int(open(path).read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f'/sys/devices/system/node/node{node}/cpulist'
                         ).read().strip().split(','):
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def run_b200(args):
    import torch
    import torch.distributed as dist
    import baseband_tasks_b200 as bt
    from baseband_tasks_b200 import _cabi

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    numa_node = bind_near_gpu(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    lib = _cabi.lib()
    name = args.workload
    w = WORKLOADS[name]
    if w.get('pfb'):
        return run_c3(args, rank, world, local)
    job = Job(w, rank, world)
    # One launch per kernel and step: blocks of a whole step's frames.
    bt.base.BLOCK_BYTES = max(bt.base.BLOCK_BYTES,
                              w['frames'] * job.spf * job.S * 8)

    n_block = job.in1 - job.in0
    host = torch.empty((n_block,) + tuple(w['sample_shape']),
                       dtype=torch.complex64).pin_memory()
    noise_block(w, job.in0, n_block, out=host.numpy())
    dev_in = host.to('cuda', non_blocking=True)
    torch.cuda.synchronize()

    chain = job.chain(dev_in)
    run = job.runner(chain)
    lo_u, hi_u = job.first // job.unit, job.last // job.unit

    def step_resident():
        return run(lo_u, hi_u)

    # End to end: the block goes from pinned host memory to the device in
    # pieces of whole frames on a copy stream; the chain works on the output
    # the pieces that have arrived determine (successive reads through the
    # public API), and every partial result is copied back.
    # Two staging buffers, each with a chain of its own, take turns: the
    # input of step k+1 arrives while step k is still being processed and
    # its result copied back.
    dev_stages = [torch.empty_like(dev_in) for _ in range(2)]
    chains_e2e = [job.chain(d) for d in dev_stages]
    runs_e2e = [job.runner(c) for c in chains_e2e]
    stage_free = [None, None]     # recorded when a step is done with a buffer
    n_e2e = [0]
    copy_stream = torch.cuda.Stream()
    d2h_stream = torch.cuda.Stream()
    n_pieces = 4 if not w.get('fold') else 1
    f_first = job.in0 // job.spf
    f_count = (n_block - job.pad) // job.spf
    cuts = [f_first + (f_count * i) // n_pieces for i in range(n_pieces + 1)]
    pieces, reads = [], []
    done = lo_u
    for i in range(n_pieces):
        a = (cuts[i] - f_first) * job.spf + (job.pad if i else 0)
        b = (cuts[i + 1] - f_first) * job.spf + job.pad
        pieces.append((a, b))
        upto = hi_u if i == n_pieces - 1 else min(
            hi_u, max(done, (cuts[i + 1] * job.spf) // job.unit))
        reads.append((done, upto))
        done = upto
    out_host = {}

    def step_e2e(src_host=None, decode=None):
        main = torch.cuda.current_stream()
        k = n_e2e[0] & 1
        n_e2e[0] += 1
        dev_stage, run_e2e = dev_stages[k], runs_e2e[k]
        raw = decode[0][k] if decode is not None else None
        if stage_free[k] is not None:      # the step before last is done with it
            copy_stream.wait_event(stage_free[k])
        events = []
        with torch.cuda.stream(copy_stream):
            for a, b in pieces:
                if decode is None:
                    dev_stage[a:b].copy_(host[a:b], non_blocking=True)
                else:
                    raw[a:b].copy_(src_host[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                events.append(ev)
        res = None
        for i, ((a, b), (r0, r1)) in enumerate(zip(pieces, reads)):
            main.wait_event(events[i])
            if decode is not None:
                lib.check(lib.bbt_decode_exec(
                    ctypes.c_void_p(raw[a:b].data_ptr()),
                    ctypes.c_void_p(dev_stage[a:b].data_ptr()),
                    ctypes.c_void_p(decode[1].data_ptr()),
                    (b - a) * 2 * job.S, 8, ctypes.c_void_p(main.cuda_stream)))
            if r1 <= r0 and i < n_pieces - 1:
                continue
            res = run_e2e(lo_u, hi_u) if n_pieces == 1 else run_e2e(r0, r1)
            if res is None:            # partial sums, kept on the device
                continue
            key = (i, tuple(res.shape))
            if key not in out_host:
                out_host[key] = torch.empty(res.shape, dtype=res.dtype,
                                            pin_memory=True)
            ev = torch.cuda.Event()
            ev.record(main)
            d2h_stream.wait_event(ev)
            with torch.cuda.stream(d2h_stream):
                out_host[key].copy_(res, non_blocking=True)
                res.record_stream(d2h_stream)
        stage_free[k] = torch.cuda.Event()
        stage_free[k].record(main)
        return res

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    windows = []

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.monotonic()
        e0.record()
        for _ in range(steps):
            fn()
        # (the last results' copies to the host are part of the region)
        torch.cuda.current_stream().wait_stream(d2h_stream)
        e1.record()
        barrier()
        windows.append((t0, time.monotonic()))
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    warm = max(args.warmup, 3)
    sampler = ClocksSampler(local) if rank == 0 else None
    for _ in range(warm):
        res = step_resident()
    out_bytes = res.numel() * res.element_size()
    l0 = lib.bbt_launch_count()
    ms = timed(step_resident, args.steps)
    launches = lib.bbt_launch_count() - l0

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    d2h_bytes = sum(t.numel() * t.element_size() for t in out_host.values())
    out_host.clear()

    # The same block as recorded baseband data are stored: 8-bit (re, im)
    # codes, a quarter of the bytes over PCIe, decoded on the device.
    levels8 = bt.payload_levels(8) / 30.
    host8 = torch.from_numpy(bt.encode_payload(
        host.numpy(), 8, levels8).reshape(n_block, -1)).pin_memory()
    dev8 = [torch.empty_like(host8, device='cuda') for _ in range(2)]
    d_levels8 = torch.from_numpy(levels8).cuda()
    for _ in range(2):
        step_e2e(host8, (dev8, d_levels8))
    ms_e2e8 = timed(lambda: step_e2e(host8, (dev8, d_levels8)), args.steps)
    out_host.clear()
    del dev8, host8

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
        kname, count, total = line.split()
        kernels[kname] = (int(count), float(total))

    # Strong scaling: a fixed stream of 64 frames split over the ranks.
    strong = None
    if not w.get('only_dedisperse') and args.strong:
        del dev_stages, chains_e2e, runs_e2e
        torch.cuda.empty_cache()
        sjob = Job(w, rank, world, frames=64)
        g = torch.Generator(device='cuda').manual_seed(w['seed'] + rank)
        sdata = torch.view_as_complex(torch.randn(
            (sjob.in1 - sjob.in0,) + tuple(w['sample_shape']) + (2,),
            device='cuda', generator=g))
        srun = sjob.runner(sjob.chain(sdata))
        s_lo, s_hi = sjob.first // sjob.unit, sjob.last // sjob.unit
        for _ in range(3):
            srun(s_lo, s_hi)
        ms_s = timed(lambda: srun(s_lo, s_hi), args.steps)
        strong = {'value': sjob.total_samples * args.steps / (ms_s * 1e-3)
                  / 1e9, 'unit': 'Gsamples/s', 'scaling': 'strong',
                  'frames_total': sjob.n_frames,
                  'ms_per_step': ms_s / args.steps,
                  'data': 'device-generated noise (not part of parity)'}
        del sdata, srun

    clocks = sampler.summary(windows) if sampler is not None else None
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
    # Dominant kernel: algorithmic bytes per launch (DESIGN.md section 4:
    # bytes per FFT point x the points one launch processes).
    frames_run = (job.in1 - job.in0 - job.pad) // job.spf
    points = frames_run * job.N * job.S
    out_samples = frames_run * job.spf * job.S
    alg = {'dd_col_fwd': 16. * points, 'dd_row': 16. * points,
           'dd_col_inv': 8. * points + 8. * out_samples,
           'dd_small': 16. * points,
           'chanpow_integrate': 8. * job.samples + out_bytes,
           'fold': 8. * job.samples}
    top = max((k for k in kernels if k in alg),
              key=lambda k: kernels[k][1], default=None)
    roofline = None
    if top:
        count, total_ms = kernels[top]
        per_launch_ms = total_ms / count
        launches_per_step = count / args.steps
        bytes_per_launch = alg[top] / launches_per_step
        achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
                per_frame = json.load(f).get(name, {}).get(top)
            if per_frame is not None:
                traffic = per_frame * frames_run / launches_per_step
        except OSError:
            pass
        roofline = {'bound': 'hbm', 'kernel': top, 'achieved': achieved,
                    'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                    'frac': achieved / peak, 'traffic': traffic,
                    'ms_per_launch': per_launch_ms,
                    'frames_per_launch': frames_run / launches_per_step,
                    'algorithmic_bytes_per_launch': bytes_per_launch}
    total_kernel_ms = sum(v[1] for v in kernels.values()) or 1.
    shares = {k: {'launches': v[0], 'ms_per_launch': v[1] / v[0],
                  'share': v[1] / total_kernel_ms}
              for k, v in sorted(kernels.items())}

    value = job.total_samples * args.steps / (ms * 1e-3) / 1e9
    e2e = job.total_samples * args.steps / (ms_e2e * 1e-3) / 1e9
    mb = model_bytes(w)
    cpu = (cpu_baseline(name, budget_s=15.)
           if world == 1 and not args.no_cpu else None)
    h2d = int(n_block * job.S * 8)
    line = {
        'metric': METRIC[name], 'value': value, 'unit': 'Gsamples/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': warm,
        'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c64 (fp32)',
        'data': 'synthetic (NoiseGenerator, Philox)',
        'config': config_of(name, w),
        'e2e': {'value': e2e, 'unit': 'Gsamples/s',
                'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': int(d2h_bytes),
                'ms_per_step': ms_e2e / args.steps,
                'h2d_gbs_per_rank': h2d / (ms_e2e / args.steps * 1e-3) / 1e9,
                'host_numa_node': numa_node},
        # Supplementary: the same step fed with the block stored as 8-bit
        # (re, im) codes (as recorded baseband data are) and decoded on the
        # device; `e2e` above is the float32 stream BASELINE.json describes.
        'e2e_packed8': {'value': job.total_samples * args.steps
                        / (ms_e2e8 * 1e-3) / 1e9, 'unit': 'Gsamples/s',
                        'h2d_bytes_per_step': h2d // 4,
                        'd2h_bytes_per_step': int(d2h_bytes),
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
        'strong': strong,
        'cpu_baseline': cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_c3(args, rank, world, local):
    """configs[2]: PFB -> per-channel Dedisperse -> Power on a block of raw
    8-bit samples (channels are independent: every rank takes its own block,
    no collective)."""
    import torch
    import torch.distributed as dist
    import baseband_tasks_b200 as bt
    from baseband_tasks_b200 import _cabi
    lib = _cabi.lib()
    w = WORKLOADS['C3']
    p = w['pfb']
    n = p['n_spec'] * p['n']
    rng = np.random.default_rng(w['seed'] + rank)
    host = torch.from_numpy(np.clip(np.round(
        rng.normal(size=(n, 2)) * 20), -127, 127).astype(np.int8)).pin_memory()
    dev = host.to('cuda')
    response = bt.sinc_hamming(p['n_tap'], p['n'])

    def chain_on(data):
        src = bt.ArrayStream(data, bt.Time(T0), w['rate'],
                             samples_per_frame=1 << 20, frequency=w['freq'],
                             sideband=-1, polarization=np.array(['X', 'Y']))
        pfb = bt.PolyphaseFilterBank(src, response)
        dd = bt.Dedisperse(pfb, w['dm'], reference_frequency=pfb.frequency)
        return bt.Power(dd)

    chain = chain_on(dev)

    def step():
        chain.seek(0)
        return chain.read_device()

    stage = torch.empty_like(dev)
    chain2 = chain_on(stage)
    out_host = []

    # The result (2.1 GB of powers per step) takes four times as long over
    # PCIe as the 8-bit input: it is copied back on a stream of its own, into
    # alternating pinned buffers, while the next step's input arrives and is
    # processed; the timed region ends when the last copy has landed.
    d2h_stream = torch.cuda.Stream()
    n_e2e = [0]

    def step_e2e():
        main = torch.cuda.current_stream()
        stage.copy_(host, non_blocking=True)
        chain2.seek(0)
        res = chain2.read_device()
        while len(out_host) < 2:
            out_host.append(torch.empty(res.shape, dtype=res.dtype,
                                        pin_memory=True))
        ev = torch.cuda.Event()
        ev.record(main)
        d2h_stream.wait_event(ev)
        with torch.cuda.stream(d2h_stream):
            out_host[n_e2e[0] & 1].copy_(res, non_blocking=True)
            res.record_stream(d2h_stream)
        n_e2e[0] += 1
        return res

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    windows = []

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.monotonic()
        e0.record()
        for _ in range(steps):
            fn()
        torch.cuda.current_stream().wait_stream(d2h_stream)
        e1.record()
        barrier()
        windows.append((t0, time.monotonic()))
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    warm = max(args.warmup, 3)
    sampler = ClocksSampler(local) if rank == 0 else None
    for _ in range(warm):
        res = step()
    out_bytes = res.numel() * res.element_size()
    l0 = lib.bbt_launch_count()
    ms = timed(step, args.steps)
    launches = lib.bbt_launch_count() - l0
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    lib.bbt_profile_enable(1)
    barrier()
    for _ in range(args.steps):
        step()
    barrier()
    lib.bbt_profile_enable(0)
    buf = ctypes.create_string_buffer(1 << 16)
    lib.check(lib.bbt_profile_report(buf, len(buf)))
    kernels = {}
    for line in buf.value.decode().splitlines():
        kname, count, total = line.split()
        kernels[kname] = (int(count), float(total))
    clocks = sampler.summary(windows) if sampler is not None else None
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
    samples = n * 2                      # real samples, both polarizations
    n_out = res.shape[0] * 1025 * 2      # dedispersed complex values
    # Algorithmic bytes per launch: the filter bank reads 1 B
    # and writes 8 B x 1025/2048 per real sample; the single-pass
    # dedispersion reads and writes 8 B per point; Power reads 8, writes 8.
    alg = {'pfb': samples * (1. + 8. * 1025 / 2048),
           'dd_small': 16. * kernels.get('dd_small', (1, 0))[0] / args.steps
           and 16. * n_out / 0.879, 'power': 16. * n_out}
    top = max((k for k in kernels if k in alg), key=lambda k: kernels[k][1])
    count, total_ms = kernels[top]
    per_launch_ms = total_ms / count
    bytes_per_launch = alg[top] / (count / args.steps)
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    total_kernel_ms = sum(v[1] for v in kernels.values()) or 1.
    value = samples * world * args.steps / (ms * 1e-3) / 1e9
    mb = model_bytes(w)
    line = {
        'metric': METRIC['C3'], 'value': value, 'unit': 'Gsamples/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': warm,
        'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c64 (fp32)',
        'data': 'synthetic (8-bit rounded normal noise)',
        'config': config_of('C3', w),
        'e2e': {'value': samples * world * args.steps / (ms_e2e * 1e-3) / 1e9,
                'unit': 'Gsamples/s', 'h2d_bytes_per_step': int(samples),
                'd2h_bytes_per_step': int(out_bytes),
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches), 'clocks': clocks,
        'roofline': {'bound': 'hbm', 'kernel': top, 'achieved': achieved,
                     'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': None, 'ms_per_launch': per_launch_ms,
                     'algorithmic_bytes_per_launch': bytes_per_launch},
        'chain_roofline': {'model_bytes_per_sample': mb,
                           'achieved_gbs': value * mb, 'peak': peak * world,
                           'frac': value * mb / (peak * world)},
        'kernels': {k: {'launches': v[0], 'ms_per_launch': v[1] / v[0],
                        'share': v[1] / total_kernel_ms}
                    for k, v in sorted(kernels.items())},
        'cpu_baseline': None,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------- CPU arm
_REF = {}


def _ref_init(name, counter, base_frame):
    """Worker set-up, outside the timed region: one frame of the workload's
    stream, the dedispersion plan and its cached chirp (the reference builds
    phase_factor once, dispersion.py:115 is a lazyproperty)."""
    os.environ['OMP_NUM_THREADS'] = '1'
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import bbt_oracle as orc
    w = WORKLOADS[name]
    with counter.get_lock():
        index = counter.value
        counter.value += 1
    N, spf, pad_start, pad_end = framing(w)
    shape = tuple(w['sample_shape'])
    # Frame `index` of the stream: input samples [index*spf, index*spf + N).
    first = (base_frame + index) * spf
    f0, f1 = first // GEN_SPF, -(-(first + N) // GEN_SPF)
    x = np.concatenate([orc.noise_frame(w['seed'], f * GEN_SPF, GEN_SPF,
                                        shape, 'c8') for f in range(f0, f1)])
    x = np.ascontiguousarray(x[first - f0 * GEN_SPF:first - f0 * GEN_SPF + N])
    f = np.asarray(w['freq'], float) / 1e6
    plan = orc.DispersePlan(-w['dm'], f, 1, w['rate'] / 1e6, True, N, N,
                            shape, samples_per_frame=spf,
                            fast_len=orc.next_pow2)
    assert (plan.N, plan.pad_start, plan.pad_end) == (N, pad_start, pad_end)
    _REF.update(orc=orc, w=w, x=x, plan=plan, pf=plan.phase_factor('c8'),
                spf=spf)


def _ref_step(_):
    """One frame through the oracle chain (numpy), as the reference's tasks
    would process it."""
    orc, w, x, plan = _REF['orc'], _REF['w'], _REF['x'], _REF['plan']
    t0 = time.perf_counter()
    y = orc.disperse(x, plan, phase_factor=_REF['pf'])
    if w.get('only_dedisperse'):
        pass
    elif w.get('fold'):
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
        orc.integrate(power, ip.offsets(np.arange(ip.n_out + 1)))
    return _REF['spf'] * int(np.prod(w['sample_shape'], dtype=np.int64)), \
        time.perf_counter() - t0


class RefPool:
    """Persistent pool of ``cores`` processes, one frame each."""

    def __init__(self, name, cores):
        import multiprocessing as mp
        ctx = mp.get_context('fork')
        self.cores = cores
        self.pool = ctx.Pool(cores, initializer=_ref_init,
                             initargs=(name, ctx.Value('i', 0), 0))
        self.step()                      # every worker is set up

    def step(self):
        t0 = time.perf_counter()
        res = self.pool.map(_ref_step, range(self.cores), chunksize=1)
        wall = time.perf_counter() - t0
        return sum(r[0] for r in res), wall, sum(r[1] for r in res)

    def close(self):
        self.pool.close()
        self.pool.join()


def usable_cores(w):
    """Host threads the CPU arm uses: all of them, unless memory (about 12
    frame-sized arrays per process) is the tighter limit."""
    cores = len(os.sched_getaffinity(0))
    N = 1 << w['log2n']
    per_proc = 12 * N * int(np.prod(w['sample_shape'], dtype=np.int64)) * 8
    try:
        import psutil
        avail = psutil.virtual_memory().available
        cores = max(1, min(cores, int(0.6 * avail // per_proc)))
    except ImportError:
        pass
    return cores


def cpu_baseline(name, budget_s=15.):
    """The oracle chain on ONE host core, on a bounded sample."""
    w = WORKLOADS[name]
    pool = RefPool(name, 1)              # set-up and one warm-up unit
    n, wall, busy = pool.step()
    units = int(max(1, min(8, budget_s / max(busy, 1e-3))))
    for _ in range(units - 1):
        n2, _, busy2 = pool.step()
        n, busy = n + n2, busy + busy2
    pool.close()
    return {'value': n / busy / 1e9, 'unit': 'Gsamples/s', 'cores': 1,
            'kind': 'port',
            'sample': f'{units} frame(s) of the same stream (one N-point '
                      'frame, all series) through oracle/bbt_oracle.py '
                      '(numpy.fft); input and chirp prepared outside the '
                      'timed region'}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    name = args.workload
    w = WORKLOADS[name]
    if w.get('pfb'):
        print(json.dumps({'impl': 'reference', 'unavailable':
                          'the CPU arm covers the dedispersion chains'}))
        return
    cores = usable_cores(w)
    pool = RefPool(name, cores)          # set-up + one untimed step
    warm = max(args.warmup, 1)
    for _ in range(warm - 1):
        pool.step()
    total, wall = 0, 0.
    for _ in range(args.steps):
        n, t, _ = pool.step()
        total += n
        wall += t
    pool.close()
    value = total / wall / 1e9
    sample = (f'{cores} frames per step, one per process on {cores} host '
              'threads (numpy.fft restatement of the reference; the '
              'reference needs astropy+baseband, not installable here); the '
              'GPU arm\'s step is frames_per_gpu_and_step frames of the same '
              'stream')
    line = {
        'impl': 'reference', 'metric': METRIC[name], 'value': value,
        'unit': 'Gsamples/s', 'n_gpus': int(args.gpus), 'steps': args.steps,
        'warmup': warm, 'ms_per_step': wall / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'c64 (fp32)', 'data': 'synthetic (NoiseGenerator, Philox)',
        'config': config_of(name, w),
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
    ap.add_argument('--workload', default='C4', choices=sorted(WORKLOADS))
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true',
                    help='skip the CPU baseline leg (kernel A/B runs)')
    ap.add_argument('--no-strong', dest='strong', action='store_false',
                    help='skip the strong-scaling leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
