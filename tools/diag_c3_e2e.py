"""Where a C3 end-to-end step spends its time: host time stamps and CUDA
events around the input copy, the chain and the copy of the result."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import baseband_tasks_b200 as bt

w = bench.WORKLOADS['C3']
p = w['pfb']
n = p['n_spec'] * p['n']
rng = np.random.default_rng(1)
host = torch.from_numpy(np.clip(np.round(rng.normal(size=(n, 2)) * 20),
                                -127, 127).astype(np.int8)).pin_memory()
stage = torch.empty_like(host, device='cuda')
response = bt.sinc_hamming(p['n_tap'], p['n'])
src = bt.ArrayStream(stage, bt.Time(bench.T0), w['rate'],
                     samples_per_frame=1 << 20, frequency=w['freq'],
                     sideband=-1, polarization=np.array(['X', 'Y']))
pfb = bt.PolyphaseFilterBank(src, response)
dd = bt.Dedisperse(pfb, w['dm'], reference_frequency=pfb.frequency)
chain = bt.Power(dd)
d2h = torch.cuda.Stream()
outs = []
main = torch.cuda.current_stream()
rows = []
E = lambda: torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
base = E(); base.record()
t00 = time.monotonic()
for k in range(8):
    t0 = time.monotonic()
    e_a = E(); e_a.record(main)
    stage.copy_(host, non_blocking=True)
    e_b = E(); e_b.record(main)
    t1 = time.monotonic()
    chain.seek(0)
    res = chain.read_device()
    t2 = time.monotonic()
    e_c = E(); e_c.record(main)
    while len(outs) < 2:
        outs.append(torch.empty(res.shape, dtype=res.dtype, pin_memory=True))
    d2h.wait_event(e_c)
    with torch.cuda.stream(d2h):
        e_d = E(); e_d.record(d2h)
        outs[k & 1].copy_(res, non_blocking=True)
        e_e = E(); e_e.record(d2h)
        res.record_stream(d2h)
    t3 = time.monotonic()
    rows.append((t0 - t00, t1 - t00, t2 - t00, t3 - t00, e_a, e_b, e_c, e_d, e_e))
torch.cuda.synchronize()
for r in rows:
    print('host: start %.1f copy-enq %.1f chain-enq %.1f d2h-enq %.1f | gpu: h2d %.1f-%.1f chain-end %.1f d2h %.1f-%.1f' % (
        tuple(1e3 * x for x in r[:4]) + tuple(base.elapsed_time(e) for e in r[4:])))
