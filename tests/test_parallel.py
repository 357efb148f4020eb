"""Time-block sharding over ranks (world_size 2, gloo, CPU).

Each rank runs Dedisperse -> Power -> Fold on its own block of frames (with the
overlap-save halo) on the host-thread emulation of the kernels; the profile
sums and counts are reduced with torch.distributed and compared with the
single-process result and with the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

N, RATE, DM, N_PHASE = 24000, 1e6, 3., 32
SPF = 4096 - 923   # N = 4096 with the 460 + 463 samples of padding
COEF = [0.1, 29.946923, -3.77535e-10 / 2]


def _setup():
    for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import backend as _b
    import baseband_tasks_b200 as bt
    from baseband_tasks_b200 import _cabi
    _cabi._LIB = _cabi.CABI(_b.build_emu())
    _cabi._DEVICE = torch.device('cpu')
    return bt


def _data():
    rng = np.random.default_rng(7)
    return (rng.normal(size=(N, 2)) + 1j * rng.normal(size=(N, 2))).astype('c8')


def _chain(bt, src, t_ref):
    dd = bt.Dedisperse(src, DM, samples_per_frame=SPF)
    pw = bt.Power(dd)
    poly = bt.PolynomialPhase(COEF, t_ref)
    return dd, bt.Fold(pw, N_PHASE, poly, average=False)


def _source(bt, x):
    return bt.ArrayStream(x, bt.Time(1289567655), RATE, samples_per_frame=1000,
                          frequency=300e6, sideband=1,
                          polarization=np.array(['X', 'Y']))


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    bt = _setup()
    from baseband_tasks_b200 import parallel
    assert parallel.init('gloo') == (rank, world)
    src = _source(bt, _data())
    probe, _ = _chain(bt, src, src.start_time)
    assert probe.samples_per_frame == SPF
    pad = probe._ih_samples_per_frame - probe.samples_per_frame
    block, f0, f1 = parallel.shard_frames(src, SPF, pad, rank, world)
    dd, fold = _chain(bt, block, src.start_time)
    assert dd.shape[0] == (f1 - f0) * SPF
    # The block's samples are numbered on the grid of the whole stream.
    assert fold.phase.grid(fold.ih) == (0., f0 * SPF + dd._pad_start)
    assert abs((dd.start_time - probe.start_time) - f0 * SPF / RATE) < 1e-12
    sums, counts = fold.read_sums()
    local = (sums.clone().numpy(), counts.clone().numpy())
    parallel.reduce_sums(sums, counts)
    avg = parallel.average(sums, counts)
    np.savez(os.path.join(out_dir, f'rank{rank}.npz'), sums=sums.numpy(),
             counts=counts.numpy(), avg=avg.numpy(), lsum=local[0],
             lcount=local[1], f=np.array([f0, f1]))
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


def test_fold_sharded_over_two_ranks(tmp_path):
    import torch.multiprocessing as mp
    import bbt_oracle as orc
    import backend
    backend.build_emu()
    world = 2
    port = 29500 + os.getpid() % 1000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world,
             join=True)
    r = [np.load(tmp_path / f'rank{i}.npz') for i in range(world)]
    # Both ranks hold the same reduced result.
    np.testing.assert_array_equal(r[0]['sums'], r[1]['sums'])
    np.testing.assert_array_equal(r[0]['counts'], r[1]['counts'])
    assert r[0]['f'][1] == r[1]['f'][0] and r[0]['f'][0] == 0
    n_frames = int(r[1]['f'][1])
    # Single process, same frames.
    bt = _setup()
    x = _data()
    src = _source(bt, x)
    dd, fold = _chain(bt, src, src.start_time)
    n_out = n_frames * SPF
    pw = fold.ih
    one = bt.Fold(pw[:n_out], N_PHASE, fold.phase, average=False)
    sums1, counts1 = one.read_sums()
    np.testing.assert_array_equal(r[0]['counts'], counts1.numpy())
    np.testing.assert_allclose(r[0]['sums'], sums1.numpy(), rtol=1e-5,
                               atol=1e-5 * np.abs(sums1.numpy()).max())
    assert counts1.sum() == n_out
    # Oracle on the dedispersed stream.
    op = orc.DispersePlan(-DM, 300., 1, 1., True, N, 1000, (2,),
                          samples_per_frame=SPF, fast_len=orc.next_pow2)
    y = orc.disperse(x, op)[:n_out]
    power = orc.power(y, axis=-1)
    poly = fold.phase
    # Phases count samples on the grid of the source stream, whatever block
    # a rank was given: bins are bit-exact under sharding.
    i_ref, i_0 = poly.grid(pw)
    assert i_0 == dd._pad_start and i_ref == 0.
    want, wcount = orc.fold(power, np.array([0, n_out]), N_PHASE,
                            lambda i: poly.of_index(i + i_0, i_ref, RATE))
    got_cnt = r[0]['counts'].reshape(wcount.shape[:2])
    np.testing.assert_array_equal(got_cnt, wcount[..., 0])
    np.testing.assert_allclose(r[0]['sums'].reshape(want.shape), want,
                               rtol=1e-5, atol=1e-5 * np.abs(want).max())
    with np.errstate(invalid='ignore'):
        np.testing.assert_allclose(r[0]['avg'],
                                   r[0]['sums'] / r[0]['counts'][..., None],
                                   rtol=1e-6)


# ---------------------------------------------------------------------------
# Dedisperse -> Channelize -> Power -> Integrate, shared out in time: every
# rank builds the chain on its own block of the stream (parallel.StreamBlock),
# integrates the bins its block touches and the bins cut by a rank boundary
# are completed with one small all-reduce.
N_CHAN, STEP = 32, 17      # 17 spectra per bin: boundaries fall mid-bin


def _chain_integrate(bt, src):
    dd = bt.Dedisperse(src, DM, samples_per_frame=SPF)
    return dd, bt.Integrate(bt.Power(bt.Channelize(dd, N_CHAN)), STEP,
                            average=False)


def _worker_integrate(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1',
                      MASTER_PORT=str(port))
    bt = _setup()
    from baseband_tasks_b200 import parallel
    assert parallel.init('gloo') == (rank, world)
    x = _data()
    whole = _source(bt, x)
    probe, _ = _chain_integrate(bt, whole)
    pad = probe._ih_samples_per_frame - SPF
    n_frames = (N - pad) // SPF
    plans = [parallel.block_plan(n_frames, SPF, pad, N_CHAN, r, world)
             for r in range(world)]
    first, last, in0, in1 = plans[rank]
    # Only this rank's block of the stream is handed to its chain.
    block = parallel.StreamBlock(
        x[in0:in1].copy(), in0, N, whole.start_time, RATE,
        samples_per_frame=1000, frequency=300e6, sideband=1,
        polarization=np.array(['X', 'Y']))
    dd, it = _chain_integrate(bt, block)
    assert it.shape == _chain_integrate(bt, whole)[1].shape
    edges = it._get_offsets(np.arange(it.shape[0] + 1))
    bins = [parallel.bin_range(edges, p[0] // N_CHAN, p[1] // N_CHAN)
            for p in plans]
    b0, b1 = bins[rank]
    it.seek(b0)
    sums, counts = it.read_sums(b1 - b0, within=(first // N_CHAN,
                                                 last // N_CHAN))
    parallel.reduce_edge_bins(sums, counts, bins[rank], bins)
    # Reading outside the block fails loudly.
    if world > 1:
        other = plans[1 - rank]
        try:
            block.seek(other[2] if rank else other[3] - 1)
            block.read(1)
            raised = False
        except EOFError:
            raised = True
        assert raised
    np.savez(os.path.join(out_dir, f'int{rank}.npz'), sums=sums.numpy(),
             counts=counts.numpy(), bins=np.array([b0, b1]),
             plan=np.array(plans[rank]))
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


def test_integrate_chain_sharded_over_two_ranks(tmp_path):
    import torch.multiprocessing as mp
    import backend
    backend.build_emu()
    world = 2
    port = 29500 + (os.getpid() + 7) % 1000
    mp.spawn(_worker_integrate, args=(world, port, str(tmp_path)),
             nprocs=world, join=True)
    r = [np.load(tmp_path / f'int{i}.npz') for i in range(world)]
    bt = _setup()
    whole = _source(bt, _data())
    _, it = _chain_integrate(bt, whole)
    n_units = r[1]['plan'][1] // N_CHAN         # spectra in complete frames
    ref = it.read()
    # The ranks' ranges tile the output, cut inside a bin.
    assert r[0]['plan'][1] == r[1]['plan'][0] and r[0]['plan'][0] == 0
    assert r[0]['bins'][1] - 1 == r[1]['bins'][0]
    for i in range(world):
        b0, b1 = r[i]['bins']
        want = ref[b0:b1]
        # Bins that lie wholly in complete frames are identical to the
        # single-process result; the last may be cut by the end of the
        # complete frames.
        edges = it._get_offsets(np.arange(it.shape[0] + 1))
        full = edges[b0 + 1:b1 + 1] <= n_units
        np.testing.assert_array_equal(
            r[i]['counts'][full], want['count'][full][:, 0, 0])
        np.testing.assert_allclose(
            r[i]['sums'][full].reshape(want['data'][full].shape),
            want['data'][full], rtol=1e-5,
            atol=1e-5 * np.abs(want['data']).max())
    # Both ranks hold the completed shared bin.
    np.testing.assert_array_equal(r[0]['counts'][-1], r[1]['counts'][0])
    np.testing.assert_allclose(r[0]['sums'][-1], r[1]['sums'][0], rtol=1e-6)
