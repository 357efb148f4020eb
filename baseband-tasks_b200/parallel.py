"""Multi-GPU execution: one process per GPU, time-block sharding.

The hot path shards without any exchange of samples:
rank ``r`` of ``W`` takes a contiguous run of overlap-save frames of a padded
task (`Dedisperse`) and reads its own input range, which overlaps the next
rank's by the ``pad`` samples of the halo (reference framing: base.py:775-790).
The only collective is a sum over ranks of the ``(sum, count)`` accumulators of
`Integrate` / `Fold` (reference: integration.py:301-303, 394-395), done with
``torch.distributed`` (NCCL over NVLink on GPUs; gloo on CPU in the tests).
"""
import os

import numpy as np

from . import _buffers as B
from . import _cabi
from .base import Base
from .shaping import GetSlice

__all__ = ['init', 'frame_range', 'shard_frames', 'reduce_sums', 'average',
           'StreamBlock', 'block_plan', 'bin_range', 'reduce_edge_bins']


def _dist():
    import torch.distributed as dist
    return dist


def init(backend=None):
    """Join the process group described by the environment (torchrun).

    Returns ``(rank, world_size)``; ``(0, 1)`` when not launched distributed.
    """
    dist = _dist()
    world = int(os.environ.get('WORLD_SIZE', 1))
    if world == 1:
        return 0, 1
    if not dist.is_initialized():
        import torch
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kwargs = {}
        if backend == 'nccl':
            local = int(os.environ.get('LOCAL_RANK', 0))
            torch.cuda.set_device(local)
            kwargs['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend, **kwargs)
    return dist.get_rank(), dist.get_world_size()


def frame_range(n_frames, rank, world):
    """Contiguous, balanced share ``[f0, f1)`` of ``n_frames`` frames."""
    base, extra = divmod(n_frames, world)
    f0 = rank * base + min(rank, extra)
    return f0, f0 + base + (1 if rank < extra else 0)


def shard_frames(ih, samples_per_frame, pad, rank, world):
    """This rank's time block of ``ih`` for a padded task downstream.

    The padded task (``samples_per_frame`` outputs per frame, ``pad`` extra
    input samples) run on the returned stream produces exactly the frames
    ``[f0, f1)`` it would produce on the whole of ``ih``: the block covers
    input samples ``[f0*spf, f1*spf + pad)``.  Only complete frames are
    shared out; a trailing partial frame of the whole stream is dropped.

    Returns ``(stream, f0, f1)``.
    """
    n_frames = (ih.shape[0] - pad) // samples_per_frame
    if n_frames < world:
        raise ValueError(f"only {n_frames} frames for {world} ranks.")
    f0, f1 = frame_range(n_frames, rank, world)
    start = f0 * samples_per_frame
    stop = f1 * samples_per_frame + pad
    return GetSlice(ih, slice(start, stop)), f0, f1


class StreamBlock(Base):
    """A block of a longer stream, presented as the whole stream.

    ``data`` holds the samples ``[start, start + len(data))`` of a stream of
    ``n_total`` samples.  The object has the shape, times and sample numbering
    of the whole stream, so a chain of tasks built on it has exactly the
    framing, channelizer blocks and integration bins it would have on the
    whole stream; only reads inside the block succeed.  One rank of a job
    sharded in time builds its chain on its own block (with the overlap-save
    halo, see `block_plan`) and reads the part of the output the block
    determines.
    """

    def __init__(self, data, start, n_total, start_time, sample_rate,
                 samples_per_frame=None, **kwargs):
        self._data = data
        self._block = (int(start), int(start) + int(data.shape[0]))
        if not 0 <= self._block[0] <= self._block[1] <= n_total:
            raise ValueError("block is not inside the stream.")
        if B.is_tensor(data):
            dtype = np.dtype(str(data.dtype).replace('torch.', ''))
        else:
            dtype = data.dtype
        if samples_per_frame is None:
            samples_per_frame = max(1, min(int(data.shape[0]), 1 << 20))
        super().__init__(shape=(int(n_total),) + tuple(data.shape[1:]),
                         start_time=start_time, sample_rate=sample_rate,
                         samples_per_frame=samples_per_frame, dtype=dtype,
                         **kwargs)

    def _take(self, first, count):
        a, b = self._block
        if first < a or first + count > b:
            raise EOFError(f"samples [{first}, {first + count}) are outside "
                           f"this rank's block [{a}, {b}).")
        return self._data[first - a:first - a + count]

    def _read_data(self, count, out=None):
        data = self._take(self.offset, count)
        self.offset += count
        if out is not None:
            out[...] = data
            return out
        return data

    def _read_frame(self, frame_index):
        first = frame_index * self.samples_per_frame
        return self._take(first, min(self.samples_per_frame,
                                     self.shape[0] - first))

    def close(self):
        super().close()
        self._data = None


def block_plan(n_frames, samples_per_frame, pad, unit, rank, world):
    """Share of a padded task's output for one rank of ``world``.

    The ``n_frames`` overlap-save frames (``samples_per_frame`` outputs, ``pad``
    more inputs each) are shared out evenly (`frame_range`); the rank's output
    range is cut at multiples of ``unit`` output samples (the block length of
    a channelizer downstream), so that every unit belongs to exactly one rank.

    Returns ``(out_first, out_last, in_first, in_last)``: the rank produces
    output samples ``[out_first, out_last)`` and for that needs the input
    samples ``[in_first, in_last)`` -- its frames plus the halo, and one more
    frame where a unit straddles the end of its frames.
    """
    spf = samples_per_frame
    f0, f1 = frame_range(n_frames, rank, world)
    total = (n_frames * spf // unit) * unit
    first = min(-(-f0 * spf // unit) * unit, total)
    last = total if rank == world - 1 else min(
        -(-f1 * spf // unit) * unit, total)
    if last <= first:
        return first, first, 0, 0
    in_first = (first // spf) * spf
    in_last = -(-last // spf) * spf + pad
    return first, last, in_first, in_last


def bin_range(edges, first, last):
    """Bins ``[b0, b1)`` of an `Integrate` with upstream bin edges ``edges``
    that have samples in ``[first, last)``."""
    edges = np.asarray(edges)
    b0 = int(np.searchsorted(edges, first, side='right')) - 1
    b1 = int(np.searchsorted(edges, last, side='left'))
    return max(b0, 0), min(b1, len(edges) - 1)


def reduce_edge_bins(sums, counts, bins, all_bins):
    """Complete the bins that the cut between two ranks runs through.

    ``sums`` / ``counts`` are this rank's accumulators (`Integrate.read_sums`
    with ``within``) for the bins ``bins = (b0, b1)``; ``all_bins`` lists the
    ``(b0, b1)`` of every rank.  A bin shared with a neighbour holds partial
    sums on both: the first and last bin of every rank are summed over ranks
    (one small all-reduce: NCCL over NVLink on GPUs) and added where bins
    coincide, after which both ranks hold the complete bin.
    """
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) \
            or dist.get_world_size() == 1 or sums.shape[0] == 0:
        return sums, counts
    rank, world = dist.get_rank(), dist.get_world_size()
    t = B.torch()
    inner = sums[0].numel()
    edge_s = t.zeros((world, 2, inner), dtype=sums.dtype, device=sums.device)
    edge_c = t.zeros((world, 2, counts[0].numel()), dtype=counts.dtype,
                     device=counts.device)
    edge_s[rank, 0] = sums[0].reshape(-1)
    edge_s[rank, 1] = sums[-1].reshape(-1)
    edge_c[rank, 0] = counts[0].reshape(-1)
    edge_c[rank, 1] = counts[-1].reshape(-1)
    dist.all_reduce(edge_s, op=dist.ReduceOp.SUM)
    dist.all_reduce(edge_c, op=dist.ReduceOp.SUM)
    b0, b1 = bins
    if rank > 0 and all_bins[rank - 1][1] - 1 == b0 \
            and all_bins[rank - 1][1] > all_bins[rank - 1][0]:
        sums[0] += edge_s[rank - 1, 1].reshape(sums[0].shape)
        counts[0] += edge_c[rank - 1, 1].reshape(counts[0].shape)
    if rank < world - 1 and all_bins[rank + 1][0] == b1 - 1 \
            and all_bins[rank + 1][1] > all_bins[rank + 1][0]:
        sums[-1] += edge_s[rank + 1, 0].reshape(sums[-1].shape)
        counts[-1] += edge_c[rank + 1, 0].reshape(counts[-1].shape)
    return sums, counts


def reduce_sums(sums, counts, dst=None):
    """Sum the ``(sum, count)`` accumulators over all ranks, in place.

    ``sums`` (float32) and ``counts`` (int64) are device tensors as returned
    by ``Integrate.read_sums`` / ``Fold.read_sums``.  With ``dst`` the result
    lands on that rank only (reduce), otherwise on all ranks (all-reduce).
    """
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) \
            or dist.get_world_size() == 1:
        return sums, counts
    for t in (sums, counts):
        if dst is None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        else:
            dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    return sums, counts


def average(sums, counts):
    """``sums / counts`` on the device (NaN for empty bins), like the
    division at the end of ``Integrate._read_frame`` (integration.py:268-269)."""
    lib = _cabi.lib()
    sums = sums.contiguous()
    counts = counts.contiguous()
    out = B.empty(sums.shape, np.float32)
    lib.check(lib.bbt_average_exec(
        B.ptr(sums), B.ptr(counts), B.ptr(out), counts.numel(),
        sums.numel() // max(counts.numel(), 1), _cabi.stream_ptr()))
    return out
