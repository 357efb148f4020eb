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
from .shaping import GetSlice

__all__ = ['init', 'frame_range', 'shard_frames', 'reduce_sums', 'average']


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
