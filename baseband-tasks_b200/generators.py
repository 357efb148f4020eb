"""Source generators: streams of synthetic data for tests and benchmarks.

Mirrors baseband_tasks/generators.py of the reference: `StreamGenerator`
(:16-90), `EmptyStreamGenerator` (:93-151), `Noise` (:154-190, Philox keyed by
the seed with ``counter[1]`` = frame start offset, so frames are reproducible
and the stream is seekable) and `NoiseGenerator` (:193-245).  These run on the
host (they are inputs, not part of the accelerated path).  `ArrayStream` is
new: it presents an array -- in particular a tensor already resident in HBM --
as a stream, so a chain can be fed without any host round trip.
`PayloadStream` is new too: a packed 1/2/4/8-bit baseband payload that is
copied to the GPU as bytes and decoded there.
"""
import numpy as np

from . import _buffers as B
from . import _cabi
from .base import Base

__all__ = ['StreamGenerator', 'EmptyStreamGenerator',
           'Noise', 'NoiseGenerator', 'ArrayStream', 'PayloadStream',
           'payload_levels', 'encode_payload']


class _FrameSource(Base):
    """A stream whose frames are made on demand: subclasses say how in
    ``_make_frame``, which may rely on the sample pointer being at the first
    sample of the frame asked for."""

    def _make_frame(self):
        raise NotImplementedError

    def _read_frame(self, frame_index):
        return self._make_frame()


class StreamGenerator(_FrameSource):
    """Generator of data produced by a user-provided function.

    Parameters are as for the reference (generators.py:16-90): ``function``
    is called with the stream itself and returns one frame --
    ``samples_per_frame`` samples of sample shape ``shape[1:]`` -- for the
    position ``stream.tell()`` is at; then the shape of the whole stream, its
    start time, sample rate, frame length and dtype, and optionally
    ``frequency``, ``sideband`` and ``polarization``.
    """

    def __init__(self, function, shape, start_time, sample_rate,
                 samples_per_frame=1, dtype=np.complex64, **kwargs):
        self._function = function
        super().__init__(shape=shape, start_time=start_time,
                         sample_rate=sample_rate,
                         samples_per_frame=samples_per_frame, dtype=dtype,
                         **kwargs)

    def _make_frame(self):
        return self._function(self)


class EmptyStreamGenerator(_FrameSource):
    """Generator of a stream of uninitialised frames, to be filled by a
    `Task` (generators.py:93-151); arguments as for `Base`."""

    def _make_frame(self):
        return np.empty((self.samples_per_frame,) + self.sample_shape,
                        dtype=self.dtype)


class Noise:
    """Callable giving reproducible, normally distributed frames.

    Every frame is drawn from a Philox generator keyed by ``seed`` whose
    counter is set from the position of the frame in the stream, so a frame
    is the same whenever and in whatever order it is read, and the stream can
    be sought (generators.py:154-190).  The draw itself is what makes streams
    bit-identical to the reference's: ``normal`` in float64 over the frame,
    the last axis doubled and viewed as complex128 for complex streams, then
    cast to the stream's dtype.
    """

    def __init__(self, seed=None):
        self.seed = seed
        self._philox = np.random.Philox(seed)
        self._fresh = self._philox.state      # counter zero, nothing buffered
        self._rng = np.random.Generator(self._philox)

    def __call__(self, sh):
        state = dict(self._fresh)
        counter = self._fresh['state']['counter'].copy()
        counter[1] = sh.tell()
        state['state'] = dict(self._fresh['state'], counter=counter)
        self._philox.state = state
        shape = (sh.samples_per_frame,) + tuple(sh.sample_shape)
        if sh.complex_data:
            shape = shape[:-1] + (2 * shape[-1],)
        draws = self._rng.normal(size=shape)
        if sh.complex_data:
            draws = draws.view(np.complex128)
        return draws.astype(sh.dtype, copy=False)


class NoiseGenerator(StreamGenerator):
    """Generator of a stream of normally distributed noise.

    Arguments as for `StreamGenerator` without the function, plus ``seed``.
    Frames are reproducible (see `Noise`); ``samples_per_frame`` should be
    large (millions of samples) to keep the cost of re-keying negligible.
    """

    def __init__(self, shape, start_time, sample_rate, samples_per_frame,
                 dtype=np.complex64, seed=None, **kwargs):
        super().__init__(Noise(seed), shape, start_time, sample_rate,
                         samples_per_frame=samples_per_frame, dtype=dtype,
                         **kwargs)


class ArrayStream(Base):
    """A numpy array or device tensor presented as a stream.

    Reads return views of the data; with a device tensor the samples are
    already in HBM and downstream GPU tasks use them in place.

    ``grid=(time, index)`` states that the array is a block of a longer
    stream: its first sample is sample ``index`` of a stream that started at
    ``time`` (``start_time`` is then ``time + index / sample_rate``).  Tasks
    that number samples (`Fold` with a `PolynomialPhase`) count on that grid,
    so every block of a stream shared out over GPUs gets the bins the whole
    stream would.
    """

    def __init__(self, data, start_time, sample_rate, samples_per_frame=None,
                 grid=None, **kwargs):
        self._data = data
        self._grid = grid
        if grid is not None and start_time is None:
            start_time = grid[0] + grid[1] / sample_rate
        if B.is_tensor(data):
            dtype = np.dtype(str(data.dtype).replace('torch.', ''))
        else:
            dtype = data.dtype
        if samples_per_frame is None:
            samples_per_frame = data.shape[0]
        super().__init__(shape=tuple(data.shape), start_time=start_time,
                         sample_rate=sample_rate,
                         samples_per_frame=samples_per_frame, dtype=dtype,
                         **kwargs)

    def _sample_grid(self):
        if self._grid is None:
            return super()._sample_grid()
        return self._grid[0], int(self._grid[1])

    def _read_data(self, count, out=None):
        data = self._data[self.offset:self.offset + count]
        self.offset += count
        if out is not None:
            out[...] = data
            return out
        return data

    def _read_frame(self, frame_index):
        start = frame_index * self.samples_per_frame
        return self._data[start:start + self.samples_per_frame]

    def close(self):
        super().close()
        self._data = None


# Decoder levels of the VDIF-style encodings in the `baseband` package, which
# the reference's coded HDF5 payloads reuse (io/hdf5/payload.py:165-166).
# `baseband` is not vendored with the reference, so these defaults are a
# restatement from its documentation (parity unpinned); any table of
# 1 << bps levels can be passed instead.
_OPTIMAL_2BIT_HIGH = 3.316505
_FOUR_BIT_1_SIGMA = 2.95


def payload_levels(bps):
    """Default float32 value of each ``bps``-bit code."""
    if bps == 1:
        levels = [-1., 1.]
    elif bps == 2:
        levels = [-_OPTIMAL_2BIT_HIGH, -1., 1., _OPTIMAL_2BIT_HIGH]
    elif bps == 4:
        levels = (np.arange(16) - 8.) / _FOUR_BIT_1_SIGMA
    elif bps == 8:
        levels = np.arange(256) - 127.5    # 0..255 encode -127.5..127.5
    else:
        raise NotImplementedError("bits per sample must be 1, 2, 4 or 8")
    return np.asarray(levels, dtype=np.float32)


def encode_payload(data, bps, levels=None):
    """Pack samples into ``bps``-bit codes (nearest level), as uint8 bytes.

    Host-side helper for tests and benchmarks: the first value goes in the
    least significant bits; complex samples are (re, im) value pairs.
    """
    levels = payload_levels(bps) if levels is None else np.asarray(levels)
    data = np.ascontiguousarray(data)
    if data.dtype.kind == 'c':
        data = data.view(data.real.dtype)
    values = data.reshape(-1)
    order = np.argsort(levels)
    sorted_levels = levels[order]
    edges = ((sorted_levels[1:] + sorted_levels[:-1]) / 2).astype(values.dtype)
    codes = np.empty(len(values), np.uint8)
    chunk = 1 << 24              # bounds the int64 temporaries
    steps = np.diff(levels.astype(np.float64))
    uniform = len(levels) > 2 and np.allclose(steps, steps[0]) and steps[0] > 0
    for i in range(0, len(values), chunk):
        v = values[i:i + chunk]
        if uniform:      # evenly spaced, increasing levels: round
            c = np.rint((v - levels[0]) * values.dtype.type(1. / steps[0]))
            codes[i:i + chunk] = np.clip(c, 0, len(levels) - 1)
        else:
            codes[i:i + chunk] = order[np.searchsorted(edges, v)]
    per_byte = 8 // bps
    pad = -len(codes) % per_byte
    if pad:
        codes = np.concatenate([codes, np.zeros(pad, np.uint8)])
    codes = codes.reshape(-1, per_byte)
    shifts = (np.arange(per_byte) * bps).astype(np.uint8)
    return np.bitwise_or.reduce(codes << shifts, axis=1).astype(np.uint8)


class PayloadStream(Base):
    """A packed baseband payload presented as a stream, decoded on the GPU.

    ``words`` holds ``bps``-bit codes (any integer dtype, used as bytes; first
    value in the least significant bits), time-major over ``sample_shape``
    with complex samples stored as (re, im) pairs.  Only the packed bytes
    cross PCIe; `bbt_decode_exec` expands them to float32/complex64 in HBM.
    """

    def __init__(self, words, bps, sample_shape, start_time, sample_rate,
                 samples_per_frame=None, complex_data=False, levels=None,
                 **kwargs):
        if bps not in (1, 2, 4, 8):
            raise NotImplementedError("bits per sample must be 1, 2, 4 or 8")
        if B.is_tensor(words):
            self._words = words.contiguous().view(B.torch_dtype(np.uint8))
            n_bytes = self._words.numel()
        else:
            self._words = np.ascontiguousarray(words).reshape(-1).view(np.uint8)
            n_bytes = self._words.size
        self.bps = bps
        sample_shape = tuple(sample_shape)
        self._values_per_sample = (int(np.prod(sample_shape, dtype=np.int64))
                                   * (2 if complex_data else 1))
        n_sample = n_bytes * 8 // (bps * self._values_per_sample)
        if samples_per_frame is None:
            samples_per_frame = n_sample
        levels = payload_levels(bps) if levels is None else np.asarray(
            levels, dtype=np.float32)
        if levels.shape != (1 << bps,):
            raise ValueError("need one level for each of the 1 << bps codes")
        self._levels = levels
        self._d_levels = None
        super().__init__(shape=(n_sample,) + sample_shape,
                         start_time=start_time, sample_rate=sample_rate,
                         samples_per_frame=samples_per_frame,
                         dtype=np.complex64 if complex_data else np.float32,
                         **kwargs)

    def _decode(self, start, count):
        bit0 = start * self._values_per_sample * self.bps
        if bit0 % 8:
            raise ValueError("read does not start on a byte of the payload")
        n = count * self._values_per_sample
        words = B.as_device(
            self._words[bit0 // 8:(bit0 + n * self.bps + 7) // 8])
        if self._d_levels is None:
            self._d_levels = B.as_device(self._levels)
        out = B.empty((n,), np.float32)
        lib = _cabi.lib()
        lib.check(lib.bbt_decode_exec(B.ptr(words), B.ptr(out),
                                      B.ptr(self._d_levels), n, self.bps,
                                      _cabi.stream_ptr()))
        if self.complex_data:
            out = B.torch().view_as_complex(out.view(-1, 2))
        return out.view((count,) + self.sample_shape)

    def _read_data(self, count, out=None):
        start = self.offset
        if (start * self._values_per_sample * self.bps) % 8:
            return super()._read_data(count, out)   # frame by frame
        data = self._decode(start, count)
        self.offset = start + count
        if out is not None:
            out[...] = data if B.is_tensor(out) else B.as_host(data)
            return out
        return data

    def _read_frame(self, frame_index):
        start = frame_index * self.samples_per_frame
        return self._decode(start, min(self.samples_per_frame,
                                       self.shape[0] - start))

    def close(self):
        super().close()
        self._words = None
