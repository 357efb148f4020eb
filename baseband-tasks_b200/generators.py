"""Source generators: streams of synthetic data for tests and benchmarks.

Mirrors baseband_tasks/generators.py of the reference: `StreamGenerator`
(:16-90), `EmptyStreamGenerator` (:93-151), `Noise` (:154-190, Philox keyed by
the seed with ``counter[1]`` = frame start offset, so frames are reproducible
and the stream is seekable) and `NoiseGenerator` (:193-245).  These run on the
host (they are inputs, not part of the accelerated path).  `ArrayStream` is
new: it presents an array -- in particular a tensor already resident in HBM --
as a stream, so a chain can be fed without any host round trip.
"""
import numpy as np

from . import _buffers as B
from .base import Base

__all__ = ['StreamGenerator', 'EmptyStreamGenerator',
           'Noise', 'NoiseGenerator', 'ArrayStream']


class StreamGenerator(Base):
    """Generator of data produced by a user-provided function.

    ``function(stream)`` returns ``samples_per_frame`` samples of sample
    shape ``shape[1:]``; it can count on ``stream.tell()`` being at the start
    of the frame.
    """

    def __init__(self, function, shape, start_time, sample_rate,
                 samples_per_frame=1, dtype=np.complex64, **kwargs):
        super().__init__(shape=shape, start_time=start_time,
                         sample_rate=sample_rate,
                         samples_per_frame=samples_per_frame, dtype=dtype,
                         **kwargs)
        self._function = function

    def _read_frame(self, frame_index):
        return self._function(self)


class EmptyStreamGenerator(Base):
    """Generator of an empty data stream, to be filled by a `Task`."""

    def _read_frame(self, frame_index):
        return np.empty((self.samples_per_frame,) + self.shape[1:],
                        self.dtype)


class Noise:
    """Source callable providing reproducible normally distributed frames."""

    def __init__(self, seed=None):
        self.seed = seed
        self.rng = np.random.Generator(np.random.Philox(self.seed))
        self.bg_state = self.rng.bit_generator.state

    def __call__(self, sh):
        self.bg_state['state']['counter'][1] = sh.tell()
        self.rng.bit_generator.state = self.bg_state
        shape = (sh.samples_per_frame,) + sh.sample_shape
        if sh.complex_data:
            shape = shape[:-1] + (shape[-1] * 2,)
        numbers = self.rng.normal(size=shape)
        if sh.complex_data:
            numbers = numbers.view(np.complex128)
        return numbers.astype(sh.dtype, copy=False)


class NoiseGenerator(StreamGenerator):
    """Generator of a stream of normally distributed noise.

    Data are identical if read multiple times, since the random number
    generator is re-keyed by the frame offset; choose ``samples_per_frame``
    large (of order millions of samples).
    """

    def __init__(self, shape, start_time, sample_rate, samples_per_frame,
                 dtype=np.complex64, seed=None, **kwargs):
        generator = Noise(seed)
        super().__init__(function=generator, shape=shape,
                         start_time=start_time, sample_rate=sample_rate,
                         samples_per_frame=samples_per_frame,
                         dtype=dtype, **kwargs)


class ArrayStream(Base):
    """A numpy array or device tensor presented as a stream.

    Reads return views of the data; with a device tensor the samples are
    already in HBM and downstream GPU tasks use them in place.
    """

    def __init__(self, data, start_time, sample_rate, samples_per_frame=None,
                 **kwargs):
        self._data = data
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
