"""FFT plugin layer: `FFTBase`, the maker registry and the `fft_maker` state.

Same plugin seam as the reference (fourier/base.py): maker classes register
themselves in `FFT_MAKER_CLASSES` under their lower-cased name without the
``fftmaker`` suffix (:221-253); a maker instance builds per-transform objects
(:262-311) that know their time/frequency shapes and dtypes, ``axis``,
``ortho``, ``sample_rate``, ``frequency`` and ``inverse()`` (:59-218);
``fft_maker.set(name, **kwargs)`` selects the maker new tasks will capture, also
as a context manager (:348-466; astropy's ScienceState re-stated here since
astropy is optional).
"""
import operator

import numpy as np

__all__ = ['FFTMakerBase', 'FFTBase', 'fft_maker',
           'FFTMakerMeta', 'FFT_MAKER_CLASSES']

FFT_MAKER_CLASSES = {}
"""Dict for storing FFT maker classes, indexed by their name or prefix."""


class FFTBase:
    """Single pre-defined FFT and its metadata (fourier/base.py:59-218)."""

    def __init__(self, direction):
        self._direction = direction if direction == 'backward' else 'forward'

    direction = property(lambda self: self._direction)
    time_shape = property(lambda self: self._time_shape)
    time_dtype = property(lambda self: self._time_dtype)
    frequency_shape = property(lambda self: self._frequency_shape)
    frequency_dtype = property(lambda self: self._frequency_dtype)
    axis = property(lambda self: self._axis)
    ortho = property(lambda self: self._ortho)
    sample_rate = property(lambda self: self._sample_rate)

    @property
    def frequency(self):
        """FFT sample frequencies, with trailing unit dimensions so that they
        broadcast against the transformed data (fourier/base.py:114-157)."""
        sample_rate = 1. if self.sample_rate is None else self.sample_rate
        a_length = self._time_shape[self.axis]
        if self._time_dtype.kind == 'f':
            frequency = np.fft.rfftfreq(a_length)
        else:
            frequency = np.fft.fftfreq(a_length)
        frequency = frequency.reshape(
            frequency.shape + (len(self._time_shape) - self.axis - 1) * (1,))
        return frequency * sample_rate

    def __call__(self, a):
        """Transform ``a`` along ``axis``."""
        return self._fft(a)

    _DESCRIPTION = ('direction', 'time_shape', 'time_dtype', 'frequency_shape',
                    'frequency_dtype', 'axis', 'ortho', 'sample_rate')

    def inverse(self):
        """The same transform in the opposite direction."""
        flipped = {'forward': 'backward', 'backward': 'forward'}
        return type(self)(direction=flipped[self.direction])

    def __copy__(self):
        return type(self)(direction=self.direction)

    def __eq__(self, other):
        # Two transforms are equal if everything that describes them is.
        try:
            return all(getattr(self, name) == getattr(other, name)
                       for name in self._DESCRIPTION)
        except AttributeError:
            return NotImplemented

    __hash__ = None

    def __repr__(self):
        return (f"<{type(self).__name__} direction={self.direction},\n"
                f"    axis={self.axis}, ortho={self.ortho}, "
                f"sample_rate={self.sample_rate}\n"
                f"    Time domain: shape={self.time_shape}, "
                f"dtype={self.time_dtype}\n"
                f"    Frequency domain: shape={self.frequency_shape}, "
                f"dtype={self.frequency_dtype}>")


class FFTMakerMeta(type):
    """Registry of FFT maker classes (fourier/base.py:221-253)."""
    _registry = FFT_MAKER_CLASSES

    def __init__(cls, name, bases, dct):
        if name != 'FFTMakerBase':
            key = name.lower()
            if key.endswith('fftmaker') and len(key) > 8:
                key = key[:-8]
            if key in FFTMakerMeta._registry:
                raise ValueError("key {0} already registered in "
                                 "FFT_MAKER_CLASSES.".format(key))
            FFTMakerMeta._registry[key] = cls
        super().__init__(name, bases, dct)


class FFTMakerBase(metaclass=FFTMakerMeta):
    """Base class for all FFT factories (fourier/base.py:256-346)."""
    _FFTBase = FFTBase
    _repr_kwargs = {}

    def __call__(self, shape, dtype, direction='forward', axis=0, ortho=False,
                 sample_rate=None, **kwargs):
        time_shape, time_dtype = tuple(shape), np.dtype(dtype)
        axis = operator.index(axis)
        frequency_shape, frequency_dtype = self.get_frequency_data_info(
            time_shape, time_dtype, axis=axis)
        # The transform is an instance of a class made on the spot, with the
        # description of the data as (private) class attributes, so that
        # ``inverse()`` and ``copy`` only need the direction
        # (fourier/base.py:262-311).
        described = dict(time_shape=time_shape, time_dtype=time_dtype,
                         frequency_shape=frequency_shape,
                         frequency_dtype=frequency_dtype, axis=axis,
                         ortho=bool(ortho), sample_rate=sample_rate, **kwargs)
        cls = type(self._FFTBase.__name__.replace('Base', ''),
                   (self._FFTBase,),
                   {'_' + key: value for key, value in described.items()})
        return cls(direction)

    def get_frequency_data_info(self, shape, dtype, axis=0):
        """Shape and dtype of the frequency-domain array: real data keep the
        ``n // 2 + 1`` non-negative frequencies and become complex."""
        if dtype.kind != 'f':
            return shape, dtype
        half = shape[:axis] + (shape[axis] // 2 + 1,) + shape[axis + 1:]
        return tuple(half), np.dtype(f'c{2 * dtype.itemsize}')

    def __repr__(self):
        settings = ', '.join(f'{key}={value}'
                             for key, value in self._repr_kwargs.items())
        return f'{type(self).__name__}({settings})'


class _StateContext:
    def __init__(self, parent, value):
        self._parent = parent
        self._value = value

    def __enter__(self):
        pass

    def __exit__(self, type, value, tb):
        self._parent._value = self._value

    def __repr__(self):
        return f"<ScienceState {self._parent.__name__}: " \
               f"{self._parent._value!r}>"


class _FFTMakerState(type):
    @property
    def system_default(cls):
        """System default FFT factory."""
        return cls._system_default


class fft_maker(metaclass=_FFTMakerState):
    """Create an FFT with the default maker, or select that default.

    ``fft_maker(shape, dtype, direction=..., axis=..., ortho=...,
    sample_rate=...)`` builds a transform with the current default maker;
    ``fft_maker.set('cuda')`` (optionally ``with``) changes the default that
    new tasks capture at construction; ``fft_maker.set(None)`` restores the
    system default; ``fft_maker.get()`` returns it.
    """
    _system_default = None
    _value = None

    def __new__(cls, shape, dtype, *,
                direction='forward', axis=0, ortho=False, sample_rate=None):
        fft_engine = cls.get()
        return fft_engine(shape, dtype, direction=direction, axis=axis,
                          ortho=ortho, sample_rate=sample_rate)

    @classmethod
    def get(cls):
        return cls.validate(cls._value)

    @classmethod
    def validate(cls, value):
        if value is None:
            value = cls._system_default
        if not isinstance(value, FFTMakerBase):
            raise TypeError("Can only set the default to an instance of "
                            "a FFT maker such as CudaFFTMaker().")
        return value

    @classmethod
    def set(cls, fft_engine, **kwargs):
        """Set the FFT factory to be used in new tasks."""
        if fft_engine is None:
            fft_engine = cls._system_default
        elif not isinstance(fft_engine, FFTMakerBase):
            fft_engine = FFT_MAKER_CLASSES[fft_engine](**kwargs)
        elif kwargs:
            raise TypeError("cannot pass keyword arguments except if "
                            "fft_engine is the name of an FFT maker.")
        ctx = _StateContext(cls, cls._value)
        cls._value = cls.validate(fft_engine)
        return ctx
