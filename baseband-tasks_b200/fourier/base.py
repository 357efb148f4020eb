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


def _registry_key(class_name):
    """``'CudaFFTMaker'`` -> ``'cuda'``: the name a maker is selected by."""
    key, suffix = class_name.lower(), 'fftmaker'
    if key.endswith(suffix) and key != suffix:
        key = key[:-len(suffix)]
    return key


class FFTMakerMeta(type):
    """Metaclass that enters every FFT maker class in `FFT_MAKER_CLASSES`.

    Defining ``class FooFFTMaker(FFTMakerBase)`` is all it takes to make
    ``fft_maker.set('foo')`` work (the plugin seam of the reference,
    fourier/base.py:221-253); a second class for the same key is an error.
    """

    def __new__(mcls, name, bases, namespace, **kwargs):
        cls = super().__new__(mcls, name, bases, namespace, **kwargs)
        if bases:                   # the root of the hierarchy stays out
            key = _registry_key(name)
            if FFT_MAKER_CLASSES.setdefault(key, cls) is not cls:
                raise ValueError(f"key {key} already registered in "
                                 "FFT_MAKER_CLASSES.")
        return cls


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


class _Restore:
    """What `fft_maker.set` returns: leaving the ``with`` block (if it is
    used as one) puts the previous selection back."""

    def __init__(self, selection, previous):
        self._selection, self._previous = selection, previous

    def __enter__(self):
        return self._selection.get()

    def __exit__(self, *exc):
        self._selection._value = self._previous
        return False

    def __repr__(self):
        return f"<ScienceState fft_maker: {self._selection._value!r}>"


class _DefaultFFTMaker:
    """Create an FFT with the default maker, or select that default.

    ``fft_maker(shape, dtype, direction=..., axis=..., ortho=...,
    sample_rate=...)`` builds a transform with the current default maker;
    ``fft_maker.set('cuda')`` (optionally ``with``) changes the default that
    new tasks capture at construction; ``fft_maker.set(None)`` restores the
    system default; ``fft_maker.get()`` returns it.  (The reference keeps this
    state in an astropy ``ScienceState``, fourier/base.py:348-466; astropy
    is optional here, so the few operations used are provided directly.)
    """
    __name__ = 'fft_maker'

    def __init__(self):
        self._system_default = None
        self._value = None

    @property
    def system_default(self):
        """System default FFT factory."""
        return self._system_default

    def __call__(self, shape, dtype, *, direction='forward', axis=0,
                 ortho=False, sample_rate=None):
        return self.get()(shape, dtype, direction=direction, axis=axis,
                          ortho=ortho, sample_rate=sample_rate)

    def validate(self, value):
        """The maker ``value`` stands for: itself, or the system default."""
        maker = self._system_default if value is None else value
        if not isinstance(maker, FFTMakerBase):
            raise TypeError("Can only set the default to an instance of "
                            "a FFT maker such as CudaFFTMaker().")
        return maker

    def get(self):
        return self.validate(self._value)

    def set(self, fft_engine, **kwargs):
        """Set the FFT factory to be used in new tasks: a maker instance, the
        registered name of one (with keyword arguments for it) or None."""
        if isinstance(fft_engine, str):
            fft_engine = FFT_MAKER_CLASSES[fft_engine](**kwargs)
        elif kwargs:
            raise TypeError("cannot pass keyword arguments except if "
                            "fft_engine is the name of an FFT maker.")
        restore = _Restore(self, self._value)
        self._value = self.validate(fft_engine)
        return restore


fft_maker = _DefaultFFTMaker()
