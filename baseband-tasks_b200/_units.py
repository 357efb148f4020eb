"""Unit and time conventions.

The reference passes `astropy.units.Quantity` and `astropy.time.Time` around
(base.py:9-10).  astropy is an optional dependency here: plain numbers are
taken to be in SI base units (Hz, seconds), a `Time` stand-in with the few
operations the task framework needs is provided, and anything that quacks
like a Quantity (``to_value``) or an astropy Time is passed through unchanged
so that the reference's expressions (``start_time + offset / sample_rate``)
keep working with either kind.
"""
import math
import operator

import numpy as np

__all__ = ['Time', 'to_hz', 'to_mhz', 'to_seconds', 'to_float', 'is_index']


def is_index(n):
    """Whether ``n`` can be used as an integer index (integration.py:42-49)."""
    try:
        operator.index(n)
    except TypeError:
        return False
    return True


def _value(x, unit):
    if hasattr(x, 'to_value'):
        return x.to_value(unit)
    return x


def to_hz(x):
    """Frequency-like quantity or number (Hz) as float or float array."""
    v = _value(x, 'Hz')
    return np.asarray(v, dtype=float) if np.ndim(v) else float(v)


def to_mhz(x):
    v = to_hz(x)
    return v / 1e6


def to_seconds(x):
    """Time-interval quantity or number (s) as float."""
    v = _value(x, 's')
    return np.asarray(v, dtype=float) if np.ndim(v) else float(v)


def to_float(x):
    """Dimensionless quantity or number as float (array)."""
    if hasattr(x, 'decompose'):
        x = x.decompose().value
    elif hasattr(x, 'to_value'):
        x = x.to_value('')
    return np.asarray(x, dtype=float) if np.ndim(x) else float(x)


class Time:
    """Minimal stand-in for `astropy.time.Time`: seconds since an epoch.

    Kept as integer seconds plus a float64 fraction so that adding sample
    offsets keeps sub-nanosecond precision over long streams (the role of
    astropy's two-double representation).  Scalar or array valued.  Supports
    what the task framework uses: ``time + seconds``, ``time - seconds``,
    ``time - time`` (float seconds), comparisons and indexing.
    """
    __slots__ = ('sec', 'frac')
    __array_priority__ = 1000

    def __init__(self, sec=0, frac=0.):
        if isinstance(sec, Time):
            sec, frac = sec.sec, sec.frac + frac
        sec = np.asarray(sec)
        frac = np.asarray(frac, dtype=np.float64)
        if sec.dtype.kind not in 'iu':
            whole = np.floor(sec)
            frac = frac + (sec - whole)
            sec = whole.astype(np.int64)
        carry = np.floor(frac)
        sec = sec.astype(np.int64) + carry.astype(np.int64)
        frac = frac - carry
        if sec.ndim == 0 and frac.ndim == 0:
            self.sec, self.frac = int(sec), float(frac)
        else:
            self.sec, self.frac = np.broadcast_arrays(sec, frac)

    @property
    def shape(self):
        return np.shape(self.frac)

    def __getitem__(self, item):
        return Time(np.asarray(self.sec)[item], np.asarray(self.frac)[item])

    def __add__(self, other):
        if isinstance(other, Time):
            return NotImplemented
        other = to_seconds(other)
        return Time(self.sec, self.frac + other)

    __radd__ = __add__

    def __sub__(self, other):
        if isinstance(other, Time):
            return (self.sec - other.sec) + (self.frac - other.frac)
        return self + (-to_seconds(other))

    def _cmp(self, other):
        return np.sign(self - other)

    def __eq__(self, other):
        if not isinstance(other, Time):
            return False
        return self._cmp(other) == 0

    def __lt__(self, other):
        return self._cmp(other) < 0

    def __le__(self, other):
        return self._cmp(other) <= 0

    def __gt__(self, other):
        return self._cmp(other) > 0

    def __ge__(self, other):
        return self._cmp(other) >= 0

    __hash__ = None

    def __float__(self):
        return float(self.sec + self.frac)

    def __repr__(self):
        return f"Time({self.sec!r}, {self.frac!r})"
