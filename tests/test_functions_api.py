"""Square/Power metadata semantics after the reference's
tests/test_functions.py:97-215, on synthetic complex streams."""
import numpy as np
import pytest
from numpy.testing import assert_array_equal

from test_tasks import bt, start_time  # noqa: F401  (fixture)


def empty(bt, shape, dtype='c8', **kwargs):
    return bt.EmptyStreamGenerator(shape, start_time(bt), 1., dtype=dtype,
                                   **kwargs)


def test_polarization_propagation(bt):
    eh = empty(bt, (1000, 2))
    pt = bt.Power(bt.SetAttribute(eh, polarization=np.array(['L', 'R'])))
    assert_array_equal(pt.polarization, np.array(['LL', 'RR', 'LR', 'RL']))
    assert repr(pt).startswith('Power(ih)')
    pt = bt.Power(bt.SetAttribute(eh, polarization=np.array(['R', 'L'])))
    assert_array_equal(pt.polarization, np.array(['RR', 'LL', 'RL', 'LR']))
    # other axes, or an overly detailed array
    eh = empty(bt, (10000, 2, 4), polarization=[['L'], ['R']])
    pt = bt.Power(eh)
    expected = np.array([['LL'], ['RR'], ['LR'], ['RL']])
    assert_array_equal(pt.polarization, expected)
    assert pt.shape == (10000, 4, 4)
    pt = bt.Power(eh, polarization=np.array([['LL'] * 4, ['RR'] * 4,
                                             ['LR'] * 4, ['RL'] * 4]))
    assert_array_equal(pt.polarization, expected)


def test_frequency_sideband_propagation(bt):
    frequency = np.array([[320.25], [320.25], [336.25], [336.25]]) * 1e6
    sideband = np.array([[-1], [1], [-1], [1]])
    eh = empty(bt, (10000, 4, 2), frequency=frequency, sideband=sideband,
               polarization=['R', 'L'])
    pt = bt.Power(eh)
    assert_array_equal(pt.polarization, np.array(['RR', 'LL', 'RL', 'LR']))
    assert_array_equal(pt.frequency, eh.frequency)
    assert_array_equal(pt.sideband, eh.sideband)
    pt = bt.Power(eh, polarization=pt.polarization)
    assert_array_equal(pt.polarization, np.array(['RR', 'LL', 'RL', 'LR']))
    assert_array_equal(pt.frequency, eh.frequency)
    sq = bt.Square(eh)
    assert_array_equal(sq.polarization, np.array(['RR', 'LL']))
    assert_array_equal(sq.frequency, eh.frequency)
    assert_array_equal(sq.sideband, eh.sideband)


def test_power_failures(bt):
    eh = empty(bt, (1000, 2))
    with pytest.raises(AttributeError):
        bt.Power(eh)                                   # no polarization
    with pytest.raises(ValueError):
        bt.Power(eh, polarization=['L'])               # only one
    with pytest.raises(ValueError):
        bt.Power(eh, polarization=['L', 'L', 'R', 'R'])
    with pytest.raises(ValueError):                    # wrong axis
        bt.Power(eh, polarization=[['LL'], ['RR'], ['LR'], ['RL']])
    real = empty(bt, (1000, 2, 4), dtype='f4', polarization=[['L'], ['R']])
    with pytest.raises(ValueError):
        bt.Power(real)                                 # real time stream
    many = empty(bt, (1000, 8), polarization=np.array(['L', 'R'] * 4))
    with pytest.raises(ValueError):
        bt.Power(many)                                 # too many
    # frequency or sideband differing between the two polarizations
    sideband = np.array([[-1], [1], [-1], [1]])
    frequency = np.array([[320.25], [320.25], [336.25], [336.25]]) * 1e6
    pol = ['RR', 'LL', 'RL', 'LR']
    bad_freq = np.array([[320, 320], [320, 320], [336, 336], [336, 337]]) * 1e6
    with pytest.raises(ValueError):
        bt.Power(empty(bt, (1000, 4, 2), frequency=bad_freq,
                       sideband=sideband), polarization=pol)
    bad_side = np.array([[-1, -1], [1, -1], [-1, -1], [1, 1]])
    with pytest.raises(ValueError):
        bt.Power(empty(bt, (1000, 4, 2), frequency=frequency,
                       sideband=bad_side), polarization=pol)
