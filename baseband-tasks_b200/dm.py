"""Dispersion measure with the delay and phase formulas of the reference.

Mirrors baseband_tasks/dm.py: the constant is hard-coded to Tempo's
1/2.41e-4 s MHz^2 cm^3/pc (dm.py:37); ``time_delay`` dm.py:42-76,
``phase_delay`` :78-105, ``phase_factor`` :107-120.  Values are pc/cm^3;
frequencies are `astropy` quantities or numbers in Hz; delays are returned in
seconds and phases in cycles (plain floats -- astropy is optional).
"""
import numpy as np

from ._units import to_mhz

__all__ = ['DispersionMeasure']


class DispersionMeasure(float):
    """Dispersion measure in pc/cm^3."""
    dispersion_delay_constant = 1. / 2.41e-4
    """Dispersion delay constant in s MHz^2 cm^3 / pc (as for Tempo)."""

    def __new__(cls, dm):
        if hasattr(dm, 'to_value'):
            dm = dm.to_value('pc / cm3')
        return super().__new__(cls, dm)

    def __neg__(self):
        return DispersionMeasure(-float(self))

    def time_delay(self, freq, ref_freq=None):
        """Time delay in seconds at ``freq`` relative to ``ref_freq``."""
        d = self.dispersion_delay_constant * float(self)
        freq = to_mhz(freq)
        ref_freq_inv2 = 0. if ref_freq is None else 1. / to_mhz(ref_freq)**2
        return d * (1. / freq**2 - ref_freq_inv2)

    def phase_delay(self, freq, ref_freq=None):
        """Phase delay in cycles at ``freq`` relative to ``ref_freq``."""
        d = self.dispersion_delay_constant * float(self)
        freq = to_mhz(freq)
        ref_freq_inv = 0. if ref_freq is None else 1. / to_mhz(ref_freq)
        # s MHz^2 * MHz / MHz^2 = s MHz = 1e6 cycles.
        return d * freq * (ref_freq_inv - 1. / freq)**2 * 1e6

    def phase_factor(self, freq, ref_freq=None):
        """Complex exponential of the phase delay."""
        return np.exp(self.phase_delay(freq, ref_freq) * (2j * np.pi))

    def __repr__(self):
        return f"<DispersionMeasure {float(self)} pc / cm3>"
