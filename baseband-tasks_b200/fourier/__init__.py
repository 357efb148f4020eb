"""Fourier transform module: the FFT maker plugin layer and the 'cuda' maker.

Mirrors baseband_tasks/fourier of the reference.  The system default maker
is the CUDA one; there is no CPU maker in this package (no CPU fallback).
"""
from .base import (fft_maker, FFTMakerBase, FFTBase, FFTMakerMeta,
                   FFT_MAKER_CLASSES)
from .cuda import CudaFFTMaker

fft_maker._system_default = CudaFFTMaker()

__all__ = ['fft_maker', 'FFTMakerBase', 'FFTBase', 'FFTMakerMeta',
           'FFT_MAKER_CLASSES', 'CudaFFTMaker']
