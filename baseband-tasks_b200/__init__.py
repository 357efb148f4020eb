"""B200-native implementation of the baseband-tasks dedispersion and
channelization hot path, behind the baseband-tasks Task / FFTMaker API.

Python host code calls hand-written sm_100a CUDA kernels through the C ABI in
``include/bbt_b200.h``; PyTorch is used only for device buffers and streams.
There is no CPU fallback: using a task without the built CUDA library or
without a GPU raises.
"""
__version__ = '0.1'

from . import _cabi  # noqa: F401
from .base import (Base, BaseTaskBase, TaskBase, PaddedTaskBase, Task,  # noqa
                   SetAttribute)
from ._units import Time  # noqa: F401
from .fourier import fft_maker, CudaFFTMaker  # noqa: F401
from .dm import DispersionMeasure  # noqa: F401
from .dispersion import (Disperse, Dedisperse, DisperseSamples,  # noqa: F401
                         DedisperseSamples)
from .sampling import ShiftSamples  # noqa: F401
from .convolution import Convolve  # noqa: F401
from .channelize import Channelize, Dechannelize  # noqa: F401
from .pfb import (sinc_hamming, PolyphaseFilterBankSamples,  # noqa: F401
                  PolyphaseFilterBank)
from .functions import Square, Power  # noqa: F401
from .integration import (Integrate, Fold, PulseStack,  # noqa: F401
                          PolynomialPhase)
from .generators import (StreamGenerator, EmptyStreamGenerator, Noise,  # noqa
                         NoiseGenerator, ArrayStream, PayloadStream,
                         payload_levels, encode_payload)
