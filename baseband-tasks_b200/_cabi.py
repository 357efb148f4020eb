"""ctypes binding of the C ABI in include/bbt_b200.h.

The shared library ``csrc/libbbt_b200.so`` holds the hand-written sm_100a
kernels.  There is no CPU fallback: if the library is missing or no CUDA
device is present, using any task raises.
"""
import ctypes
import os
from ctypes import (POINTER, c_char_p, c_double, c_int, c_int8, c_int32,
                    c_int64, c_void_p)

__all__ = ['CABI', 'BBTError', 'lib', 'device', 'check']

_HERE = os.path.dirname(os.path.abspath(__file__))
# BBT_B200_LIB selects another build of the same CUDA library (kernel A/B runs).
LIB_PATH = os.environ.get('BBT_B200_LIB') or os.path.join(
    _HERE, 'csrc', 'libbbt_b200.so')

BBT_C2C, BBT_R2C, BBT_C2R = 0, 1, 2
BBT_FORWARD, BBT_BACKWARD = 0, 1


class BBTError(RuntimeError):
    """Error reported by the CUDA library."""


_SIGNATURES = {
    'bbt_version': (c_int, []),
    'bbt_last_error': (c_char_p, []),
    'bbt_fft_plan_create': (c_int, [POINTER(c_void_p), c_int64, c_int64,
                                    c_int64, c_int, c_int, c_double]),
    'bbt_fft_plan_work_bytes': (c_int64, [c_void_p]),
    'bbt_fft_exec': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p]),
    'bbt_fft_plan_destroy': (c_int, [c_void_p]),
    'bbt_dedisperse_plan_create': (c_int, [
        POINTER(c_void_p), c_int64, c_int64, c_int64, c_int64, c_int64,
        POINTER(c_int32), POINTER(c_double), POINTER(c_double),
        POINTER(c_int8), c_double, c_double, c_double, c_int]),
    'bbt_dedisperse_plan_set_response': (c_int, [c_void_p, c_void_p]),
    'bbt_dedisperse_plan_get_response': (c_int, [c_void_p, c_void_p]),
    'bbt_dedisperse_work_bytes': (c_int64, [c_void_p, c_int64]),
    'bbt_dedisperse_exec': (c_int, [c_void_p, c_void_p, c_int64, c_int64,
                                    c_int64, c_void_p, c_int64, c_void_p,
                                    c_void_p]),
    'bbt_dedisperse_power_supported': (c_int, [c_void_p]),
    'bbt_dedisperse_power_exec': (c_int, [c_void_p, c_void_p, c_int64,
                                          c_int64, c_int64, c_void_p, c_int64,
                                          c_void_p, c_void_p]),
    'bbt_dedisperse_plan_destroy': (c_int, [c_void_p]),
    'bbt_power_exec': (c_int, [c_void_p, c_void_p, c_int64, c_int64,
                               c_void_p]),
    'bbt_multiply_exec': (c_int, [c_void_p, c_void_p, c_void_p, c_int64,
                                  c_void_p]),
    'bbt_square_exec': (c_int, [c_void_p, c_void_p, c_int64, c_int,
                                c_void_p]),
    'bbt_channelize_power_exec': (c_int, [c_void_p, c_void_p, c_int64,
                                          c_int64, c_int64, c_void_p]),
    'bbt_channelize_power_integrate_exec': (c_int, [
        c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64,
        c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    'bbt_integrate_exec': (c_int, [c_void_p, c_int64, c_int64, c_int64,
                                   c_void_p, c_int64, c_int64, c_void_p,
                                   c_void_p, c_int, c_void_p]),
    'bbt_fold_exec': (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64,
                              c_int64, c_void_p, c_void_p, c_int64, c_int64,
                              c_void_p,
                              POINTER(c_double), c_int, c_double, c_double,
                              c_int, c_void_p, c_void_p, c_void_p]),
    'bbt_pfb_exec': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                             c_int64, c_int64, c_int, c_void_p]),
    'bbt_shift_exec': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                               c_int, c_void_p]),
    'bbt_convert_exec': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    'bbt_pair_frames_exec': (c_int, [c_void_p, c_void_p, c_int64, c_int64,
                                     c_int64, c_int64, c_int64, c_void_p]),
    'bbt_unpair_frames_exec': (c_int, [c_void_p, c_void_p, c_int64, c_int64,
                                       c_int64, c_void_p]),
    'bbt_decode_exec': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                c_void_p]),
    'bbt_average_exec': (c_int, [c_void_p, c_void_p, c_void_p, c_int64,
                                 c_int64, c_void_p]),
    'bbt_launch_count': (c_int64, []),
    'bbt_tune_set': (c_int, [ctypes.c_char_p, c_int]),
    'bbt_profile_enable': (c_int, [c_int]),
    'bbt_profile_report': (c_int, [ctypes.c_char_p, c_int64]),
    'bbt_strided_copy_bench': (c_int, [c_void_p, c_void_p, c_int64, c_int64,
                                       c_int64, c_int64, c_void_p]),
}


class CABI:
    """The loaded library with typed entry points."""

    def __init__(self, path):
        self.path = path
        self._dll = ctypes.CDLL(path)
        for name, (restype, argtypes) in _SIGNATURES.items():
            func = getattr(self._dll, name)  # AttributeError if not exported
            func.restype = restype
            func.argtypes = argtypes
            setattr(self, name, func)

    def check(self, status):
        """Turn a status code into the reference's exception types."""
        if status == 0:
            return
        msg = self.bbt_last_error().decode()
        if status == -1:
            raise ValueError(msg)
        if status == -2:
            raise NotImplementedError(msg)
        if status == -4:
            raise MemoryError(msg)
        raise BBTError(msg)


# Module state.  ``_LIB`` and ``_DEVICE`` are set on first use; the test-suite's
# kernel-emulation harness (tests/emu) overrides them explicitly, the package
# itself never does.
_LIB = None
_DEVICE = None


def lib():
    """The CUDA library; loading fails loudly if it was not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise BBTError(
                f"{LIB_PATH} not found: build the CUDA kernels first "
                "(python -c 'import __graft_entry__ as g; g.build()'). "
                "There is no CPU fallback.")
        _LIB = CABI(LIB_PATH)
    return _LIB


def device():
    """The torch device buffers live on (a CUDA device, or fail)."""
    global _DEVICE
    if _DEVICE is None:
        import torch
        if not torch.cuda.is_available():
            raise BBTError("no CUDA device available; baseband_tasks_b200 "
                           "has no CPU fallback.")
        _DEVICE = torch.device('cuda', torch.cuda.current_device())
    return _DEVICE


def stream_ptr():
    """Current torch CUDA stream as a raw pointer (0 on the emulation rig)."""
    import torch
    if device().type != 'cuda':
        return None
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def check(status):
    lib().check(status)
