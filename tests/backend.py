"""Test backends for the C-ABI parity tests.

``cuda``: the product library (baseband-tasks_b200/csrc/libbbt_b200.so) with
torch CUDA tensors as device buffers -- the parity tests proper (marked gpu).

``emu``: the same kernel sources compiled by g++ with -DBBT_EMULATE (one OS
thread per CUDA thread, tests/emu) with numpy arrays as "device" buffers.
TEST INFRASTRUCTURE ONLY: it lets the index arithmetic of the kernels be
checked on machines without a GPU; the product package never loads it.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, 'tests', 'emu')
EMU_LIB = os.path.join(EMU_DIR, 'libbbt_emu.so')
CSRC = os.path.join(ROOT, 'baseband-tasks_b200', 'csrc')


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_emu():
    sources = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
               if f.endswith(('.cu', '.cuh'))]
    sources += [os.path.join(EMU_DIR, 'bbt_emu.cpp'),
                os.path.join(ROOT, 'include', 'bbt_b200.h')]
    import fcntl
    with open(os.path.join(EMU_DIR, '.build.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)     # one builder at a time
        if _stale(EMU_LIB, sources):
            subprocess.run(['sh', os.path.join(EMU_DIR, 'build.sh')],
                           check=True)
    return EMU_LIB


class EmuBackend:
    name = 'emu'
    # Largest problem sizes worth running on host threads.
    big = False

    def __init__(self):
        from baseband_tasks_b200 import _cabi
        self.lib = _cabi.CABI(build_emu())
        self.stream = None

    def to_dev(self, a):
        return np.ascontiguousarray(a).copy()

    def empty(self, shape, dtype):
        return np.empty(shape, dtype)

    def zeros(self, shape, dtype):
        return np.zeros(shape, dtype)

    def ptr(self, h):
        if h is None:
            return None
        return ctypes.c_void_p(h.ctypes.data)

    def to_host(self, h):
        return np.array(h)

    def sync(self):
        pass


class CudaBackend:
    name = 'cuda'
    big = True

    def __init__(self):
        import torch
        from baseband_tasks_b200 import _cabi
        self.torch = torch
        self.lib = _cabi.lib()
        self.device = _cabi.device()

    @property
    def stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def to_dev(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def _tdtype(self, dtype):
        return getattr(self.torch, np.dtype(dtype).name)

    def empty(self, shape, dtype):
        return self.torch.empty(shape, dtype=self._tdtype(dtype),
                                device=self.device)

    def zeros(self, shape, dtype):
        return self.torch.zeros(shape, dtype=self._tdtype(dtype),
                                device=self.device)

    def ptr(self, h):
        if h is None:
            return None
        return ctypes.c_void_p(h.data_ptr())

    def to_host(self, h):
        return h.cpu().numpy()

    def sync(self):
        self.torch.cuda.synchronize()


def make_backend(name):
    return EmuBackend() if name == 'emu' else CudaBackend()
