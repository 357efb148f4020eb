"""Test configuration: markers, import paths and the kernel backends."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a CUDA device (run on the B200 box)")


_BACKENDS = {}


@pytest.fixture(params=['emu', pytest.param('cuda', marks=pytest.mark.gpu)])
def backend(request):
    """Kernel backend: host-thread emulation here, the CUDA library on GPU."""
    import backend as _b
    name = request.param
    if name not in _BACKENDS:
        _BACKENDS[name] = _b.make_backend(name)
    return _BACKENDS[name]
