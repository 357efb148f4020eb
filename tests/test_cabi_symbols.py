"""The CUDA library loads (no GPU needed) and exports every entry point that
include/bbt_b200.h declares; the ctypes table binds exactly those."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'bbt_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(bbt_\w+)\s*\(', text)))


def test_library_exports_header_symbols():
    import __graft_entry__ as entry
    from baseband_tasks_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        entry.build()
    dll = ctypes.CDLL(_cabi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(dll, name), f"{name} not exported"
    assert sorted(_cabi._SIGNATURES) == names
    assert dll.bbt_version() >= 100


def test_no_cpu_fallback(monkeypatch):
    """Without a CUDA device the product path raises instead of computing."""
    import numpy as np
    import pytest
    import torch
    import baseband_tasks_b200 as bt
    from baseband_tasks_b200 import _cabi
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    monkeypatch.setattr(_cabi, '_DEVICE', None)
    src = bt.ArrayStream(np.zeros((4096, 2), 'c8'), bt.Time(0), 1e6,
                         polarization=np.array(['X', 'Y']))
    with pytest.raises(_cabi.BBTError):
        bt.Power(src).read()
