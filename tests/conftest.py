import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def host(monkeypatch):
    """Host-plumbing harness for CPU tests: tensors stay on the host and the REAL libsc_b200.so is called, so ctypes
    marshals every argument and the library validates it; calls stop at their first CUDA API call (no device here),
    which this fixture -- and nothing in the product -- tolerates.  Yields (package, list of (rc, message))."""
    import torch
    import spectral_cube_b200 as S
    from spectral_cube_b200 import cube as C, _lib

    class _Stream(object):
        cuda_stream = 0

    calls = []

    def check(rc):
        msg = _lib.load().sc_last_error().decode() if rc else ''
        calls.append((rc, msg))
        if rc and 'CUDA error' not in msg and 'not available from the driver' not in msg:
            raise AssertionError("the library refused the arguments: %d %s" % (rc, msg))

    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: _Stream())
    monkeypatch.setattr(_lib, 'require_cuda', lambda: torch)
    monkeypatch.setattr(_lib, 'check', check)
    monkeypatch.setattr(C, '_stream', lambda: 0)
    return S, calls
