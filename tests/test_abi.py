"""
CPU-side checks of the drop-in boundary: libsc_b200.so builds/loads here (nvcc cross-compiles),
exports every symbol include/sc_b200.h declares, and validates arguments before launching
anything (so these calls are safe without a GPU).
"""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'sc_b200.h')


@pytest.fixture(scope='module')
def lib():
    from spectral_cube_b200 import _lib, build
    build.build_library()
    return _lib.load()


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    names = re.findall(r'^\s*(?:const\s+char\s*\*\s*|(?:int64_t|size_t|int|float|void)\s+)(sc_\w+)\s*\(', text, flags=re.M)
    return sorted(set(names))


def test_header_declares_entry_points():
    names = declared_functions()
    for must in ('sc_moments_axis0', 'sc_spectral_smooth', 'sc_spatial_smooth_sep', 'sc_spectral_interp',
                 'sc_reproject', 'sc_moments_axis0_host', 'sc_last_error'):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    from spectral_cube_b200 import _lib
    for name in declared_functions():
        assert hasattr(lib, name), "libsc_b200.so does not export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes prototype for %s" % name
    for name in _lib.SIGNATURES:
        assert name in declared_functions(), "%s is bound but not declared in the header" % name


def test_version_and_error_slot(lib):
    assert lib.sc_version() == 1
    assert isinstance(lib.sc_last_error(), bytes)


def test_argument_validation_happens_before_any_launch(lib):
    from spectral_cube_b200 import _lib
    n0 = lib.sc_launch_count()
    # NULL cube
    rc = lib.sc_moments_axis0(None, 4, 4, 4, 16, 4, None, None, 1.0, 0.0, 1, None, None, None, None, 0, None)
    assert rc == -1 and b'NULL' in lib.sc_last_error()
    # bad shape
    buf = (C.c_float * 64)()
    out = (C.c_double * 16)()
    p, o = C.addressof(buf), C.addressof(out)
    rc = lib.sc_moments_axis0(p, 0, 4, 4, 16, 4, None, None, 1.0, 0.0, 1, o, None, None, None, 0, None)
    assert rc == -1 and b'shape' in lib.sc_last_error()
    # stride smaller than the row
    rc = lib.sc_moments_axis0(p, 4, 4, 4, 16, 2, None, None, 1.0, 0.0, 1, o, None, None, None, 0, None)
    assert rc == -1 and b'stride_y' in lib.sc_last_error()
    # missing output
    rc = lib.sc_moments_axis0(p, 4, 4, 4, 16, 4, None, None, 1.0, 0.0, 3, o, None, None, None, 0, None)
    assert rc == -1 and b'out_m1' in lib.sc_last_error()
    # moment 1 without channel coordinates
    rc = lib.sc_moments_axis0(p, 4, 4, 4, 16, 4, None, None, 1.0, 0.0, 2, None, o, None, None, 0, None)
    assert rc == -1 and b'chan_offset' in lib.sc_last_error()
    # workspace too small
    x = (C.c_double * 4)(0, 1, 2, 3)
    rc = lib.sc_moments_axis0(p, 4, 4, 4, 16, 4, None, x, 1.0, 0.0, 2, None, o, None, None, 0, None)
    assert rc == -4 and b'workspace' in lib.sc_last_error()
    # malformed mask: child index does not precede the node
    m = _lib.MaskDesc()
    m.n_nodes = 1
    m.nodes[0].kind = _lib.MASK_NOT
    m.nodes[0].a = 0
    rc = lib.sc_moments_axis0(p, 4, 4, 4, 16, 4, m, None, 1.0, 0.0, 1, o, None, None, None, 0, None)
    assert rc == -1 and b'mask' in lib.sc_last_error()
    assert lib.sc_launch_count() == n0


def test_workspace_sizes(lib):
    from spectral_cube_b200 import _lib
    assert lib.sc_workspace_bytes(_lib.OP_MOMENTS, 1024, 2048, 2048, 0) >= 1024 * 24
    assert lib.sc_workspace_bytes(_lib.OP_SPECTRAL_SMOOTH, 1024, 2048, 2048, 17) >= 17 * 8


def test_product_refuses_to_run_without_a_gpu():
    """No CPU fallback: constructing a cube without a CUDA device must raise, loudly."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import spectral_cube_b200 as scb
    from tests.golden import reference_goldens as G
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        scb.SpectralCube(np.zeros((3, 3, 3), dtype=np.float32), scb.CubeWCS(**G.MOMENT_WCS))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'spectral_cube_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
