"""Shared helpers for the parity tests: build the oracle cube and the GPU cube from one spec,
and compare float maps with a tolerance that states its scale."""
import numpy as np

from oracle.cube import OracleCube
from oracle.wcs import OWCS

RTOL = 1e-5          # north star: "within 1e-5 relative (float32)"


def oracle_cube(data, wcs_kw, unit='K', use_dask=False, **kw):
    return OracleCube(np.asarray(data), OWCS(**wcs_kw), unit=unit, use_dask=use_dask, **kw)


def gpu_cube(data, wcs_kw, unit='K', use_dask=False, **kw):
    import spectral_cube_b200 as scb
    from spectral_cube_b200.masks import LazyMask
    cls = scb.DaskSpectralCube if use_dask else scb.SpectralCube
    if not hasattr(data, 'is_cuda'):
        data = np.asarray(data, dtype=np.float32)
    cube = cls(data, scb.CubeWCS(**wcs_kw), unit=unit, **kw)
    if 'mask' not in kw:
        # io/fits.py:214: cubes read from FITS carry LazyMask(np.isfinite)
        cube._mask = LazyMask(np.isfinite, cube=cube)
    return cube


def assert_maps_close(got, want, rtol=RTOL, atol=0.0, what=''):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    gn, wn = np.isnan(got), np.isnan(want)
    assert np.array_equal(gn, wn), "%s: NaN pattern differs at %d positions" % (what, int((gn != wn).sum()))
    ginf, winf = np.isinf(got), np.isinf(want)
    assert np.array_equal(ginf, winf) and np.array_equal(got[ginf], want[winf]), "%s: inf pattern differs" % what
    ok = ~(gn | ginf)
    err = np.abs(got[ok] - want[ok])
    tol = atol + rtol * np.abs(want[ok])
    bad = err > tol
    assert not bad.any(), "%s: %d of %d values differ; worst |err|=%g at |want|=%g (rtol=%g atol=%g)" % (
        what, int(bad.sum()), int(ok.sum()), float(err[bad].max()), float(np.abs(want[ok][bad][np.argmax(err[bad])])),
        rtol, atol)
