"""
Every method of the hot-path surface, driven on CPU tensors into the real library (see the `host` fixture in
tests/conftest.py): no Python-level error on the way, result types / shapes / units / beams as the reference's,
and the library accepts every argument set it is handed (it fails only at its first CUDA call -- there is no device
here).  Numbers are not looked at: that is what the `-m gpu` parity tests do.
"""
import warnings

import numpy as np
import pytest

from tests.golden import reference_goldens as G


@pytest.mark.parametrize('use_dask', [False, True])
def test_the_method_surface_reaches_the_library(host, use_dask):
    S, calls = host
    from spectral_cube_b200.masks import LazyMask
    rng = np.random.default_rng(0)
    cls = S.DaskSpectralCube if use_dask else S.SpectralCube
    cube = cls(rng.normal(size=(16, 24, 32)).astype(np.float32), S.CubeWCS(**G.ADV_WCS), unit='K',
               header={'BMAJ': 1 / 3600., 'BMIN': 1 / 3600., 'BPA': 0.0})
    cube._mask = LazyMask(np.isfinite, cube=cube)
    beam = cube.beam
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = cube.with_mask(cube > 0.5)
        for order in (0, 1, 2):
            p = m.moment(order=order)
            assert isinstance(p, S.Projection) and p.shape == (24, 32)
        assert len(m.moments012()) == 3 and m.linewidth_fwhm().shape == (24, 32)
        for axis in (1, 2):
            assert m.moment(order=1, axis=axis).shape == tuple(n for i, n in enumerate(m.shape) if i != axis)
        s = m.spectral_smooth(S.Gaussian1DKernel(1.0))
        assert s.moment1().shape == (24, 32) and s.unmasked_data[:].shape == m.shape and s.beam == beam
        sp = m.spatial_smooth(S.Gaussian2DKernel(1.0))
        assert sp.filled_data[:].shape == m.shape and sp.beam == beam and sp.mask is m.mask
        sp = m.spatial_smooth(S.Tophat2DKernel(2))
        axis = cube.spectral_axis
        i = m.spectral_interpolate(np.linspace(axis[0], axis[-1], 9))
        assert i.shape == (9, 24, 32) and i.beam == beam
        v = m[2:10, 3:20, 4:28]
        assert v.shape == (8, 17, 24) and v.moment0().shape == (17, 24) and v.beam == beam
        assert m.spectral_slab(axis[2], axis[9]).shape[0] == 8
        assert m.sum(axis=0).shape == (24, 32) and m.std(axis=0).shape == (24, 32)
        assert m.argmax(axis=0).dtype == np.int64
        for scalar in (m.mean(), m.max(), m.sum(), m.std()):
            assert np.ndim(scalar) == 0
        try:
            r = m.reproject(dict(cube.header))
            assert type(r) is cls and r.shape == m.shape and r.beam == beam
        except ValueError as exc:              # the any-valid flag is whatever the never-run kernel left there
            assert "All values in reprojected cube are nan" in str(exc)
        if use_dask:
            assert set(m.statistics()) == {'npts', 'min', 'max', 'sum', 'sumsq', 'mean', 'sigma', 'rms'}
    assert len(calls) > 15 and all(rc != 0 for rc, msg in calls)      # every call got as far as the missing device


def test_mask_channels(host):
    S, calls = host
    cube = S.SpectralCube(np.zeros((4, 6, 8), dtype=np.float32), S.CubeWCS(**G.ADV_WCS), unit='K')
    masked = cube.mask_channels([True, False, True, True])
    assert isinstance(masked.mask, S.BooleanArrayMask) and masked.shape == cube.shape
    with pytest.raises(ValueError, match="one-dimensional"):
        cube.mask_channels(np.ones((4, 1), dtype=bool))
    with pytest.raises(ValueError, match="length equal"):
        cube.mask_channels([True, False])


def test_spatial_axis_reductions_reach_the_library(host, monkeypatch):
    S, calls = host
    from spectral_cube_b200.masks import LazyMask
    cube = S.SpectralCube(np.zeros((4, 6, 8), dtype=np.float32), S.CubeWCS(**G.ADV_WCS), unit='K')
    cube._mask = LazyMask(np.isfinite, cube=cube)
    del calls[:]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for axis, shape in ((1, (4, 8)), (2, (4, 6))):
            for name in ('sum', 'mean', 'std', 'max', 'min'):
                out = getattr(cube, name)(axis=axis)
                assert isinstance(out, S.Projection) and out.shape == shape and out.meta['collapse_axis'] == axis
            assert cube.argmax(axis=axis).shape == shape and cube.argmin(axis=axis).dtype == np.int64
    assert len(calls) == 14 and all(rc != 0 for rc, msg in calls)
    with pytest.raises(NotImplementedError):
        cube.sum(axis=3)
