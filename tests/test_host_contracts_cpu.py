"""
Error / refusal contracts of the reference that live in the host layer (CPU: the `host` fixture drives the real
library without a device).  Each case cites the reference lines it mirrors.
"""
import warnings

import numpy as np
import pytest

from tests.golden import reference_goldens as G


def _cube(S, cls=None, shape=(4, 6, 8), unit='K', **kw):
    from spectral_cube_b200.masks import LazyMask
    cls = cls or S.SpectralCube
    cube = cls(np.zeros(shape, dtype=np.float32), S.CubeWCS(**G.ADV_WCS), unit=unit, **kw)
    cube._mask = LazyMask(np.isfinite, cube=cube)
    return cube


@pytest.mark.parametrize('use_dask', [False, True])
def test_huge_operations_are_refused_like_warn_slow(host, monkeypatch, use_dask):
    """utils.py:41-75 + tests/test_spectral_cube.py:104-133: with the threshold lowered the cube `_is_huge`; the numpy
    class refuses reductions (unless how='slice'/'ray'), reproject and convolve_to with the reference's text; the dask
    class overrides the reductions and convolve_to without the decorator (dask_spectral_cube.py:641-767, 1412) but
    inherits `reproject`."""
    S, calls = host
    from spectral_cube_b200 import cube as C
    cls = S.DaskSpectralCube if use_dask else S.SpectralCube
    cube = _cube(S, cls, header={'BMAJ': 1 / 3600., 'BMIN': 1 / 3600., 'BPA': 0.0})
    assert not cube._is_huge
    monkeypatch.setattr(C, 'MEMORY_THRESHOLD', 10)
    assert cube._is_huge
    with pytest.raises(ValueError, match='entire cube into memory') as exc:
        cube.reproject(dict(cube.header))
    assert '`cube.allow_huge_operations=True`' in str(exc.value) and 'big_data.html' in str(exc.value)
    assert '(%d pixels)' % cube.size in str(exc.value).replace('\n', ' ')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        if use_dask:
            assert cube.max(axis=0, how='cube').shape == (6, 8)
            cube.convolve_to(S.Beam.from_arcsec(3.0))
        else:
            with pytest.raises(ValueError, match='entire cube into memory') as exc:
                cube.max(how='cube')
            assert "how='slice' or how='ray'" in str(exc.value)
            with pytest.raises(ValueError, match='entire cube into memory'):
                cube.sum(axis=0)
            assert cube.max(axis=0, how='slice').shape == (6, 8)          # loads_whole_cube is False
            with pytest.raises(ValueError, match='entire cube into memory'):
                cube.convolve_to(S.Beam.from_arcsec(3.0))
        cube.allow_huge_operations = True
        assert cube.max(axis=0).shape == (6, 8)
        cube.convolve_to(S.Beam.from_arcsec(3.0))
        try:
            cube.reproject(dict(cube.header))
        except ValueError as exc:              # the any-valid flag is whatever the never-run kernel left there
            assert "All values in reprojected cube are nan" in str(exc)
        # the flag travels with derived cubes (spectral_cube.py:283)
        assert cube.with_mask(cube > 1).allow_huge_operations is True


def _named(name):
    def f(*a, **k):
        raise AssertionError("the convolution runs on the device; this must never be called")
    f.__name__ = name
    return f


@pytest.mark.parametrize('use_dask', [False, True])
def test_smooths_refuse_a_user_supplied_convolve(host, use_dask):
    """spectral_cube.py:2810, 3188 take `convolve=`; a callable other than astropy's two cannot run on the device and
    is refused rather than silently ignored."""
    S, calls = host
    cube = _cube(S, S.DaskSpectralCube if use_dask else S.SpectralCube, shape=(9, 6, 8))
    with pytest.raises(NotImplementedError, match='astropy.convolution.convolve'):
        cube.spectral_smooth(S.Gaussian1DKernel(1.0), convolve=lambda a, k, **kw: a)
    with pytest.raises(NotImplementedError, match='astropy.convolution.convolve'):
        cube.spatial_smooth(S.Gaussian2DKernel(1.0), convolve=lambda a, k, **kw: a)
    with pytest.raises(NotImplementedError, match='nan_treatment'):
        cube.spatial_smooth(S.Gaussian2DKernel(1.0), nan_treatment='fill')
    for fn in (None, _named('convolve'), _named('convolve_fft')):
        assert cube.spectral_smooth(S.Gaussian1DKernel(1.0), convolve=fn).shape == cube.shape
        assert cube.spatial_smooth(S.Gaussian2DKernel(1.0), convolve=fn).shape == cube.shape


def test_projection_header_drops_the_collapsed_axis(tmp_path):
    """lower_dimensional_structures.py:66-97: the header comes from the 2-D WCS; cards of the parent cube's third
    axis must not reach a NAXIS=2 file."""
    import spectral_cube_b200 as S
    from spectral_cube_b200 import io_fits
    parent = {'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN', 'CTYPE3': 'VRAD', 'CRVAL3': 5.0, 'CDELT3': 2.0, 'CRPIX3': 1.0,
              'CUNIT3': 'm/s', 'PC3_3': 1.0, 'PC1_3': 0.0, 'CD3_3': 2.0, 'CROTA3': 0.0, 'NAXIS3': 7, 'NAXIS4': 1, 'CTYPE4': 'STOKES',
              'WCSAXES': 3, 'OBJECT': 'x', 'BMAJ': 0.1, 'RESTFRQ': 1.0e9}
    wcs = S.CubeWCS(**G.ADV_WCS).drop_axis(0)
    p = S.Projection(np.arange(12, dtype=np.float64).reshape(3, 4), unit='K', wcs=wcs, header=parent)
    hdr = p.header
    assert hdr['NAXIS'] == 2 and hdr['NAXIS1'] == 4 and hdr['NAXIS2'] == 3
    for gone in ('CTYPE3', 'CRVAL3', 'CDELT3', 'CRPIX3', 'CUNIT3', 'PC3_3', 'PC1_3', 'CD3_3', 'CROTA3', 'NAXIS3', 'NAXIS4', 'CTYPE4'):
        assert gone not in hdr, gone
    assert hdr['OBJECT'] == 'x' and hdr['BMAJ'] == 0.1 and hdr['RESTFRQ'] == 1.0e9 and hdr['CTYPE1'].startswith('RA')
    path = str(tmp_path / 'm.fits')
    p.write(path)
    with open(path, 'rb') as f:
        back, off = io_fits.read_header(f)
    assert back['NAXIS'] == 2 and 'CTYPE3' not in back and 'CRVAL3' not in back


def test_wcs_accepts_archive_unit_spellings_and_any_projection():
    """wcslib normalises CUNIT strings ('HZ', 'M/S', 'DEG'); moments / smoothing never evaluate the projection, so a
    CAR / ARC / SFL cube must be constructible -- only `reproject` needs the closed forms."""
    import spectral_cube_b200 as S
    kw = dict(G.ADV_WCS)
    w = S.CubeWCS(ctype=['RA---CAR', 'DEC--CAR', 'FREQ'], crval=[10.0, 20.0, 1.4], crpix=[1, 1, 1], cdelt=[-60.0, 60.0, 1.0],
                  cunit=['ARCSEC', 'arcsec', 'GHZ'])
    assert w.cunit == ['deg', 'deg', 'Hz'] and w.cdelt[0] == -60.0 / 3600 and w.crval[2] == 1.4e9
    assert S.CubeWCS(ctype=kw['ctype'], crval=kw['crval'], crpix=kw['crpix'], cdelt=kw['cdelt'], cunit=['DEG', 'Deg', 'M/S']).cunit[2] == 'm/s'
    with pytest.raises(NotImplementedError, match='TAN and SIN'):
        w.celestial_params()
    with pytest.raises(ValueError, match='unknown spectral unit'):
        S.CubeWCS(ctype=kw['ctype'], crval=kw['crval'], crpix=kw['crpix'], cdelt=kw['cdelt'], cunit=['deg', 'deg', 'furlong/fortnight'])
    # a blank CUNIT3 means the SI unit of the axis type (example_cube.fits has none)
    from spectral_cube_b200.wcs import as_cube_wcs
    hdr = {'CTYPE1': 'RA---ARC', 'CTYPE2': 'DEC--ARC', 'CTYPE3': 'VRAD', 'CRVAL3': 7000.0, 'CDELT3': -103.68, 'CRPIX3': 77.6}
    assert as_cube_wcs(hdr).cunit[2] == 'm/s'
