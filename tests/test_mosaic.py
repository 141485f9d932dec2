"""
mosaic_cubes / combine_headers (spectral_cube/cube_utils.py:744-856).

CPU: the common-grid construction against the oracle restatement and its defining properties.
GPU (``-m gpu``): the reference's own test of the function (spectral_cube/tests/test_regrid.py:602-634: two
overlapping parts of one cube mosaic back to the cube, nearest-neighbour, 3 decimals -- the one value-carrying pin
the reference holds on `reproject`) and the device pipeline against the oracle on random overlapping cubes.
"""
import numpy as np
import pytest

from oracle import mosaic as omosaic
from oracle.wcs import OWCS
from tests.golden import reference_goldens as G
from tests.helpers import oracle_cube, gpu_cube, assert_maps_close, RTOL


def _wkw(crpix, ctype=('RA---TAN', 'DEC--TAN', 'VRAD'), cdelt=(-5.55555561268e-4, 5.55555561268e-4, 1.28821496879), pc=None):
    kw = dict(ctype=list(ctype), crval=[24.0, 30.0, -321.214698632], crpix=list(crpix), cdelt=list(cdelt), cunit=['deg', 'deg', 'km/s'])
    if pc is not None:
        kw['pc'] = pc
    return kw


def test_common_grid_matches_the_oracle_and_contains_every_corner():
    from spectral_cube_b200.mosaic import find_optimal_celestial_wcs
    import spectral_cube_b200 as S
    a = np.radians(20.0)
    rot = [[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]]
    specs = [((40, 56), _wkw([28.5, 20.5, 1.0])), ((33, 47), _wkw([-10.0, 35.0, 1.0], ctype=('RA---SIN', 'DEC--SIN', 'VRAD'))),
             ((25, 30), _wkw([60.0, -5.0, 1.0], cdelt=(-8.0e-4, 8.0e-4, 1.28821496879), pc=rot))]
    got, gshape = find_optimal_celestial_wcs([(s, S.CubeWCS(**kw)) for s, kw in specs])
    want, wshape = omosaic.optimal_celestial_wcs([(s, OWCS(**kw)) for s, kw in specs])
    assert gshape == wshape and got.ctype == ['RA---TAN', 'DEC--TAN']
    np.testing.assert_allclose(got.crval, want.crval[:2], rtol=0, atol=1e-12)
    np.testing.assert_allclose(got.crpix, want.crpix[:2], rtol=0, atol=1e-8)
    np.testing.assert_allclose(got.cdelt, want.cdelt[:2], rtol=1e-15)
    assert abs(got.cdelt[1] - 5.55555561268e-4) < 1e-18                   # the finest input scale
    # every corner of every input falls inside [0.5, naxis + 0.5] of the common grid (FITS pixel counting)
    for (ny, nx), kw in specs:
        w = S.CubeWCS(**kw).celestial()
        lon, lat = w.pix2world(np.array([-0.5, nx - 0.5, nx - 0.5, -0.5]), np.array([-0.5, -0.5, ny - 0.5, ny - 0.5]))
        xp, yp = got.world2pix(lon, lat, origin=1)
        assert xp.min() > 0.5 - 1e-6 and xp.max() < gshape[1] + 0.5 + 0.51 and yp.min() > 0.5 - 1e-6 and yp.max() < gshape[0] + 0.5 + 0.51


def test_combining_a_header_with_itself_keeps_the_grid():
    """tests/test_regrid.py:614: ``combine_headers(cube.header, cube.header)`` is the expected WCS of the mosaic."""
    import spectral_cube_b200 as S
    hdr = dict(S.CubeWCS(**_wkw([28.5, 20.5, 1.0])).to_header(), NAXIS=3, NAXIS1=56, NAXIS2=40, NAXIS3=7, BUNIT='K', CROTA2=0.0)
    out = S.combine_headers(hdr, hdr)
    assert (out['NAXIS1'], out['NAXIS2'], out['NAXIS3']) == (56, 40, 7) and out['BUNIT'] == 'K' and 'CROTA2' not in out
    assert out['CTYPE1'] == 'RA---TAN' and abs(out['CRPIX1'] - 28.5) < 1e-6 and abs(out['CRPIX2'] - 20.5) < 1e-6
    assert abs(out['CRVAL1'] - 24.0) < 1e-12 and out['CDELT1'] == -5.55555561268e-4


def test_oracle_mosaic_of_two_overlapping_parts_restores_the_cube():
    """The reference's test restated on the oracle: tests/test_regrid.py:602-634."""
    data = G.adv_data()
    cube = oracle_cube(data, G.ADV_WCS)
    n = cube.shape[1]
    p1w, p2w = dict(G.ADV_WCS), dict(G.ADV_WCS)
    cut = int(round(n / 3.))
    p2w['crpix'] = [G.ADV_WCS['crpix'][0], G.ADV_WCS['crpix'][1] - cut, G.ADV_WCS['crpix'][2]]
    part1 = oracle_cube(data[:, :int(round(n * 2. / 3.)), :], p1w)
    part2 = oracle_cube(data[:, cut:, :], p2w)
    result, grid = omosaic.mosaic_cubes([part1, part2], order='nearest-neighbor')
    assert result.shape == cube.shape
    np.testing.assert_almost_equal(result, data, decimal=3)


@pytest.mark.gpu
@pytest.mark.parametrize('use_dask', [False, True])
def test_mosaic_cubes_reference_case(use_dask):
    """spectral_cube/tests/test_regrid.py:602-634 through the device pipeline."""
    import spectral_cube_b200 as S
    data = G.adv_data().astype(np.float32)
    cube = gpu_cube(data, G.ADV_WCS, use_dask=use_dask)
    n = cube.shape[1]
    part1 = cube[:, :int(round(n * 2. / 3.)), :]
    part2 = cube[:, int(round(n / 3.)):, :]
    result = S.mosaic_cubes([part1, part2], order='nearest-neighbor', roundtrip_coords=False, spectral_block_size=100)
    assert type(result) is type(cube) and result.shape == cube.shape and result.unit == cube.unit
    expected = S.combine_headers(cube.header, cube.header)
    for key in ('CTYPE1', 'CTYPE2', 'NAXIS1', 'NAXIS2'):
        assert result.header[key] == expected[key], key
    for key in ('CRVAL1', 'CRVAL2', 'CRPIX1', 'CRPIX2', 'CDELT1', 'CDELT2'):
        assert abs(result.header[key] - expected[key]) < 1e-9, key
    np.testing.assert_almost_equal(result.filled_data[:], cube.filled_data[:], decimal=3)


@pytest.mark.gpu
@pytest.mark.parametrize('order', ['bilinear', 'nearest-neighbor'])
def test_mosaic_cubes_matches_oracle(order):
    import spectral_cube_b200 as S
    rng = np.random.default_rng(5)
    a = np.radians(15.0)
    rot = [[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]]
    specs = [((4, 40, 56), _wkw([28.5, 20.5, 1.0])), ((4, 33, 47), _wkw([5.0, 30.0, 1.0], ctype=('RA---SIN', 'DEC--SIN', 'VRAD'))),
             ((4, 25, 30), _wkw([40.0, -2.0, 1.0], pc=rot))]
    gcubes, ocubes = [], []
    for shape, kw in specs:
        d = rng.normal(2.0, 1.0, shape).astype(np.float32)
        d[rng.random(shape) < 0.03] = np.nan
        gcubes.append(gpu_cube(d, kw))
        ocubes.append(oracle_cube(d, kw))
    got = S.mosaic_cubes(gcubes, order=order)
    want, grid = omosaic.mosaic_cubes(ocubes, order=order)
    assert got.shape == want.shape
    np.testing.assert_allclose(got.wcs.crpix[:2], grid.crpix[:2], atol=1e-7)
    gd = got.unmasked_data[:]
    assert gd.dtype == np.float64
    assert_maps_close(gd, want, rtol=RTOL, atol=1e-9, what='mosaic ' + order)
    assert np.isnan(gd).any() and np.isfinite(gd).sum() > gd.size // 3        # uncovered corners: 0 / 0
