"""
Pins the oracle (oracle/) to the golden vectors the reference's own tests hold for the
hot path.  CPU only.  Test names mirror the reference tests.
"""
import operator
import warnings

import numpy as np
import pytest

from oracle.cube import OracleCube, VarianceWarning
from oracle.wcs import OWCS
from oracle import convolve as oconv
from tests.golden import reference_goldens as G


def moment_cube(use_dask=False, **kw):
    return OracleCube(G.moment_cube_data(), OWCS(**G.MOMENT_WCS), unit='K', use_dask=use_dask, **kw)


def adv_cube(use_dask=False):
    return OracleCube(G.adv_data(), OWCS(**G.ADV_WCS), unit='K', use_dask=use_dask)


def delta_cube(data, use_dask=False, flip=False):
    w = dict(G.ADV_WCS)
    if flip:
        w['cdelt'] = [w['cdelt'][0], w['cdelt'][1], -w['cdelt'][2]]
    return OracleCube(data, OWCS(**w), unit='K', use_dask=use_dask, spectral_unit='km/s')


use_dask = pytest.mark.parametrize('use_dask', [False, True])
axis_order = pytest.mark.parametrize(('axis', 'order'),
                                     [(a, o) for a in (0, 1, 2) for o in (0, 1, 2)])
rtol, atol = 2e-7, 1e-30


# ---- spectral_cube/tests/test_moments.py ------------------------------------------------------
@use_dask
@pytest.mark.parametrize(('order', 'axis', 'how'),
                         [(o, a, h) for o in [0, 1, 2] for a in [0, 1, 2]
                          for h in ['cube', 'slice', 'auto', 'ray']])
def test_reference(order, axis, how, use_dask):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', VarianceWarning)
        mom, unit = moment_cube(use_dask).moment(order=order, axis=axis, how=how)
    np.testing.assert_allclose(mom, G.MOMENTS[order][axis], rtol=1e-7)
    assert unit == G.MOMENT_UNITS[order][axis]


@axis_order
def test_strategies_consistent(axis, order):
    sc = moment_cube()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', VarianceWarning)
        cwise = sc.moment(axis=axis, order=order, how='cube')[0]
        swise = sc.moment(axis=axis, order=order, how='slice')[0]
        rwise = sc.moment(axis=axis, order=order, how='ray')[0]
    np.testing.assert_allclose(cwise, swise, rtol=rtol, atol=atol)
    np.testing.assert_allclose(cwise, rwise, rtol=rtol, atol=atol)


@axis_order
def test_consistent_mask_handling(axis, order):
    sc = moment_cube()
    sc._mask = sc > 4
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', VarianceWarning)
        cwise = sc.moment(axis=axis, order=order, how='cube')[0]
        swise = sc.moment(axis=axis, order=order, how='slice')[0]
        rwise = sc.moment(axis=axis, order=order, how='ray')[0]
        dwise = moment_cube(True, mask=sc._mask).moment(axis=axis, order=order)[0]
    np.testing.assert_allclose(cwise, swise, rtol=rtol, atol=atol)
    np.testing.assert_allclose(cwise, rwise, rtol=rtol, atol=atol)
    np.testing.assert_allclose(cwise, dwise, rtol=rtol, atol=atol)


@use_dask
def test_linewidth(use_dask):
    sc = moment_cube(use_dask)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        np.testing.assert_allclose(sc.moment2()[0], G.MOMENTS[2][0], rtol=1e-7)
    assert len(w) == 1
    assert w[0].category == VarianceWarning
    assert str(w[0].message) == G.VARIANCE_WARNING_TEXT
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        np.testing.assert_allclose(sc.linewidth_sigma()[0], G.MOMENTS[2][0] ** 0.5, rtol=1e-7)
        np.testing.assert_allclose(sc.linewidth_fwhm()[0],
                                   G.MOMENTS[2][0] ** 0.5 * 2.3548200450309493, rtol=1e-7)
    assert len(w) == 0


@use_dask
def test_preserve_unit(use_dask):
    sc_kms = moment_cube(use_dask).with_spectral_unit('km/s')
    m0, u0 = sc_kms.moment0(axis=0)
    m1, u1 = sc_kms.moment1(axis=0)
    np.testing.assert_allclose(m0, G.MOMENTS[0][0] * 1e-3, rtol=1e-7)
    np.testing.assert_allclose(m1, G.MOMENTS[1][0] * 1e-3, rtol=1e-7)
    assert (u0, u1) == ('K km/s', 'km/s')


# ---- spectral_cube/tests/test_spectral_cube.py:1052-1063 ----------------------------------------
@pytest.mark.parametrize('op', (operator.gt, operator.lt, operator.le, operator.ge))
def test_operator_threshold(op):
    c = adv_cube()
    thresh = c._data.ravel()[0]
    c._mask = op(c, thresh)
    np.testing.assert_allclose(c.flattened(), c._data[op(c._data, thresh)])


# ---- spectral_cube/tests/test_spectral_cube.py:2363-2421 ----------------------------------------
@use_dask
def test_spatial_smooth_g2d(use_dask):
    cube_g2d = adv_cube(use_dask).spatial_smooth(oconv.Gaussian2DKernel(3))
    np.testing.assert_almost_equal(cube_g2d._data[0], G.G2D_RESULT0)
    np.testing.assert_almost_equal(cube_g2d._data[2], G.G2D_RESULT2)


@use_dask
def test_spatial_smooth_t2d(use_dask):
    cube_t2d = adv_cube(use_dask).spatial_smooth(oconv.Tophat2DKernel(3))
    np.testing.assert_almost_equal(cube_t2d._data[0], G.T2D_RESULT0)
    np.testing.assert_almost_equal(cube_t2d._data[2], G.T2D_RESULT2)


# ---- spectral_cube/tests/test_regrid.py -----------------------------------------------------------
@use_dask
def test_spectral_smooth(use_dask):
    cube = delta_cube(G.delta_522(), use_dask)
    kernel = oconv.Gaussian1DKernel(1.0)
    result = cube.spectral_smooth(kernel)
    assert kernel.array.size == 9
    np.testing.assert_almost_equal(result._data[:, 0, 0], kernel.array[2:-2], 4)


@use_dask
def test_spectral_interpolate(use_dask):
    cube = delta_cube(G.delta_522(), use_dask)
    sa = cube.spectral_axis
    sg = (sa[1:] + sa[:-1]) / 2.
    result = cube.spectral_interpolate(spectral_grid=sg)
    np.testing.assert_almost_equal(result._data[:, 0, 0], G.INTERP_MIDPOINTS)
    np.testing.assert_allclose(result.spectral_axis, sg)


def test_spectral_interpolate_varying_chunksize():
    cube = delta_cube(G.delta_255(), True)
    sa = cube.spectral_axis
    sg = (sa[1:] + sa[:-1]) / 2.
    result = cube.spectral_interpolate(spectral_grid=sg)
    np.testing.assert_almost_equal(result._data[:, 2, 2], [0.5])


@use_dask
def test_spectral_interpolate_with_fillvalue(use_dask):
    cube = delta_cube(G.delta_522(), use_dask)
    sa = cube.spectral_axis
    sg = sa[0] - (sa[1] - sa[0]) * np.linspace(1, 4, 4)
    result = cube.spectral_interpolate(spectral_grid=sg, fill_value=42)
    np.testing.assert_almost_equal(result._data[:, 0, 0], np.ones(4) * 42)


@use_dask
def test_spectral_interpolate_with_mask(use_dask):
    cube = delta_cube(G.delta_522(), use_dask, flip=True)
    mask = np.ones(cube.shape, dtype=bool)
    mask[:2] = False
    masked_cube = cube.with_mask(mask)
    sa = cube.spectral_axis
    sg = (sa[1:] + sa[:-1]) / 2.
    result = masked_cube.spectral_interpolate(spectral_grid=sg[::-1])
    np.testing.assert_almost_equal(result._data[:, 0, 0], G.INTERP_WITH_MASK)
    # the data of the result are consistent with its own mask where finite
    filled = result.unitless_filled_data[:, 0, 0]
    np.testing.assert_almost_equal(filled, G.INTERP_WITH_MASK)


@use_dask
def test_spectral_interpolate_reversed(use_dask):
    cube = delta_cube(G.delta_522(), use_dask)
    sg = cube.spectral_axis[::-1]
    result = cube.spectral_interpolate(spectral_grid=sg)
    np.testing.assert_almost_equal(sg, result.spectral_axis)


# ---- reproject: no value golden exists in the reference (test_regrid.py:99-135 = shape/WCS) ----
def test_reproject_shape_and_identity():
    cube = adv_cube()
    same = cube.reproject(cube._wcs.copy(), cube.shape)
    np.testing.assert_allclose(same._data, cube._data, rtol=1e-9)
    w = dict(G.ADV_WCS)
    w['crpix'] = [2., 2., 1.]
    w['crval'] = [cube._wcs.celestial_pix2world(0.5, 1.0)[0], cube._wcs.celestial_pix2world(0.5, 1.0)[1],
                  G.ADV_WCS['crval'][2]]
    out = cube.reproject(OWCS(**w), (4, 5, 4))
    assert out.shape == (4, 5, 4)
    # output pixel (1,1) sits at input (y=1.0, x=0.5): mean of the two x-neighbours in row 1
    np.testing.assert_allclose(out._data[:, 1, 1], 0.5 * (cube._data[:, 1, 0] + cube._data[:, 1, 1]),
                               rtol=1e-6)
