"""
Host-side logic that needs no GPU: the kernel discretisation rules (astropy.convolution's published
defaults, pinned through the reference goldens in test_oracle_goldens.py on the oracle side), the WCS
stand-in against the oracle's FITS-WCS evaluator, header round trips, and the `Projection` container
(lower_dimensional_structures.py:246-292).
"""
import numpy as np
import pytest

import oracle.convolve as oconv
from oracle.wcs import OWCS
from spectral_cube_b200 import kernels as K
from spectral_cube_b200.projection import Projection
from spectral_cube_b200.wcs import CubeWCS, as_cube_wcs

WCS = dict(ctype=['RA---TAN', 'DEC--TAN', 'VRAD'], crval=[24.0, 30.0, -321214.698632], crpix=[8.5, 9.5, 1.0],
           cdelt=[-5.55555561268e-4, 5.55555561268e-4, 1288.21496879], cunit=['deg', 'deg', 'm/s'])


@pytest.mark.parametrize('sigma', [0.4, 1.0, 5 / 2.3548200450309493, 3.3, 8.0])
def test_gaussian1d_matches_the_oracle_and_astropys_size_rule(sigma):
    a, b = K.Gaussian1DKernel(sigma).array, oconv.Gaussian1DKernel(sigma).array
    assert a.dtype == np.float64 and a.shape == b.shape and a.size % 2 == 1
    np.testing.assert_allclose(a, b, rtol=1e-14, atol=0)          # (x / s)^2 against x^2 / s^2: an ulp
    assert a.size == int(np.ceil(8 * sigma)) + (1 - int(np.ceil(8 * sigma)) % 2)       # ceil(8 sigma) rounded up to odd
    assert K.Gaussian1DKernel(1.0).array.size == 9                                     # tests/test_regrid.py:152


@pytest.mark.parametrize('args', [(1.0,), (8 / 2.3548200450309493,), (1.0, 2.0), (1.0, 2.0, 0.5)])
def test_gaussian2d_and_tophat_match_the_oracle(args):
    kw = dict(zip(('x_stddev', 'y_stddev', 'theta'), args))
    a, b = K.Gaussian2DKernel(**kw).array, oconv.Gaussian2DKernel(**kw).array
    assert a.shape == b.shape and all(n % 2 == 1 for n in a.shape)
    np.testing.assert_allclose(a, b, rtol=1e-13, atol=0)           # different association of the exponent: a few ulps
    np.testing.assert_allclose(K.Tophat2DKernel(3).array, oconv.Tophat2DKernel(3).array, rtol=1e-14, atol=0)
    assert np.array_equal(K.Tophat2DKernel(3).array > 0, oconv.Tophat2DKernel(3).array > 0)
    np.testing.assert_array_equal(K.Box1DKernel(5).array, oconv.Box1DKernel(5).array)
    if len(args) == 1:
        assert a.shape == (K.Gaussian1DKernel(args[0]).array.size,) * 2


def test_cube_wcs_agrees_with_the_oracle_evaluator():
    w, o = CubeWCS(**WCS), OWCS(**WCS)
    pz = np.arange(12)
    np.testing.assert_allclose(w.spectral_pix2world(pz), o.spectral_pix2world(pz), rtol=1e-15)
    np.testing.assert_allclose(w.pixel_scale_matrix, o.pixel_scale_matrix, rtol=1e-15)
    # header round trip, CD-matrix form, duck typing
    hdr = w.to_header()
    w2 = CubeWCS.from_header(hdr)
    for a in ('crval', 'crpix', 'cdelt'):
        np.testing.assert_array_equal(getattr(w, a), getattr(w2, a))
    assert list(w2.ctype) == list(w.ctype)
    cd = dict(hdr)
    for i in (1, 2, 3):
        cd.pop('CDELT%d' % i)
        cd['CD%d_%d' % (i, i)] = float(w.cdelt[i - 1])
    np.testing.assert_allclose(CubeWCS.from_header(cd).pixel_scale_matrix, w.pixel_scale_matrix, rtol=1e-15)
    np.testing.assert_array_equal(as_cube_wcs(o).crpix, w.crpix)
    assert as_cube_wcs(w) is w
    with pytest.raises(TypeError):
        as_cube_wcs(3.0)
    # dropping the spectral axis keeps the celestial cards (wcs_utils.py:28-45)
    c = w.drop_axis(0).to_header()
    assert c['CTYPE1'] == 'RA---TAN' and c['CRPIX2'] == 9.5 and 'CTYPE3' not in c


def test_projection_container():
    w = CubeWCS(**WCS)
    p = Projection(np.arange(6.0).reshape(2, 3) - 1.0, unit='km/s2', wcs=w.drop_axis(0), meta={'moment_order': 2}, header={'OBJECT': 'x'})
    assert p.unit == 'km/s2' and p.meta['moment_order'] == 2 and p.shape == (2, 3)
    np.testing.assert_array_equal(p.value, np.arange(6.0).reshape(2, 3) - 1.0)
    s = p.sqrt()                                               # linewidth_sigma: sqrt of the variance map, NaN for negatives
    assert s.unit == 'km/s' and np.isnan(s.value[0, 0]) and s.value[1, 2] == 2.0
    h = p.header
    assert h['NAXIS'] == 2 and h['NAXIS1'] == 3 and h['NAXIS2'] == 2 and h['BUNIT'] == 'km/s2' and h['CTYPE1'] == 'RA---TAN'
    with pytest.raises(ValueError):
        Projection(np.zeros(3))


@pytest.mark.parametrize('shape,kshape,nw', [((70, 40), (5, 7), 8), ((20, 20), (17, 11), 8), ((45, 33), (9, 5), 4)])
def test_model_of_the_tiled_direct_kernel_matches_the_oracle(shape, kshape, nw):
    """direct2d_tiled_kernel (opt-in, csrc/spatial_smooth.cu) restated thread by thread in Python: its box staging,
    register-window walk and kernel flip give astropy's convolution (asymmetric random kernel, NaN holes)."""
    from tools.dryrun.model_direct2d_tiled import tiled
    rng = np.random.default_rng(0)
    img = rng.normal(size=shape).astype(np.float32)
    img[rng.random(img.shape) < 0.1] = np.nan
    img[3:3 + kshape[0] + 4, 5:5 + kshape[1] + 6] = np.nan
    kernel = rng.random(kshape) + 0.01
    want = oconv.convolve(img.astype(np.float64), kernel, normalize_kernel=True)
    got = tiled(img, kernel, nw=nw)
    assert np.isnan(want).any() and np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-6, atol=1e-7)
