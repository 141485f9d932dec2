"""
`Projection.reproject` (``-m gpu``; lower_dimensional_structures.py:496-538): the 2-D image through the same device
kernels as `SpectralCube.reproject`, against the oracle's bilinear restatement (oracle unpinned for the weights, see
DESIGN.md section 2) and against the cube path on the equivalent one-channel cube.
"""
import warnings

import numpy as np
import pytest

from tests.helpers import oracle_cube, gpu_cube, assert_maps_close, RTOL
from tests.test_moments_gpu import BENCH_WCS, _random_cube
from tests.test_regrid_gpu import rotated_header, owcs_from_header

pytestmark = pytest.mark.gpu


def _image_header(hdr3):
    """The 2-D header of the target: the cube header of `rotated_header` without its third axis."""
    out = {k: v for k, v in hdr3.items() if not k.endswith('3') and not k.startswith(('PC3_', 'PC1_3', 'PC2_3'))}
    out['NAXIS'] = 2
    return out


@pytest.mark.parametrize('proj', ['TAN', 'SIN'])
@pytest.mark.parametrize('order', ['bilinear', 'nearest-neighbor'])
@pytest.mark.parametrize('angle,scale', [(0.0, 1.0), (30.0, 1.0), (-75.0, 0.7)])
def test_projection_reproject_matches_oracle_and_the_cube_path(angle, scale, order, proj):
    data = _random_cube((1, 40, 56), seed=int(abs(angle)) + 11, nan_frac=0.02)
    w = dict(BENCH_WCS)
    w['ctype'] = ['RA---' + proj, 'DEC--' + proj, 'VRAD']
    w['crpix'] = [28.5, 20.5, 1.0]
    sc, oc = gpu_cube(data, w), oracle_cube(data, w)
    hdr3 = rotated_header(w, (1, 48, 50), angle, scale, shift=(1.3, -2.1))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        image = sc.moment0()                                  # a Projection carrying the cube's celestial WCS
        got = image.reproject(_image_header(hdr3), order=order)
        # the cube path on the one-channel cube that holds the same (float32-rounded) image
        plane = gpu_cube(np.asarray(image.value, dtype=np.float32)[None], w)
        via_cube = plane.reproject(hdr3, order=order)._data_hi[0].cpu().numpy()
    assert got.shape == (48, 50) and got.dtype == np.float64
    np.testing.assert_array_equal(np.isnan(got.value), np.isnan(via_cube))
    np.testing.assert_array_equal(np.nan_to_num(got.value, nan=-7.0), np.nan_to_num(via_cube, nan=-7.0))
    assert got.wcs.ctype[0].endswith(proj) and got.header['NAXIS1'] == 50 and got.header['NAXIS'] == 2
    if order == 'bilinear':
        oimg = oracle_cube(np.asarray(image.value, dtype=np.float32)[None], w)
        want = oimg.reproject(owcs_from_header(hdr3), (1, 48, 50))._data[0]
        assert_maps_close(got.value, want, rtol=RTOL, atol=1e-9, what='%s %g %g' % (proj, angle, scale))


def test_projection_reproject_needs_a_celestial_image():
    data = _random_cube((6, 8, 12), seed=2)
    sc = gpu_cube(data, BENCH_WCS)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        pv = sc.moment(order=0, axis=1)                        # (spectral, lon): not a celestial image
    with pytest.raises(ValueError, match="two spatial axes"):
        pv.reproject({'NAXIS': 2, 'NAXIS1': 4, 'NAXIS2': 4})
