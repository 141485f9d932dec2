"""
Sub-cube views (``-m gpu``): `cube[view]` and `spectral_slab` (spectral_cube.py:1290-1380, :1822-1876) return
views of the same device memory whose WCS and mask are sliced alongside; every kernel takes the strides.
Checked against the oracle built on the sliced ARRAY with the sliced header written by hand.
"""
import warnings

import numpy as np
import pytest

from tests.helpers import oracle_cube, gpu_cube, assert_maps_close, RTOL
from tests.test_moments_gpu import BENCH_WCS, _random_cube, quiet

pytestmark = pytest.mark.gpu


def _sliced_wcs(view, shape):
    w = dict(BENCH_WCS)
    crpix, cdelt = list(w['crpix']), list(w['cdelt'])
    for np_axis, sl in enumerate(view):
        start, stop, step = sl.indices(shape[np_axis])
        i = 2 - np_axis
        crpix[i] = (crpix[i] - start - 0.5) / step + 0.5
        cdelt[i] = cdelt[i] * step
    w['crpix'], w['cdelt'] = crpix, cdelt
    return w


@pytest.mark.parametrize('view', [(slice(2, 30), slice(1, 8), slice(3, 14)),          # odd width, row stride != nx
                                  (slice(None), slice(None), slice(4, 20)),           # 16-byte aligned start
                                  (slice(0, 32, 2), slice(None), slice(None)),        # every other channel
                                  (slice(5, 9), slice(2, 3), slice(0, 16))])
def test_moments_and_reductions_of_a_view(view):
    data = _random_cube((32, 9, 24), seed=13)
    sc = gpu_cube(data, BENCH_WCS)
    sc = sc.with_mask(sc > 0.5)
    sub = sc[view]
    assert sub.shape == data[view].shape
    assert sub._data.data_ptr() == sc._data[view].data_ptr()                          # a view, not a copy
    oc = oracle_cube(np.ascontiguousarray(data[view]), _sliced_wcs(view, data.shape))
    oc = oc.with_mask(oc > 0.5)
    for order in (0, 1, 2):
        got = quiet(sub.moment, order=order).value
        want = quiet(oc.moment, order=order, how='cube')[0]
        assert_maps_close(got, want, rtol=RTOL, atol=1e-7 if order == 2 else 0.0, what='moment%d of %r' % (order, view))
    np.testing.assert_array_equal(sub.max(axis=0).value, oc.max(axis=0))
    np.testing.assert_array_equal(sub.mask.include(), oc._mask_include())
    np.testing.assert_allclose(sub.spectral_axis, oc.spectral_axis, rtol=1e-12)


def test_spectral_slab_matches_the_sliced_cube():
    data = _random_cube((40, 6, 16), seed=2)
    sc = gpu_cube(data, BENCH_WCS)
    ax = sc.spectral_axis
    slab = sc.spectral_slab(ax[25] + 0.2 * (ax[1] - ax[0]), ax[10])                  # reversed order, off-centre value
    assert slab.shape == (16, 6, 16)
    np.testing.assert_allclose(slab.spectral_axis, ax[10:26], rtol=1e-12)
    want = quiet(sc[10:26].moment, order=1).value
    np.testing.assert_array_equal(quiet(slab.moment, order=1).value, want)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        sc.spectral_slab(ax[3], ax[3])
    assert len(w) == 1 and 'identical' in str(w[0].message)


def test_smoothing_a_view_leaves_the_parent_untouched():
    import spectral_cube_b200 as scb
    import oracle.convolve as oconv
    data = _random_cube((24, 8, 40), seed=6)
    sc = gpu_cube(data, BENCH_WCS, use_dask=True)
    before = sc._data.clone()
    view = (slice(2, 22), slice(1, 7), slice(4, 36))
    sub = sc[view]
    got = sub.spectral_smooth(scb.Gaussian1DKernel(1.0)).unmasked_data[:]
    oc = oracle_cube(np.ascontiguousarray(data[view]), _sliced_wcs(view, data.shape), use_dask=True)
    want = oc.spectral_smooth(oconv.Gaussian1DKernel(1.0))._data
    assert_maps_close(got, want, rtol=RTOL, what='smoothed view')
    import torch
    assert torch.equal(torch.nan_to_num(sc._data, nan=-1.0), torch.nan_to_num(before, nan=-1.0))


def test_apply_function_parallel_hands_the_filled_cube_to_a_device_function():
    """spectral_cube.py:3049-3159 / dask :501-638 with the device as the one chunk: the callable gets the FILLED cube as a
    CUDA tensor and returns one of the same shape; the mask object is left as it is (:3043-3045).  The reference's own
    use of the seam: tests/test_dask.py:230-252 (a function adding a constant)."""
    import torch
    import pytest
    import numpy as np
    from tests.helpers import gpu_cube
    from tests.test_moments_gpu import BENCH_WCS, _random_cube
    data = _random_cube((6, 5, 8), seed=3, nan_frac=0.1)
    for use_dask in (False, True):
        cube = gpu_cube(data, BENCH_WCS, use_dask=use_dask)
        cube = cube.with_mask(cube > -0.5)

        def func(x, add=None):
            assert isinstance(x, torch.Tensor) and x.is_cuda and tuple(x.shape) == data.shape
            return x + add
        out = cube.apply_function_parallel_spectral(func, add=1, accepts_chunks=True, num_cores=4)
        want = np.where(np.isfinite(data) & (data > -0.5), data, np.nan) + 1
        np.testing.assert_array_equal(out.filled_data[:], want.astype(np.float32))
        assert out.mask is cube.mask and type(out) is type(cube)
        out2 = cube.apply_function_parallel_spatial(lambda x: torch.flip(x, dims=(2,)))
        np.testing.assert_array_equal(np.isnan(out2.unmasked_data[:]), np.isnan(want[:, :, ::-1]))
        with pytest.raises(ValueError, match="must return a CUDA tensor"):
            cube.apply_function_parallel_spectral(lambda x: x[:2])
        with pytest.raises(NotImplementedError, match="no CPU fallback"):
            cube.apply_function_parallel_spectral(lambda x, y: x)
