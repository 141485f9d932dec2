"""
GPU parity tests for the spectral-axis reductions (``-m gpu``): `sum`, `mean`, `std`, `max`, `min`,
`argmax`, `argmin` against the oracle restatement of `apply_numpy_function` (numpy's nan-functions
on the mask-filled cube, spectral_cube.py:361-470 and :578-826) and the reference's own checks
(spectral_cube/tests/test_spectral_cube.py: `test_sum`/`test_max`/`test_argmax`... compare with numpy
on the filled data).  Tolerance: exact for extrema and indices; 1e-5 relative for sum / mean / std,
with an absolute floor of 1e-5 x the spaxel's sum of |values| for the float32-accumulating numpy sum
(the product accumulates in float64 and rounds once).
"""
import warnings
import zlib

import numpy as np
import pytest

from tests.helpers import oracle_cube, gpu_cube, assert_maps_close, RTOL
from tests.test_moments_gpu import BENCH_WCS, MASKS, _random_cube

pytestmark = pytest.mark.gpu


def _pair(shape, maskname, seed_extra=''):
    data = _random_cube(shape, seed=zlib.crc32(repr((shape, maskname, seed_extra)).encode()))
    sc, oc = gpu_cube(data, BENCH_WCS), oracle_cube(data, BENCH_WCS)
    ms, mo = MASKS[maskname](sc), MASKS[maskname](oc)
    if ms is not None:
        sc, oc = sc.with_mask(ms), oc.with_mask(mo)
    return data, sc, oc


@pytest.mark.parametrize('maskname', sorted(MASKS))
@pytest.mark.parametrize('shape', [(32, 9, 16), (33, 7, 13), (40, 5, 6)])
def test_reductions_match_oracle_under_lazy_masks(shape, maskname):
    data, sc, oc = _pair(shape, maskname)
    inc = oc._mask_include() & ~np.isnan(data)
    scale = np.where(inc, np.abs(data), 0).sum(axis=0).astype(np.float64)      # sum of |values| per spaxel
    for name in ('sum', 'mean', 'std'):
        got = getattr(sc, name)(axis=0)
        want = getattr(oc, name)(axis=0)
        assert got.value.dtype == np.float32 and got.unit == 'K' and got.meta['collapse_axis'] == 0
        gn, wn = np.isnan(got.value), np.isnan(want)
        np.testing.assert_array_equal(gn, wn, err_msg=name)
        ok = ~wn
        div = 1.0 if name == 'sum' else np.maximum(inc.sum(axis=0), 1)
        tol = RTOL * np.abs(want[ok]) + RTOL * (scale / div)[ok]
        assert (np.abs(got.value[ok].astype(np.float64) - want[ok]) <= tol).all(), name
    for name in ('max', 'min'):
        got = getattr(sc, name)(axis=0)
        want = getattr(oc, name)(axis=0)
        np.testing.assert_array_equal(got.value, want, err_msg=name)               # exact, NaN where nothing is included
    any_inc = inc.any(axis=0)
    for name in ('argmax', 'argmin'):
        got = getattr(sc, name)(axis=0)
        assert got.dtype == np.int64
        # all-excluded spaxels are "arbitrary" in the reference (and all-NaN ones raise in numpy): compare the rest
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            fill = -np.inf if name == 'argmax' else np.inf
            filled = np.where(inc, data, fill)
            want = getattr(np, name)(filled, axis=0)
        np.testing.assert_array_equal(got[any_inc], want[any_inc], err_msg=name)
        assert (got[~any_inc] == 0).all()


@pytest.mark.parametrize('maskname', ['isfinite', 'gt3', 'xor_not'])
def test_tma_ring_kernel_equals_the_direct_kernel(maskname, monkeypatch):
    """Both kernels on the same cube (the ring is picked on its own only for planes that fill the chip):
    every statistic bit for bit, including a ragged last tile (nx = 520) and a channel count that is not a
    multiple of the slab depth."""
    import torch
    data = _random_cube((37, 5, 520), seed=41)
    res = {}
    for choice in ('1', '2'):
        monkeypatch.setenv('SC_REDUCE_KERNEL', choice)
        sc = gpu_cube(data, BENCH_WCS)
        m = MASKS[maskname](sc)
        if m is not None:
            sc = sc.with_mask(m)
        res[choice] = sc._reduce_axis0_raw({'sum', 'count', 'm2', 'min', 'max', 'argmin', 'argmax'})
    for k in res['1']:
        a, b = res['1'][k], res['2'][k]
        if a.dtype.is_floating_point:
            a, b = torch.nan_to_num(a, nan=-123.0), torch.nan_to_num(b, nan=-123.0)
        assert torch.equal(a, b), k


def test_std_ddof_and_variance_without_cancellation():
    rng = np.random.default_rng(11)
    data = (1.0e4 + rng.normal(0, 1e-2, (64, 8, 12))).astype(np.float32)            # mean >> spread
    sc, oc = gpu_cube(data, BENCH_WCS), oracle_cube(data, BENCH_WCS)
    for ddof in (0, 1):
        want = np.nanstd(data.astype(np.float64), axis=0, ddof=ddof)               # the exact answer
        got = sc.std(axis=0, ddof=ddof).value
        np.testing.assert_allclose(got, want, rtol=1e-5)


def test_first_occurrence_and_ties():
    data = np.zeros((6, 4, 4), dtype=np.float32)
    data[2] = 5.0; data[4] = 5.0; data[1] = -3.0; data[5] = -3.0
    sc = gpu_cube(data, BENCH_WCS)
    assert (sc.argmax(axis=0) == 2).all() and (sc.argmin(axis=0) == 1).all()
    assert (sc.max(axis=0).value == 5.0).all() and (sc.min(axis=0).value == -3.0).all()


@pytest.mark.parametrize('maskname', ['isfinite', 'gt3', 'or'])
def test_whole_cube_reductions(maskname):
    """axis=None (the default of the reference's `sum/mean/std/max/min/argmax/argmin`)."""
    data, sc, oc = _pair((33, 7, 13), maskname, 'whole')
    # the reference applies the nan-function to the filled cube (spectral_cube.py:446-454); float64 here so that
    # the comparison is not limited by numpy's float32 accumulation
    filled = oc._get_filled_data(fill=np.nan).astype(np.float64)
    for name, fn in (('sum', np.nansum), ('mean', np.nanmean), ('std', np.nanstd), ('max', np.nanmax), ('min', np.nanmin)):
        got, want = getattr(sc, name)(), fn(filled)
        assert np.isclose(float(got), float(want), rtol=1e-6, atol=1e-6), (name, got, want)
        assert np.isclose(float(getattr(oc, name)()), float(want), rtol=1e-3, atol=1e-3), name     # the oracle method agrees
    assert np.isclose(float(sc.std(ddof=1)), float(np.nanstd(filled, ddof=1)), rtol=1e-6)
    inc = oc._mask_include() & ~np.isnan(data)
    assert sc.argmax() == int(np.argmax(np.where(inc, data, -np.inf)))
    assert sc.argmin() == int(np.argmin(np.where(inc, data, np.inf)))
    tie = np.zeros((5, 3, 4), dtype=np.float32)
    tie[2, 1, 1] = tie[1, 2, 3] = tie[1, 0, 2] = 9.0
    assert gpu_cube(tie, BENCH_WCS).argmax() == int(np.argmax(tie))            # first in C order among ties
    blank = gpu_cube(np.full((4, 3, 4), np.nan, dtype=np.float32), BENCH_WCS)
    assert np.isnan(blank.sum()) and np.isnan(blank.max()) and np.isnan(blank.std()) and blank.argmax() == 0


def test_a_cube_has_three_axes():
    sc = gpu_cube(np.ones((4, 4, 4), dtype=np.float32), BENCH_WCS)
    assert sc.sum(axis=1).shape == (4, 4) and sc.sum(axis=2).shape == (4, 4)
    with pytest.raises(NotImplementedError):
        sc.sum(axis=3)


def test_peak_and_noise_maps_at_scale():
    """docs/examples.rst:61-93 at bench scale: size-independent properties of one pass."""
    import torch
    from spectral_cube_b200.synth import synth_cube
    import spectral_cube_b200 as scb
    nchan, ny, nx = 256, 512, 1024
    dev = synth_cube(nchan, ny, nx, border=8)
    cube = gpu_cube(dev, dict(BENCH_WCS, crpix=[nx / 2 + 0.5, ny / 2 + 0.5, 1.0]))
    r = cube._reduce_axis0_raw({'sum', 'count', 'm2', 'min', 'max', 'argmin', 'argmax'})
    cnt = r['count']
    fin = torch.isfinite(dev)
    assert torch.equal(cnt, fin.sum(dim=0).to(torch.int32))
    ok = cnt > 0
    filled_hi = torch.where(fin, dev, torch.full_like(dev, -float('inf')))
    filled_lo = torch.where(fin, dev, torch.full_like(dev, float('inf')))
    assert torch.equal(r['max'][ok], filled_hi.max(dim=0).values[ok])
    assert torch.equal(r['min'][ok], filled_lo.min(dim=0).values[ok])
    # the value at the reported channel IS the extremum
    gathered = torch.gather(dev, 0, r['argmax'].to(torch.int64)[None])[0]
    assert torch.equal(gathered[ok], r['max'][ok])
    # linearity of the sum: sum(2 x) == 2 sum(x) exactly in binary floating point
    r2 = gpu_cube(dev * 2, dict(BENCH_WCS, crpix=[nx / 2 + 0.5, ny / 2 + 0.5, 1.0]))._reduce_axis0_raw({'sum', 'm2'})
    assert torch.equal(r2['sum'][ok], 2 * r['sum'][ok])
    assert torch.allclose(r2['m2'][ok], 4 * r['m2'][ok], rtol=1e-12)
    assert torch.isnan(r['sum'][~ok]).all() and torch.isnan(r['max'][~ok]).all()


def test_reductions_without_a_mask_and_with_infinities():
    """`mask=None` (a cube built straight from an array, spectral_cube/tests/test_spectral_cube.py:2686-2700):
    NaNs are skipped like the nan-functions do, infinities take part."""
    import spectral_cube_b200 as scb
    rng = np.random.default_rng(21)
    data = rng.normal(0, 1, (20, 6, 8)).astype(np.float32)
    data[3, 1, 1] = np.nan
    data[5, 2, 2] = np.inf
    data[7, 3, 3] = -np.inf
    data[:, 4, 4] = np.nan
    sc = scb.SpectralCube(data, scb.CubeWCS(**BENCH_WCS), unit='K')            # mask is None
    oc = oracle_cube(data, BENCH_WCS, mask=None)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        np.testing.assert_array_equal(sc.max(axis=0).value, oc.max(axis=0))
        np.testing.assert_array_equal(sc.min(axis=0).value, oc.min(axis=0))
        got, want = sc.sum(axis=0).value, oc.sum(axis=0)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    assert got[2, 2] == np.inf and got[3, 3] == -np.inf
    fin = np.isfinite(want)
    np.testing.assert_allclose(got[fin], want[fin], rtol=1e-5, atol=1e-5)
    assert sc.argmax(axis=0)[2, 2] == 5 and sc.argmin(axis=0)[3, 3] == 7


def test_reductions_on_views_with_odd_widths():
    """A spatial sub-cube view (row stride != nx, odd width): the scalar-load path."""
    data = _random_cube((24, 12, 21), seed=8)
    sc, oc = gpu_cube(data, BENCH_WCS), oracle_cube(data, BENCH_WCS)
    sub = sc[:, 2:11, 3:20]
    osub = oracle_cube(data[:, 2:11, 3:20], BENCH_WCS)
    np.testing.assert_array_equal(sub.max(axis=0).value, osub.max(axis=0))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        assert_maps_close(sub.mean(axis=0).value, osub.mean(axis=0), rtol=1e-5, atol=1e-6)


# ---- along the spatial axes: sc_reduce_spatial ------------------------------------------------------------------
@pytest.mark.parametrize('axis', [1, 2])
@pytest.mark.parametrize('shape,maskname', [((5, 40, 64), 'isfinite'), ((3, 33, 70), 'gt3'), ((2, 7, 13), 'or'),
                                            ((4, 1, 31), 'isfinite'), ((2, 300, 1), 'gt3')])
def test_reductions_along_the_spatial_axes(shape, maskname, axis):
    data, sc, oc = _pair(shape, maskname, 'spatial')
    filled = oc._get_filled_data(fill=np.nan).astype(np.float64)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        cnt = (~np.isnan(filled)).sum(axis=axis)
        want = {'sum': np.where(cnt > 0, np.nansum(filled, axis=axis), np.nan), 'mean': np.nanmean(filled, axis=axis),
                'std': np.nanstd(filled, axis=axis), 'max': np.nanmax(filled, axis=axis), 'min': np.nanmin(filled, axis=axis)}
        for name in ('sum', 'mean', 'std', 'max', 'min'):
            got = getattr(sc, name)(axis=axis)
            assert got.shape == want[name].shape and got.value.dtype == np.float32, name
            assert_maps_close(got.value, want[name], rtol=RTOL, atol=1e-5, what='%s(axis=%d)' % (name, axis))
            # the oracle method (numpy on the float32 cube, like the reference) agrees with the float64 statement
            assert_maps_close(np.asarray(getattr(oc, name)(axis=axis), dtype=np.float64), want[name], rtol=1e-3, atol=1e-3,
                              what='oracle %s' % name)
        assert_maps_close(sc.std(axis=axis, ddof=1).value, np.nanstd(filled, axis=axis, ddof=1), rtol=RTOL, atol=1e-5,
                          what='std ddof=1')
    some = cnt > 0
    inc = ~np.isnan(filled)
    assert np.array_equal(sc.argmax(axis=axis)[some], np.argmax(np.where(inc, filled, -np.inf), axis=axis)[some])
    assert np.array_equal(sc.argmin(axis=axis)[some], np.argmin(np.where(inc, filled, np.inf), axis=axis)[some])
    assert (sc.argmax(axis=axis)[~some] == 0).all()
    assert sc.sum(axis=axis).meta['collapse_axis'] == axis
