"""
GPU parity tests for the moment engine (run on the B200 box with ``-m gpu``).  They call the
product (``spectral_cube_b200`` -> C ABI -> CUDA) and check it against (i) the reference's
golden vectors (spectral_cube/tests/test_moments.py:19-48), (ii) the CPU oracle on seeded
inputs, (iii) size-independent properties at benchmark scale.  Test names mirror the
reference's tests.  Tolerance: 1e-5 relative (north star), stated per assertion.
"""
import operator
import warnings
import zlib

import numpy as np
import pytest

from tests.golden import reference_goldens as G
from tests.helpers import oracle_cube, gpu_cube, assert_maps_close, RTOL

pytestmark = pytest.mark.gpu

BENCH_WCS = dict(ctype=['RA---TAN', 'DEC--TAN', 'VRAD'], crval=[24.0, 30.0, -321.214698632],
                 crpix=[8.5, 8.5, 1.0], cdelt=[-5.55555561268e-4, 5.55555561268e-4, 1.28821496879],
                 cunit=['deg', 'deg', 'km/s'])


def quiet(f, *a, **k):
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return f(*a, **k)


# ---- spectral_cube/tests/test_moments.py --------------------------------------------------------
@pytest.mark.parametrize('use_dask', [False, True])
@pytest.mark.parametrize('how', ['cube', 'slice', 'auto', 'ray'])
@pytest.mark.parametrize('order', [0, 1, 2])
def test_reference(order, how, use_dask):
    sc = gpu_cube(G.moment_cube_data(), G.MOMENT_WCS, use_dask=use_dask)
    mom = quiet(sc.moment, order=order, axis=0, how=how)
    np.testing.assert_allclose(mom.value, G.MOMENTS[order][0], rtol=1e-7)     # the reference's own rtol
    assert mom.unit == G.MOMENT_UNITS[order][0]
    assert mom.meta['moment_order'] == order and mom.meta['moment_axis'] == 0
    assert ('moment_method' in mom.meta) == (not use_dask)


@pytest.mark.parametrize('order', [0, 1, 2])
def test_consistent_mask_handling(order):
    sc = gpu_cube(G.moment_cube_data(), G.MOMENT_WCS)
    sc._mask = sc > 4
    oc = oracle_cube(G.moment_cube_data(), G.MOMENT_WCS)
    oc._mask = oc > 4
    got = quiet(sc.moment, order=order, axis=0).value
    want = quiet(oc.moment, order=order, axis=0, how='cube')[0]
    assert_maps_close(got, want, rtol=2e-7, what='masked moment%d' % order)


def test_linewidth():
    sc = gpu_cube(G.moment_cube_data(), G.MOMENT_WCS)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        sc.moment2()
    assert len(w) == 1 and str(w[0].message) == G.VARIANCE_WARNING_TEXT
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        sigma = sc.linewidth_sigma()
        fwhm = sc.linewidth_fwhm()
    assert len(w) == 0
    np.testing.assert_allclose(sigma.value, np.sqrt(G.MOMENTS[2][0]), rtol=1e-7)
    np.testing.assert_allclose(fwhm.value, np.sqrt(G.MOMENTS[2][0]) * 2.3548200450309493, rtol=1e-7)
    assert sigma.unit == 'm/s' and fwhm.unit == 'm/s'


def test_invalid_how_returns_valueerror_like_the_reference():
    sc = gpu_cube(G.moment_cube_data(), G.MOMENT_WCS)
    assert isinstance(sc.moment(order=0, how='bogus'), ValueError)          # spectral_cube.py:1687-1689


def test_preserve_unit():
    sc = gpu_cube(G.moment_cube_data(), G.MOMENT_WCS).with_spectral_unit('km/s')
    m0, m1, m2 = (quiet(sc.moment, order=o) for o in (0, 1, 2))
    assert (m0.unit, m1.unit, m2.unit) == ('K km/s', 'km/s', 'km/s2')
    np.testing.assert_allclose(m1.value, G.MOMENTS[1][0] / 1e3, rtol=1e-7)


def test_higher_order_moment_matches_oracle():
    rng = np.random.default_rng(5)
    data = rng.random((16, 6, 8)).astype(np.float32) + 0.1
    sc, oc = gpu_cube(data, BENCH_WCS), oracle_cube(data, BENCH_WCS)
    for order in (3, 4):
        assert_maps_close(sc.moment(order=order).value, oc.moment(order=order, how='cube')[0],
                          rtol=RTOL, what='order %d' % order)


# ---- oracle parity on seeded random cubes -----------------------------------------------------------
def _random_cube(shape, seed, nan_frac=0.05, inf=False):
    rng = np.random.default_rng(seed)
    nchan = shape[0]
    c = np.arange(nchan)[:, None, None]
    amp = rng.uniform(0, 10, shape[1:])
    cen = rng.uniform(nchan / 4, 3 * nchan / 4, shape[1:])
    data = amp * np.exp(-0.5 * ((c - cen) / 3.0) ** 2) + rng.normal(0, 1, shape)
    data = data.astype(np.float32)
    data[rng.random(shape) < nan_frac] = np.nan
    data[:, 0, :] = np.nan                      # fully blank rays
    if inf:
        data[3, 2, 2] = np.inf
    return data


MASKS = {
    'isfinite': lambda c: None,
    'gt3': lambda c: c > 3.0,
    'ge_data_value': lambda c: c >= float(np.float32(1.2345)),
    'lt1': lambda c: c < 1.0,
    'le_then_gt': lambda c: (c <= 5.0) & (c > 0.5),
    'or': lambda c: (c > 4.0) | (c < -1.0),
    'xor_not': lambda c: ~((c > 2.0) ^ (c < 3.0)),
    'ne': lambda c: c != 0.0,
}


@pytest.mark.parametrize('maskname', sorted(MASKS))
@pytest.mark.parametrize('shape', [(32, 9, 16), (33, 7, 13), (40, 5, 6)])
def test_moments_match_oracle_under_lazy_masks(shape, maskname):
    data = _random_cube(shape, seed=zlib.crc32(repr((shape, maskname)).encode()))
    sc, oc = gpu_cube(data, BENCH_WCS), oracle_cube(data, BENCH_WCS)
    ms, mo = MASKS[maskname](sc), MASKS[maskname](oc)
    if ms is not None:
        sc, oc = sc.with_mask(ms), oc.with_mask(mo)
    got = sc.moments012()
    for order in (0, 1, 2):
        want = quiet(oc.moment, order=order, how='cube')[0]
        # These masks give no positivity guarantee: positive and negative noise cancel in sum(w), so the honest error
        # bound of ANY float64 evaluation (the reference's included) is eps * cond with cond = sum|w| / |sum w|.
        # Both sides accumulate in float64 (eps = 1.1e-16 per operation, <= 40 channels, different summation
        # orders and a different origin for x): 1e-10 * cond leaves a factor ~1e4 for that and is still 1e5 times
        # tighter than the float32-level RTOL -- a float32 accumulator anywhere on the path fails it.
        F64_TOL = 1e-10
        w = np.where(oc._mask_include() & np.isfinite(data), data, 0.0).astype(np.float64)
        cond = np.abs(w).sum(0) / np.maximum(np.abs(w.sum(0)), 1e-300)
        x = oc._pix_cen()[0][:, 0, 0]
        scale = {0: 0.0, 1: np.ptp(x), 2: np.ptp(x) ** 2}[order]      # sum(w x^k) / sum(w): absolute error eps cond range^k
        g, wnt = got[order].value, want
        assert np.array_equal(np.isnan(g), np.isnan(wnt)), (maskname, order)
        ok = np.isfinite(wnt)
        tol = F64_TOL * np.maximum(cond[ok], 1.0) * (np.abs(wnt[ok]) + scale)
        assert np.all(np.abs(g[ok] - wnt[ok]) <= tol), \
            (maskname, order, float((np.abs(g[ok] - wnt[ok]) / tol).max()))
        single = quiet(sc.moment, order=order).value
        np.testing.assert_array_equal(single, g)          # fused pass == single-order pass, bit for bit


@pytest.mark.parametrize('use_dask', [False, True])
def test_positive_weights_meet_plain_rtol(use_dask):
    """Under a >3 sigma mask all weights are positive: plain elementwise rtol=1e-5 must hold."""
    data = _random_cube((64, 12, 20), seed=11)
    sc = gpu_cube(data, BENCH_WCS, use_dask=use_dask)
    oc = oracle_cube(data, BENCH_WCS, use_dask=use_dask)
    sc, oc = sc.with_mask(sc > 3.0), oc.with_mask(oc > 3.0)
    for order in (0, 1, 2):
        want = quiet(oc.moment, order=order)[0]
        got = quiet(sc.moment, order=order).value
        # single-voxel rays have M2 ~ 1e-26 in the reference: allow an absolute floor of 1e-12 channel^2
        atol = 1e-12 * (1.28821496879e3) ** 2 if order == 2 else 0.0
        assert_maps_close(got, want, rtol=RTOL, atol=atol, what='moment%d' % order)
    sig_o = np.sqrt(np.clip(quiet(oc.moment, order=2)[0], 0, None))
    sig_g = sc.linewidth_sigma().value
    assert np.array_equal(np.isnan(sig_g), np.isnan(sig_o))


def test_mask_none_skips_nan_like_the_cube_path():
    data = _random_cube((24, 6, 8), seed=3)
    sc, oc = gpu_cube(data, BENCH_WCS, mask=None), oracle_cube(data, BENCH_WCS, mask=None)
    assert sc.mask is None
    for order in (0, 1):
        want = oc.moment(order=order, how='cube')[0]
        got = sc.moment(order=order).value
        assert np.array_equal(np.isnan(got), np.isnan(want))


def test_boolean_array_and_broadcast_masks():
    data = _random_cube((20, 8, 12), seed=8)
    rng = np.random.default_rng(0)
    full = rng.random(data.shape) > 0.4
    plane = rng.random((1,) + data.shape[1:]) > 0.3
    chan = (np.arange(20) % 3 != 0)[:, None, None]
    thr_plane = rng.uniform(0, 2, data.shape[1:])
    for name, (mk_s, mk_o) in {
        'full': (lambda c: full, lambda c: full),
        'plane': (lambda c: plane, lambda c: np.broadcast_to(plane, data.shape)),
        'chan': (lambda c: chan, lambda c: np.broadcast_to(chan, data.shape)),
        'cmp_plane': (lambda c: c > thr_plane, lambda c: c > thr_plane),
    }.items():
        sc, oc = gpu_cube(data, BENCH_WCS), oracle_cube(data, BENCH_WCS)
        sc, oc = sc.with_mask(mk_s(sc)), oc.with_mask(mk_o(oc))
        np.testing.assert_array_equal(sc.mask.include(), oc._mask_include(), err_msg=name)
        want = oc.moment(order=0, how='cube')[0]
        assert_maps_close(sc.moment0().value, want, rtol=RTOL, atol=1e-9, what=name)


def test_comparison_threshold_is_float64_exact():
    """spectral_cube/tests/test_spectral_cube.py:1052-1063: a threshold equal to a data value must
    separate > from >= ; and a float64 threshold between two float32 values must not round."""
    v = np.float32(1.1)
    data = np.full((4, 2, 4), v, dtype=np.float32)
    data[1] = np.nextafter(v, np.float32(2))
    data[2] = np.nextafter(v, np.float32(0))
    for thr in (float(v), float(v) + 1e-9, float(v) - 1e-9):
        for op in (operator.gt, operator.ge, operator.lt, operator.le, operator.eq, operator.ne):
            sc, oc = gpu_cube(data, BENCH_WCS), oracle_cube(data, BENCH_WCS)
            sc, oc = sc.with_mask(op(sc, thr)), oc.with_mask(op(oc, thr))
            np.testing.assert_array_equal(sc.mask.include(), oc._mask_include(), err_msg=str((thr, op)))
            assert_maps_close(sc.moment0().value, oc.moment(order=0, how='cube')[0], rtol=1e-12,
                              what=str((thr, op)))


def test_strided_view_and_unaligned_widths():
    base = _random_cube((24, 10, 37), seed=21)
    for sl in [(slice(None), slice(None), slice(None)), (slice(2, 20), slice(1, 9), slice(3, 30)),
               (slice(None), slice(None), slice(0, 36)), (slice(None), slice(None), slice(1, 33))]:
        import torch
        dev = torch.from_numpy(base).cuda()[sl]
        sc = gpu_cube(dev, BENCH_WCS)
        oc = oracle_cube(base[sl], BENCH_WCS)
        for order in (0, 1):
            assert_maps_close(sc.moment(order=order).value, oc.moment(order=order, how='cube')[0],
                              rtol=RTOL, atol=1e-7, what=str(sl))


# ---- the synthetic benchmark cube: GPU generator == numpy twin, then parity on its voxels -------
def test_synthetic_generator_is_bit_identical_to_the_numpy_twin():
    from spectral_cube_b200.synth import synth_cube
    from oracle.synth import synth_block
    got = synth_cube(48, 16, 32, y0=8, x0=32, ny_total=64, nx_total=96, border=2, nan_permille=3).cpu().numpy()
    want = synth_block(48, 16, 32, y0=8, x0=32, ny_total=64, nx_total=96, border=2, nan_permille=3)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize('kernel', ['tma', 'direct'])
def test_benchmark_block_matches_oracle(kernel, monkeypatch):
    """A 64-row block of the config-2 style cube (1024 channels, width 2048), >3 sigma mask."""
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs
    from oracle.synth import synth_block
    import spectral_cube_b200 as scb
    monkeypatch.setenv('SC_MOM_KERNEL', '2' if kernel == 'tma' else '1')
    nchan, ny, nx = 1024, 8, 2048
    dev = synth_cube(nchan, ny, nx, y0=100, ny_total=2048, nx_total=2048, border=51)
    host = synth_block(nchan, ny, nx, y0=100, ny_total=2048, nx_total=2048, border=51)
    assert np.array_equal(dev.cpu().numpy().view(np.uint32), host.view(np.uint32))
    w = benchmark_wcs(nchan, 2048, 2048)
    wkw = dict(ctype=w.ctype, crval=[24.0, 30.0, -321.214698632], crpix=list(w.crpix),
               cdelt=[-5.55555561268e-4, 5.55555561268e-4, 1.28821496879], cunit=['deg', 'deg', 'km/s'])
    sc = scb.SpectralCube(dev, w, unit='K')
    sc._mask = scb.LazyMask(np.isfinite, cube=sc)
    sc = sc.with_mask(sc > 3.0)
    oc = oracle_cube(host, wkw)
    oc = oc.with_mask(oc > 3.0)
    got = sc.moments012()
    for order in (0, 1, 2):
        want = quiet(oc.moment, order=order, how='slice')[0]
        atol = 1e-12 * (1.28821496879e3) ** 2 if order == 2 else 0.0
        assert_maps_close(got[order].value, want, rtol=RTOL, atol=atol, what='%s moment%d' % (kernel, order))


def test_properties_at_scale():
    """Size-independent checks on a cube too big for the oracle (512 x 1024 x 2048 = 4.3 GB)."""
    import torch
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs
    import spectral_cube_b200 as scb
    nchan, ny, nx = 512, 1024, 2048
    dev = synth_cube(nchan, ny, nx, border=16)
    w = benchmark_wcs(nchan, ny, nx)
    sc = scb.SpectralCube(dev, w, unit='K')
    sc._mask = scb.LazyMask(np.isfinite, cube=sc)
    m0 = sc.moment0().value
    # (1) linearity of M0: scaling the data scales M0 exactly (power of two)
    sc2 = scb.SpectralCube(dev * 2.0, w, unit='K')
    sc2._mask = scb.LazyMask(np.isfinite, cube=sc2)
    np.testing.assert_array_equal(sc2.moment0().value, 2.0 * m0)
    # (2) M0 equals dv * (float64 column sums done by an independent torch reduction)
    ref = torch.nan_to_num(dev.double(), nan=0.0).sum(0).cpu().numpy() * sc._pix_size_slice(0)
    ok = np.isfinite(m0)
    assert ok.sum() == (ny - 32) * (nx - 32)
    np.testing.assert_allclose(m0[ok], ref[ok], rtol=1e-9, atol=1e-6)
    # (3) a constant velocity shift of the axis shifts M1 by the same amount and leaves M2 alone
    masked = sc.with_mask(sc > 3.0)
    a0, a1, a2 = masked.moments012()
    w2 = benchmark_wcs(nchan, ny, nx)
    w2.crval[2] += 12345.0
    shifted = scb.SpectralCube(dev, w2, unit='K')
    shifted._mask = scb.LazyMask(np.isfinite, cube=shifted)
    shifted = shifted.with_mask(shifted > 3.0)
    b0, b1, b2 = shifted.moments012()
    np.testing.assert_array_equal(a0.value, b0.value)
    okm = np.isfinite(a1.value)
    np.testing.assert_allclose(b1.value[okm] - a1.value[okm], 12345.0, rtol=0, atol=1e-6)
    np.testing.assert_allclose(b2.value[okm], a2.value[okm], rtol=1e-9, atol=1e-3)
    # (4) M2 >= 0 wherever the mask keeps only positive weights, up to rounding
    assert np.nanmin(a2.value) > -1e-3


# ---- moments along the spatial axes (spectral_cube/tests/test_moments.py:22-27, 33-38, 43-48) -------
@pytest.mark.parametrize('use_dask', [False, True])
@pytest.mark.parametrize('axis', [1, 2])
@pytest.mark.parametrize('order', [0, 1, 2])
def test_reference_spatial_axes(order, axis, use_dask):
    sc = gpu_cube(G.moment_cube_data(), G.MOMENT_WCS, use_dask=use_dask)
    mom = quiet(sc.moment, order=order, axis=axis)
    np.testing.assert_allclose(mom.value, G.MOMENTS[order][axis], rtol=1e-7)
    assert mom.unit == G.MOMENT_UNITS[order][axis]


@pytest.mark.parametrize('axis', [1, 2])
def test_spatial_axis_moments_match_oracle(axis):
    data = _random_cube((6, 40, 70), seed=90 + axis, nan_frac=0.03) + np.float32(5.0)
    w = dict(BENCH_WCS)
    w['crpix'] = [30.0, 22.0, 1.0]
    sc, oc = gpu_cube(data, w), oracle_cube(data, w)
    sc, oc = sc.with_mask(sc > 2.0), oc.with_mask(oc > 2.0)
    for order in (0, 1, 2):
        want = quiet(oc.moment, order=order, axis=axis, how='cube')[0]
        got = quiet(sc.moment, order=order, axis=axis).value
        assert_maps_close(got, want, rtol=RTOL, atol=1e-16, what='axis %d order %d' % (axis, order))


def test_host_pipeline_matches_device_path():
    """sc_moments_axis0_host (pinned/pageable host cube streamed in row blocks) == resident-cube path."""
    import ctypes as C
    import torch
    from spectral_cube_b200 import _lib
    from spectral_cube_b200.masks import lower_mask
    lib = _lib.load()
    data = _random_cube((40, 37, 96), seed=123)
    sc = gpu_cube(data, BENCH_WCS)
    sc = sc.with_mask(sc > 1.0)
    ref = [m.value for m in sc.moments012()]
    desc, keep = lower_mask(sc._mask, sc._data)
    for i in range(desc.n_nodes):                     # the host pipeline sees the cube as "self"
        desc.nodes[i].data = None
    host = np.ascontiguousarray(data)
    outs = [np.empty((37, 96)) for _ in range(3)]
    xoff, xptr = _lib.as_double_array(sc._spectral_offsets())
    for staging in (0, 40 * 96 * 4 * 2 * 5):          # default blocks, and 5-row blocks
        _lib.check(lib.sc_moments_axis0_host(host.ctypes.data, 40, 37, 96, 37 * 96, 96, desc, xptr,
                                             float(sc._pix_size_slice(0)), sc._world0_spectral(), 7,
                                             outs[0].ctypes.data, outs[1].ctypes.data, outs[2].ctypes.data,
                                             staging, torch.cuda.current_device()))
        for o in range(3):
            np.testing.assert_array_equal(outs[o], ref[o])
