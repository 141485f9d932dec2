"""
GPU parity tests for spectral_interpolate and reproject.  Goldens: spectral_cube/tests/
test_regrid.py:234-248, 292-303, 318-345, 350-361 (interpolation) and :99-135 (reproject: shape
and WCS only -- the reference pins no bilinear values, so reproject is checked against the
oracle restatement, whose own parity is unpinned: see oracle/reproject.py).
"""
import warnings

import numpy as np
import pytest

from oracle.wcs import OWCS
from tests.golden import reference_goldens as G
from tests.helpers import oracle_cube, gpu_cube, assert_maps_close, RTOL
from tests.test_moments_gpu import BENCH_WCS, _random_cube
from tests.test_smooth_gpu import delta_pair

pytestmark = pytest.mark.gpu
use_dask = pytest.mark.parametrize('use_dask', [False, True])


def data_of(cube):
    return cube.unmasked_data[:]


# ---- spectral_cube/tests/test_regrid.py ---------------------------------------------------------------
@use_dask
def test_spectral_interpolate(use_dask):
    cube, _ = delta_pair(G.delta_522(), use_dask)
    sa = cube.spectral_axis
    sg = (sa[1:] + sa[:-1]) / 2.
    result = cube.spectral_interpolate(spectral_grid=sg)
    np.testing.assert_almost_equal(data_of(result)[:, 0, 0], G.INTERP_MIDPOINTS)
    np.testing.assert_allclose(result.spectral_axis, sg)
    assert result.shape == (4, 2, 2)


def test_spectral_interpolate_varying_chunksize():
    cube, _ = delta_pair(G.delta_255(), True)
    sa = cube.spectral_axis
    sg = (sa[1:] + sa[:-1]) / 2.
    result = cube.spectral_interpolate(spectral_grid=sg)
    np.testing.assert_almost_equal(data_of(result)[:, 2, 2], [0.5])


@use_dask
def test_spectral_interpolate_with_fillvalue(use_dask):
    cube, _ = delta_pair(G.delta_522(), use_dask)
    sa = cube.spectral_axis
    sg = sa[0] - (sa[1] - sa[0]) * np.linspace(1, 4, 4)
    result = cube.spectral_interpolate(spectral_grid=sg, fill_value=42)
    np.testing.assert_almost_equal(data_of(result)[:, 0, 0], np.ones(4) * 42)


@use_dask
def test_spectral_interpolate_with_mask(use_dask):
    cube, _ = delta_pair(G.delta_522(), use_dask, flip=True)
    mask = np.ones(cube.shape, dtype=bool)
    mask[:2] = False
    masked_cube = cube.with_mask(mask)
    sa = cube.spectral_axis
    sg = (sa[1:] + sa[:-1]) / 2.
    result = masked_cube.spectral_interpolate(spectral_grid=sg[::-1])
    np.testing.assert_almost_equal(data_of(result)[:, 0, 0], G.INTERP_WITH_MASK)


@use_dask
def test_spectral_interpolate_reversed(use_dask):
    cube, _ = delta_pair(G.delta_522(), use_dask)
    sg = cube.spectral_axis[::-1]
    result = cube.spectral_interpolate(spectral_grid=sg)
    np.testing.assert_almost_equal(sg, result.spectral_axis)


def test_smoothing_warning_when_decimating():
    import spectral_cube_b200 as scb
    cube, _ = delta_pair(np.zeros((12, 2, 2)), False)
    sa = cube.spectral_axis
    with pytest.warns(scb.SmoothingWarning, match="too small a spacing"):
        cube.spectral_interpolate(sa[::3])


# ---- oracle parity ---------------------------------------------------------------------------------------
GRIDS = {
    'midpoints': lambda sa: (sa[1:] + sa[:-1]) / 2.,
    'decimate2_aligned': lambda sa: sa[::2],                               # every sample is an exact knot hit
    'upsample': lambda sa: np.linspace(sa[0], sa[-1], 2 * sa.size + 3),
    'beyond_both_ends': lambda sa: np.linspace(sa[0] - 3.3 * (sa[1] - sa[0]), sa[-1] + 2.1 * (sa[1] - sa[0]), sa.size),
    'sub_range': lambda sa: np.linspace(sa[5], sa[11], 9),
    'reversed_out': lambda sa: np.linspace(sa[-2], sa[1], sa.size + 1),
}


@use_dask
@pytest.mark.parametrize('kernel', ['direct', 'tma'])
@pytest.mark.parametrize('flip_in', [False, True])
@pytest.mark.parametrize('fill_value', [None, -5.0])
@pytest.mark.parametrize('gridname', sorted(GRIDS))
def test_spectral_interpolate_matches_oracle(gridname, fill_value, flip_in, kernel, use_dask, monkeypatch):
    # both device kernels: the direct one (small / unaligned planes) and `spectral_interp_tma_kernel`, the one every
    # benchmark shape runs (auto-selected only from 148 row tiles on; forced here through the library's switch)
    monkeypatch.setenv('SC_INTERP_KERNEL', '1' if kernel == 'direct' else '2')
    data = _random_cube((24, 5, 12), seed=41, nan_frac=0.08)
    w = dict(BENCH_WCS)
    if flip_in:
        w['cdelt'] = [w['cdelt'][0], w['cdelt'][1], -w['cdelt'][2]]
    sc = gpu_cube(data, w, use_dask=use_dask, spectral_unit='km/s')
    oc = oracle_cube(data, w, use_dask=use_dask, spectral_unit='km/s')
    # a few fully masked spaxels and a partial mask on top of isfinite
    m = np.ones(data.shape, dtype=bool)
    m[:, 1, 3] = False
    m[4:9, 2, :] = False
    sc, oc = sc.with_mask(m), oc.with_mask(m)
    sa = np.sort(oc.spectral_axis)
    grid = GRIDS[gridname](sa)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        got = sc.spectral_interpolate(grid, fill_value=fill_value)
        want = oc.spectral_interpolate(grid, fill_value=fill_value)
    gd, wd = data_of(got), want._data
    assert gd.dtype == wd.dtype, (gd.dtype, wd.dtype)
    assert_maps_close(gd, wd, rtol=RTOL, atol=1e-12, what=gridname)
    np.testing.assert_array_equal(got.mask.include(), want._mask_include(), err_msg=gridname)
    np.testing.assert_allclose(got.spectral_axis, want.spectral_axis, rtol=1e-12)


@use_dask
@pytest.mark.parametrize('phase', [(0, 1), (3, 8), (7, 8), (1, 2)])
@pytest.mark.parametrize('kernel', ['direct', 'tma'])
@pytest.mark.parametrize('reverse_out', [False, True])
def test_scattered_interpolation_equals_the_plain_one(use_dask, reverse_out, kernel, phase, monkeypatch):
    """`sc_spectral_interp_scatter` (the form a row-sharded job uses to hand every output channel to its owner) with a
    pointer table that scatters the channels over two separate buffers, in a shuffled order: bit-identical with
    `spectral_interpolate` of the same cube -- also when the march over the spectrum starts in the middle and wraps
    around (`phase`: how the ranks of a job avoid storing to the same owner at the same time)."""
    import torch
    monkeypatch.setenv('SC_INTERP_KERNEL', '1' if kernel == 'direct' else '2')
    data = _random_cube((44, 6, 16), seed=43, nan_frac=0.08)
    sc = gpu_cube(data, BENCH_WCS, use_dask=use_dask, spectral_unit='km/s')
    sa = np.sort(sc.spectral_axis)
    grid = np.linspace(sa[0] - 2 * (sa[1] - sa[0]), sa[-1] + 2 * (sa[1] - sa[0]), 11)   # incl. samples left / right of the axis and on knots
    if reverse_out:
        grid = grid[::-1]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        want = sc.spectral_interpolate(grid)
        ny, nx = data.shape[1:]
        bufs = [torch.full((6, ny + 3, nx), -1.0, dtype=torch.float32, device='cuda') for _ in range(2)]   # rows 2 .. 2 + ny
        where = [(j % 2, [3, 0, 5, 1, 4, 2][j // 2]) for j in range(11)]   # (buffer, channel slot) of output channel j
        assert len(set(where)) == 11
        ptrs = torch.tensor([bufs[b].data_ptr() + ((slot * (ny + 3) + 2) * nx) * 4 for b, slot in where],
                            dtype=torch.int64, device='cuda')
        sc._spectral_interpolate_scatter(grid, ptrs, phase=phase[0], nphases=phase[1])
    ref = want._data
    for j, (b, slot) in enumerate(where):
        got = bufs[b][slot, 2:2 + ny]
        assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref[j], nan=-7.0)), j
        assert bool((bufs[b][slot, :2] == -1.0).all()) and bool((bufs[b][slot, 2 + ny:] == -1.0).all())   # nothing else touched


@use_dask
@pytest.mark.parametrize('reverse_out', [False, True])
def test_interpolate_benchmark_block_matches_oracle(use_dask, reverse_out, monkeypatch):
    """Config 5's interpolation (2048 -> 1024 channels on 4096-wide rows) on an 8-row block of the benchmark cube,
    through `spectral_interp_tma_kernel`, the kernel the benchmark runs (forced: on its own the library picks it from
    148 row tiles on, this block has 64), against np.interp / interp1d per spaxel (spectral_cube.py:3298-3315,
    dask_spectral_cube.py:1342-1364)."""
    import spectral_cube_b200 as scb
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs
    from oracle.synth import synth_block
    monkeypatch.setenv('SC_INTERP_KERNEL', '2')
    nchan, ny, nx, nout = 2048, 8, 4096, 1024
    w = benchmark_wcs(nchan, 4096, nx)
    wkw = dict(ctype=list(w.ctype), crval=[w.crval[0], w.crval[1], w.crval[2] / 1e3], crpix=list(w.crpix),
               cdelt=[w.cdelt[0], w.cdelt[1], w.cdelt[2] / 1e3], cunit=['deg', 'deg', 'km/s'])
    y0 = 98                                           # rows 98 .. 105 straddle the 102-row blank frame
    dev = synth_cube(nchan, ny, nx, y0=y0, ny_total=4096, nx_total=nx, nan_permille=1, border=102)
    host = synth_block(nchan, ny, nx, y0=y0, ny_total=4096, nx_total=nx, nan_permille=1, border=102)
    assert np.array_equal(dev.cpu().numpy().view(np.uint32), host.view(np.uint32))
    sc = gpu_cube(dev, wkw, use_dask=use_dask, spectral_unit='km/s')
    oc = oracle_cube(host, wkw, use_dask=use_dask, spectral_unit='km/s')
    sa = oc.spectral_axis
    grid = np.linspace(sa[0], sa[-1], nout)
    if reverse_out:
        grid = grid[::-1]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        got = sc.spectral_interpolate(grid, suppress_smooth_warning=True)
        want = oc.spectral_interpolate(grid, suppress_smooth_warning=True)
    assert_maps_close(data_of(got), want._data, rtol=RTOL, atol=1e-12, what='c5 block')
    np.testing.assert_array_equal(got.mask.include(), want._mask_include())


def test_interpolate_identity_and_linearity_at_scale():
    """512 x 256 x 2048 cube: resampling onto the input axis returns the input bits; resampling a ramp
    (linear in the spectral coordinate) onto any grid reproduces the ramp."""
    import torch
    import spectral_cube_b200 as scb
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs
    nchan, ny, nx = 512, 256, 2048
    w = benchmark_wcs(nchan, ny, nx)
    dev = synth_cube(nchan, ny, nx, nan_permille=1, border=3)
    cube = scb.SpectralCube(dev, w, unit='K')
    cube._mask = scb.LazyMask(np.isfinite, cube=cube)
    same = cube.spectral_interpolate(cube.spectral_axis)
    assert torch.equal(torch.nan_to_num(same._data, nan=-9.0), torch.nan_to_num(dev, nan=-9.0))
    sa = cube.spectral_axis
    ramp = torch.from_numpy(((sa - sa[0]) / (sa[-1] - sa[0])).astype(np.float32)).cuda()[:, None, None].expand(nchan, ny, nx).contiguous()
    rc = scb.SpectralCube(ramp, w, unit='K')
    grid = np.linspace(sa[3], sa[-7], 300)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out = rc.spectral_interpolate(grid)._data
    want = torch.from_numpy(((grid - sa[0]) / (sa[-1] - sa[0]))).cuda()[:, None, None]
    assert float((out.double() - want).abs().max()) < 2e-7


# ---- reproject -------------------------------------------------------------------------------------------
def rotated_header(wkw, shape, angle_deg, scale=1.0, shift=(0.0, 0.0)):
    """FITS-header-like dict: same projection and CRVAL, pixel grid rotated by `angle_deg`."""
    a = np.radians(angle_deg)
    cd = np.array(wkw['cdelt'][:2]) * scale
    pc = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
    hdr = {'NAXIS': 3, 'NAXIS1': shape[2], 'NAXIS2': shape[1], 'NAXIS3': shape[0],
           'CTYPE1': wkw['ctype'][0], 'CTYPE2': wkw['ctype'][1], 'CTYPE3': wkw['ctype'][2],
           'CRVAL1': wkw['crval'][0], 'CRVAL2': wkw['crval'][1], 'CRVAL3': wkw['crval'][2],
           'CRPIX1': shape[2] / 2.0 + 0.5 + shift[0], 'CRPIX2': shape[1] / 2.0 + 0.5 + shift[1], 'CRPIX3': wkw['crpix'][2],
           'CDELT1': cd[0], 'CDELT2': cd[1], 'CDELT3': wkw['cdelt'][2],
           'CUNIT1': 'deg', 'CUNIT2': 'deg', 'CUNIT3': wkw['cunit'][2],
           'PC1_1': pc[0, 0], 'PC1_2': pc[0, 1], 'PC2_1': pc[1, 0], 'PC2_2': pc[1, 1]}
    return hdr


def owcs_from_header(h):
    pc = np.eye(3)
    pc[0, 0], pc[0, 1], pc[1, 0], pc[1, 1] = h['PC1_1'], h['PC1_2'], h['PC2_1'], h['PC2_2']
    return OWCS(ctype=[h['CTYPE1'], h['CTYPE2'], h['CTYPE3']], crval=[h['CRVAL1'], h['CRVAL2'], h['CRVAL3']],
                crpix=[h['CRPIX1'], h['CRPIX2'], h['CRPIX3']], cdelt=[h['CDELT1'], h['CDELT2'], h['CDELT3']],
                cunit=[h['CUNIT1'], h['CUNIT2'], h['CUNIT3']], pc=pc)


def test_reproject_shape_and_wcs():
    """spectral_cube/tests/test_regrid.py:99-135: output shape and WCS follow the header."""
    cube = gpu_cube(G.adv_data(), G.ADV_WCS)
    hdr = rotated_header(G.ADV_WCS, (4, 5, 4), 0.0)
    hdr['CRPIX1'], hdr['CRPIX2'] = G.ADV_WCS['crpix'][0] + 1, G.ADV_WCS['crpix'][1] + 1
    out = cube.reproject(hdr)
    assert out.shape == (4, 5, 4)
    assert out.wcs.crpix[0] == hdr['CRPIX1'] and out.wcs.cdelt[0] == hdr['CDELT1']
    # shifted by exactly one pixel: interior values are the input values
    np.testing.assert_allclose(data_of(out)[:, 1:4, 1:3], G.adv_data().astype(np.float32), rtol=1e-9)   # the cube holds float32
    assert np.isnan(data_of(out)[:, 0, :]).all()
    assert not out.mask.include()[:, 0, :].any() and out.mask.include()[:, 1:4, 1:3].all()


@pytest.mark.parametrize('proj', ['TAN', 'SIN'])
@pytest.mark.parametrize('order', ['bilinear', 'nearest-neighbor'])
@pytest.mark.parametrize('angle,scale', [(0.0, 1.0), (30.0, 1.0), (-75.0, 0.7), (12.0, 1.9)])
def test_reproject_matches_oracle(angle, scale, order, proj):
    data = _random_cube((3, 40, 56), seed=int(abs(angle)) + 3, nan_frac=0.02)
    w = dict(BENCH_WCS)
    w['ctype'] = ['RA---' + proj, 'DEC--' + proj, 'VRAD']
    w['crpix'] = [28.5, 20.5, 1.0]
    sc, oc = gpu_cube(data, w), oracle_cube(data, w)
    hdr = rotated_header(w, (3, 48, 50), angle, scale, shift=(1.3, -2.1))
    got = sc.reproject(hdr, order=order)
    if order == 'bilinear':
        want = oc.reproject(owcs_from_header(hdr), (3, 48, 50))
        wd, wm = want._data, want._mask_include()
    else:
        # nearest neighbour: restate with scipy order 0 on the oracle's coordinates
        import scipy.ndimage
        from oracle.reproject import input_pixel_coords
        yin, xin = input_pixel_coords(oc._wcs, owcs_from_header(hdr), (48, 50))
        filled = oc.unitless_filled_data
        wd = np.stack([scipy.ndimage.map_coordinates(np.pad(filled[c].astype(float), 1, mode='edge'), [yin + 1, xin + 1],
                                                     order=0, mode='constant', cval=np.nan) for c in range(3)])
        outside = (yin < -0.5) | (yin > 39.5) | (xin < -0.5) | (xin > 55.5)
        wd[:, outside] = np.nan
        wm = ~np.isnan(wd)
    gd = data_of(got)
    assert gd.dtype == np.float64
    assert_maps_close(gd, wd, rtol=RTOL, atol=1e-9, what='%s %g %g' % (order, angle, scale))
    np.testing.assert_array_equal(got.mask.include(), wm)


def test_pixel_map_matches_oracle_wcs():
    """The device-side WCS chain agrees with the numpy restatement to ~1e-9 pixel over a 2048^2 grid
    far from the reference point (TAN and SIN, rotated)."""
    from oracle.reproject import input_pixel_coords
    for proj in ('TAN', 'SIN'):
        w = dict(BENCH_WCS)
        w['ctype'] = ['RA---' + proj, 'DEC--' + proj, 'VRAD']
        w['crpix'] = [1024.5, 1024.5, 1.0]
        sc = gpu_cube(np.zeros((1, 8, 8), dtype=np.float32), w)
        hdr = rotated_header(w, (1, 2048, 2048), 30.0)
        import spectral_cube_b200 as scb
        yin, xin = sc._pixel_map(scb.CubeWCS.from_header(hdr), 2048, 2048)
        oy, ox = input_pixel_coords(OWCS(**w), owcs_from_header(hdr), (2048, 2048))
        assert np.abs(yin.cpu().numpy() - oy).max() < 1e-8 and np.abs(xin.cpu().numpy() - ox).max() < 1e-8


def test_reproject_refuses_all_nan_result():
    cube = gpu_cube(G.adv_data(), G.ADV_WCS)
    hdr = rotated_header(G.ADV_WCS, (4, 5, 4), 0.0)
    hdr['CRVAL1'] += 5.0                                   # five degrees away: no overlap
    with pytest.raises(ValueError, match="All values in reprojected cube are nan"):
        cube.reproject(hdr)


def test_reproject_identity_and_integer_shift_at_scale():
    """Size-independent properties of the tiled kernel on a plane far larger than its tiles: reprojecting
    onto the cube's own header returns the data (the pixel map goes through the sphere, so the weights are
    1 - O(1e-12) and O(1e-12): 1e-6 relative), onto a header shifted by whole pixels returns the shifted
    data, and everything mapped from outside the image is NaN with a False footprint."""
    import torch
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs
    import spectral_cube_b200 as scb
    nchan, ny, nx = 6, 1024, 2048
    dev = synth_cube(nchan, ny, nx, border=0, nan_permille=0)
    w = benchmark_wcs(nchan, ny, nx)
    cube = scb.SpectralCube(dev, w, unit='K')
    same = cube.reproject(w, shape_out=(nchan, ny, nx))
    assert torch.allclose(same._data, dev, rtol=1e-6, atol=1e-6)
    assert bool(same._mask.include().all())
    w2 = w.copy()
    w2.crpix[0] += 7          # output pixel (x, y) looks at input pixel (x - 7, y + 5)
    w2.crpix[1] -= 5
    shifted = cube.reproject(w2, shape_out=(nchan, ny, nx))
    got = shifted._data
    assert torch.allclose(got[:, :ny - 5, 7:], dev[:, 5:, :nx - 7], rtol=1e-6, atol=1e-6)
    assert bool(torch.isnan(got[:, ny - 4:, :]).all()) and bool(torch.isnan(got[:, :, :6]).all())
    inc = shifted._mask.include()
    assert not inc[:, ny - 4:, :].any() and inc[:, :ny - 5, 7:].all()
