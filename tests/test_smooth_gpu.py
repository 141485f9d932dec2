"""
GPU parity tests for spectral_smooth / spatial_smooth and the fused spectral_smooth -> moment
chain.  Goldens: spectral_cube/tests/test_regrid.py:138-172 and
spectral_cube/tests/test_spectral_cube.py:2363-2421; everything else against the CPU oracle
(oracle/convolve.py restates astropy.convolution.convolve).  Tolerance 1e-5 relative.
"""
import warnings

import numpy as np
import pytest

import oracle.convolve as oconv
from tests.golden import reference_goldens as G
from tests.helpers import oracle_cube, gpu_cube, assert_maps_close, RTOL
from tests.test_moments_gpu import BENCH_WCS, _random_cube, quiet

pytestmark = pytest.mark.gpu


def _kernels():
    import spectral_cube_b200 as scb
    return scb


def delta_pair(data, use_dask, flip=False):
    w = dict(G.ADV_WCS)
    if flip:
        w['cdelt'] = [w['cdelt'][0], w['cdelt'][1], -w['cdelt'][2]]
    return (gpu_cube(data, w, use_dask=use_dask, spectral_unit='km/s'),
            oracle_cube(data, w, use_dask=use_dask, spectral_unit='km/s'))


# ---- spectral_cube/tests/test_regrid.py:138-172 ----------------------------------------------------
@pytest.mark.parametrize('use_dask', [False, True])
def test_spectral_smooth(use_dask):
    scb = _kernels()
    cube, _ = delta_pair(G.delta_522(), use_dask)
    kernel = scb.Gaussian1DKernel(1.0)
    result = cube.spectral_smooth(kernel)
    assert kernel.array.size == 9
    np.testing.assert_almost_equal(result.unmasked_data[:, 0, 0], kernel.array[2:-2], 4)
    # dtype contract: numpy class -> float64 cube (:2953/:2963), dask class keeps float32 (:829)
    assert result.unmasked_data[:].dtype == (np.float32 if use_dask else np.float64)
    assert result.mask is cube.mask                      # mask object untouched (:3043-3045)


def test_spectral_smooth_rejects_even_kernels_and_units():
    scb = _kernels()
    cube, _ = delta_pair(G.delta_522(), False)
    with pytest.raises(Exception, match="odd"):
        cube.spectral_smooth(np.ones(4))

    class WithUnit(object):
        class array(object):
            unit = 'K'
    with pytest.raises(Exception, match="without a unit"):
        cube.spectral_smooth(WithUnit())


# ---- oracle parity ----------------------------------------------------------------------------------
@pytest.mark.parametrize('use_dask', [False, True])
@pytest.mark.parametrize('ntaps', [1, 3, 5, 9, 13, 17, 21, 25, 33, 41])
def test_spectral_smooth_matches_oracle(ntaps, use_dask):
    scb = _kernels()
    rng = np.random.default_rng(ntaps)
    data = _random_cube((50, 6, 16), seed=100 + ntaps, nan_frac=0.04)
    data[10:40, 2, 3] = np.nan                            # a long NaN run: bot == 0 inside for short kernels
    taps = rng.random(ntaps) + 0.05                      # asymmetric on purpose: catches a missing flip
    sc, oc = gpu_cube(data, BENCH_WCS, use_dask=use_dask), oracle_cube(data, BENCH_WCS, use_dask=use_dask)
    got = sc.spectral_smooth(scb.CustomKernel(taps)).unmasked_data[:]
    want = oc.spectral_smooth(oconv.Kernel(taps))._data
    assert got.dtype == want.dtype
    assert_maps_close(got, want, rtol=RTOL, what='ntaps=%d' % ntaps)


@pytest.mark.parametrize('use_dask', [False, True])
def test_spectral_smooth_gaussian_fwhm5_under_mask(use_dask):
    """BASELINE config 3's kernel (17 taps) on a cube carrying a > threshold mask: masked voxels are
    NaN-filled before convolving and interpolated over."""
    scb = _kernels()
    data = _random_cube((96, 5, 132), seed=7, nan_frac=0.01)
    sc, oc = gpu_cube(data, BENCH_WCS, use_dask=use_dask), oracle_cube(data, BENCH_WCS, use_dask=use_dask)
    sc, oc = sc.with_mask(sc > -0.5), oc.with_mask(oc > -0.5)
    k = 5 / 2.3548200450309493
    got = sc.spectral_smooth(scb.Gaussian1DKernel(k)).unmasked_data[:]
    want = oc.spectral_smooth(oconv.Gaussian1DKernel(k))._data
    assert_maps_close(got, want, rtol=RTOL, what='fwhm5')


def test_spectral_smooth_views_and_odd_widths():
    import torch
    scb = _kernels()
    base = _random_cube((40, 9, 50), seed=9)
    for sl in [(slice(None), slice(None), slice(0, 48)), (slice(3, 37), slice(1, 8), slice(2, 45)),
               (slice(None), slice(None), slice(None))]:
        dev = torch.from_numpy(base).cuda()[sl]
        sc, oc = gpu_cube(dev, BENCH_WCS, use_dask=True), oracle_cube(base[sl], BENCH_WCS, use_dask=True)
        got = sc.spectral_smooth(scb.Gaussian1DKernel(1.5)).unmasked_data[:]
        want = oc.spectral_smooth(oconv.Gaussian1DKernel(1.5))._data
        assert_maps_close(got, want, rtol=RTOL, what=str(sl))


@pytest.mark.parametrize('kernel_path', ['tma', 'generic'])
def test_smooth_then_moment_fused_matches_oracle(kernel_path, monkeypatch):
    """config 3: spectral_smooth(Gaussian FWHM 5 ch) then moment0/1/2 on the dask class; the GPU path
    never writes the smoothed cube."""
    scb = _kernels()
    monkeypatch.setenv('SC_SMOOTH_KERNEL', '2' if kernel_path == 'tma' else '1')
    data = _random_cube((128, 6, 64), seed=17, nan_frac=0.01) + np.float32(4.0)     # positive weights
    sc, oc = gpu_cube(data, BENCH_WCS, use_dask=True), oracle_cube(data, BENCH_WCS, use_dask=True)
    k = 5 / 2.3548200450309493
    ssm, osm = sc.spectral_smooth(scb.Gaussian1DKernel(k)), oc.spectral_smooth(oconv.Gaussian1DKernel(k))
    assert ssm._data_t is None                           # still lazy
    for order in (0, 1, 2):
        got = quiet(ssm.moment, order=order).value
        want = quiet(osm.moment, order=order)[0]
        assert_maps_close(got, want, rtol=RTOL, what='fused moment%d' % order)
    assert ssm._data_t is None                           # the smoothed cube was never materialised
    # and the materialised route gives the same maps
    mat = sc.spectral_smooth(scb.Gaussian1DKernel(k), save_to_tmp_dir=True)
    assert mat._data_t is not None
    for order in (0, 1, 2):
        assert_maps_close(quiet(mat.moment, order=order).value, quiet(ssm.moment, order=order).value,
                          rtol=1e-9, atol=1e-9, what='materialised vs fused %d' % order)


@pytest.mark.parametrize('use_dask', [False, True])
@pytest.mark.parametrize('shape', [(64, 6, 64), (40, 5, 13)])
def test_moments_of_a_smoothed_cube_under_the_mask_of_its_source(shape, use_dask):
    """The mask object survives smoothing untouched (spectral_cube.py:3043-3045): a `> threshold` LazyMask keeps
    testing the ORIGINAL data while the moments sum the smoothed data.  The materialised route streams both
    cubes (interval test on the other cube); odd widths take the scalar loads."""
    scb = _kernels()
    data = _random_cube(shape, seed=shape[2], nan_frac=0.02) + np.float32(2.0)
    sc, oc = gpu_cube(data, BENCH_WCS, use_dask=use_dask), oracle_cube(data, BENCH_WCS, use_dask=use_dask)
    sc, oc = sc.with_mask(sc > 3.0), oc.with_mask(oc > 3.0)
    k = 5 / 2.3548200450309493
    kw = dict(save_to_tmp_dir=True) if use_dask else {}
    ssm, osm = sc.spectral_smooth(scb.Gaussian1DKernel(k), **kw), oc.spectral_smooth(oconv.Gaussian1DKernel(k))
    for order in (0, 1, 2):
        got = quiet(ssm.moment, order=order).value
        want = quiet(osm.moment, order=order)[0]
        assert_maps_close(got, want, rtol=RTOL, atol=1e-6 if order == 2 else 0.0, what='moment%d' % order)


def test_smoothing_is_linear_and_preserves_constants_at_scale():
    """Size-independent properties on a 512 x 256 x 2048 cube (1 GB): a constant cube stays constant
    away from the spectral edges; smoothing 2x the data gives 2x the result exactly."""
    import torch
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs
    scb = _kernels()
    nchan, ny, nx = 512, 256, 2048
    w = benchmark_wcs(nchan, ny, nx)
    k = scb.Gaussian1DKernel(5 / 2.3548200450309493)
    const = scb.DaskSpectralCube(torch.full((nchan, ny, nx), 3.25, device='cuda'), w, unit='K')
    out = const.spectral_smooth(k)._data
    assert torch.allclose(out[8:-8], torch.full_like(out[8:-8], 3.25), rtol=1e-6, atol=0)
    assert float(out[0].max()) < 3.25                    # zero-filled boundary tapers the edge channels
    dev = synth_cube(nchan, ny, nx, nan_permille=1, border=4)
    a = scb.DaskSpectralCube(dev, w, unit='K').spectral_smooth(k)._data
    b = scb.DaskSpectralCube(dev * 2.0, w, unit='K').spectral_smooth(k)._data
    assert torch.equal(torch.nan_to_num(b, nan=-1.0), torch.nan_to_num(a * 2.0, nan=-1.0))
    # NaN-interpolation: isolated NaN voxels are filled in; an all-NaN spectrum stays NaN except in the
    # first/last 8 channels, where the window reaches the (valid) zero padding and the result is 0
    assert not bool(torch.isnan(a[:, 10, 10:-10]).any())
    assert bool(torch.isnan(a[8:-8, 0, :]).all()) and bool((a[:8, 0, :] == 0).all())


# ---- spatial smoothing: spectral_cube/tests/test_spectral_cube.py:2363-2421 ---------------------------
@pytest.mark.parametrize('use_dask', [False, True])
def test_spatial_smooth_g2d(use_dask):
    scb = _kernels()
    cube = gpu_cube(G.adv_data(), G.ADV_WCS, use_dask=use_dask)
    res = cube.spatial_smooth(scb.Gaussian2DKernel(3)).unmasked_data[:]
    np.testing.assert_almost_equal(res[0], G.G2D_RESULT0)
    np.testing.assert_almost_equal(res[2], G.G2D_RESULT2)


@pytest.mark.parametrize('use_dask', [False, True])
def test_spatial_smooth_t2d(use_dask):
    scb = _kernels()
    cube = gpu_cube(G.adv_data(), G.ADV_WCS, use_dask=use_dask)
    res = cube.spatial_smooth(scb.Tophat2DKernel(3)).unmasked_data[:]
    np.testing.assert_almost_equal(res[0], G.T2D_RESULT0)
    np.testing.assert_almost_equal(res[2], G.T2D_RESULT2)


def test_spatial_smooth_preserves_unit_and_refuses_jybeam():
    scb = _kernels()
    cube = gpu_cube(G.adv_data(), G.ADV_WCS, unit='Jy/beam')
    with pytest.raises(scb.BeamUnitsError, match="Jy/beam"):
        cube.spatial_smooth(scb.Gaussian2DKernel(3))
    out = cube.spatial_smooth(scb.Gaussian2DKernel(3), raise_error_jybm=False)
    assert out.unit == 'Jy/beam'


@pytest.mark.parametrize('use_dask', [False, True])
@pytest.mark.parametrize('sigma', [0.5, 1.0, 8 / 2.3548200450309493, 4.0])
def test_spatial_smooth_gaussian_matches_oracle(sigma, use_dask):
    """Separable (marching) path incl. config 4's FWHM 8 px kernel (29 x 29)."""
    scb = _kernels()
    data = _random_cube((3, 70, 300), seed=int(sigma * 10), nan_frac=0.01)
    data[1, 20:60, 100:160] = np.nan                     # a hole wider than the kernel: bot == 0 inside
    sc, oc = gpu_cube(data, BENCH_WCS, use_dask=use_dask), oracle_cube(data, BENCH_WCS, use_dask=use_dask)
    got = sc.spatial_smooth(scb.Gaussian2DKernel(sigma)).unmasked_data[:]
    want = oc.spatial_smooth(oconv.Gaussian2DKernel(sigma))._data
    assert got.dtype == want.dtype
    assert_maps_close(got, want, rtol=RTOL, what='sigma=%g' % sigma)


# the separable kernels the library can be told to run (SC_SPATIAL_KERNEL / SC_SPATIAL_J): "auto" picks the pipelined kernel
# or, with more than a third of the 8 x 128 blocks crowded with missing samples, the convolved-denominator march -- forcing each one keeps every denominator path under test
SEPARABLE_KERNELS = {'auto': (None, None), 'pipe_j8': ('5', '8'), 'pipe_j16': ('5', '16'), 'sparse_r1': ('4', None), 'march': ('3', None)}


def _force_kernel(monkeypatch, name):
    kern, j = SEPARABLE_KERNELS[name]
    if kern is not None:
        monkeypatch.setenv('SC_SPATIAL_KERNEL', kern)
    if j is not None:
        monkeypatch.setenv('SC_SPATIAL_J', j)


@pytest.mark.parametrize('kernel', sorted(SEPARABLE_KERNELS))
@pytest.mark.parametrize('sigma', [1.0, 8 / 2.3548200450309493])
def test_spatial_smooth_blank_bands_and_crowded_blocks(sigma, kernel, monkeypatch):
    """The denominator paths of the separable kernels in one image: isolated NaNs (listed inputs -> scattered row
    deficits), a half-masked region (crowded blocks: deficits from the runs of missing samples / integer convolution of
    the flags) and a blank band wider than a CTA's window (closed form; outputs deep inside keep the filled input)."""
    _force_kernel(monkeypatch, kernel)
    scb = _kernels()
    rng = np.random.default_rng(int(sigma * 100))
    data = _random_cube((2, 96, 640), seed=int(sigma * 7), nan_frac=0.002)
    data[:, 24:72, 90:520] = np.nan                                           # blank band
    speckle = rng.random((2, 20, 300)) < 0.5
    data[:, 74:94, 200:500][speckle] = np.nan                                 # crowded, not blank
    sc, oc = gpu_cube(data, BENCH_WCS, use_dask=True), oracle_cube(data, BENCH_WCS, use_dask=True)
    got = sc.spatial_smooth(scb.Gaussian2DKernel(sigma)).unmasked_data[:]
    want = oc.spatial_smooth(oconv.Gaussian2DKernel(sigma))._data
    assert np.isnan(want[:, 45:50, 200:400]).all()                            # deep inside the band: nothing valid
    assert_maps_close(got, want, rtol=RTOL, what='sigma=%g' % sigma)


@pytest.mark.parametrize('kernel', ['pipe_j8', 'pipe_j16', 'march'])
@pytest.mark.parametrize('use_dask', [False, True])
def test_spatial_smooth_benchmark_block_matches_oracle(kernel, use_dask, monkeypatch):
    """Config 4's smoothing (29 x 29 Gaussian, FWHM 8 px) on a block of the benchmark cube that holds everything the
    full cube does: 4096-pixel rows (32 strips, the outer two clipped by the image and holding the blank frame's 102
    side columns), the frame's last rows (blank blocks, then a crowded one), 0.1 % NaNs -- against the astropy
    restatement (spectral_cube.py:2808-2842, dask :962-993)."""
    _force_kernel(monkeypatch, kernel)
    scb = _kernels()
    from spectral_cube_b200.synth import synth_cube
    from oracle.synth import synth_block
    ny, nx, y0 = 72, 4096, 60                              # rows 60 .. 131 of the 4096-row image: 42 blank rows, then data
    dev = synth_cube(1, ny, nx, y0=y0, ny_total=4096, nx_total=nx, nan_permille=1, border=102)
    host = synth_block(1, ny, nx, y0=y0, ny_total=4096, nx_total=nx, nan_permille=1, border=102)
    assert np.array_equal(dev.cpu().numpy().view(np.uint32), host.view(np.uint32))
    sc, oc = gpu_cube(dev, BENCH_WCS, use_dask=use_dask), oracle_cube(host, BENCH_WCS, use_dask=use_dask)
    got = sc.spatial_smooth(scb.Gaussian2DKernel(8 / 2.3548200450309493)).unmasked_data[:]
    want = oc.spatial_smooth(oconv.Gaussian2DKernel(8 / 2.3548200450309493))._data
    # (local rows < 14 see the zero padding above the block -- valid samples -- and come out as zeros)
    assert np.isnan(want[0, 15:25, 200:3800]).all() and np.isfinite(want[0, 60:, 200:3800]).all()
    assert_maps_close(got, want, rtol=RTOL, what='config-4 block, %s' % kernel)


def test_spatial_smooth_fixup_list_full_scans_every_tile(monkeypatch):
    """The pipe kernel lists the tiles that hold outputs to be redone exactly; when the list is full the fix-up kernel
    scans the whole output instead.  A two-entry list forces that path: same result as with the full-size list."""
    import torch
    _force_kernel(monkeypatch, 'pipe_j8')
    scb = _kernels()
    from spectral_cube_b200.synth import synth_cube
    dev = synth_cube(2, 72, 4096, y0=60, ny_total=4096, nx_total=4096, nan_permille=1, border=102)
    k = scb.Gaussian2DKernel(8 / 2.3548200450309493)
    ref = gpu_cube(dev, BENCH_WCS, use_dask=True).spatial_smooth(k)._data
    monkeypatch.setenv('SC_SPATIAL_FIX_CAP', '2')
    got = gpu_cube(dev, BENCH_WCS, use_dask=True).spatial_smooth(k)._data
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref, nan=-7.0))
    assert not bool((got.view(torch.int32) == 0x7FC5CB20).any())          # no marker left behind


def test_spatial_smooth_row_shards_across_a_blank_frame_equal_the_whole_image(monkeypatch):
    """The exact treatment of blank-region edges (re-fold in the column warps, fix-up kernel) with halo rows: two row
    shards of a block of the benchmark cube, cut INSIDE the rows where the blank frame ends, against the unsharded
    result -- bit for bit -- and against the oracle."""
    import torch
    from spectral_cube_b200 import _lib
    _force_kernel(monkeypatch, 'pipe_j8')
    scb = _kernels()
    lib = _lib.load()
    from spectral_cube_b200.synth import synth_cube
    from oracle.synth import synth_block
    ny, nx, y0, cut = 96, 4096, 60, 40                       # frame ends at local row 42: its edge outputs straddle the cut
    dev = synth_cube(1, ny, nx, y0=y0, ny_total=4096, nx_total=nx, nan_permille=1, border=102)
    k = scb.Gaussian2DKernel(8 / 2.3548200450309493)
    h = k.shape[0] // 2
    whole = gpu_cube(dev, BENCH_WCS, use_dask=True)
    ref = whole.spatial_smooth(k)._data
    top, bot = gpu_cube(dev[:, :cut].contiguous(), BENCH_WCS, use_dask=True), gpu_cube(dev[:, cut:].contiguous(), BENCH_WCS, use_dask=True)

    def pack(cube, row0, nrows):
        out = torch.empty((cube.shape[0], nrows, cube.shape[2]), dtype=torch.float32, device='cuda')
        desc, keep = cube._mask_desc()
        d = cube._data
        _lib.check(lib.sc_pack_filled_rows(d.data_ptr(), d.shape[0], d.shape[1], d.shape[2], d.stride(0), d.stride(1),
                                           desc, float('nan'), row0, nrows, out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
        return out
    halo_for_top, halo_for_bot = pack(bot, 0, h), pack(top, cut - h, h)
    counts = top._spatial_strategy_counts() + bot._spatial_strategy_counts()
    a = top._run_spatial_smooth(k.array, _lib.F32, halo_top=None, halo_bot=halo_for_top, halo_rows=h, strategy_counts=counts)
    b = bot._run_spatial_smooth(k.array, _lib.F32, halo_top=halo_for_bot, halo_bot=None, halo_rows=h, strategy_counts=counts)
    got = torch.cat([a, b], dim=1)
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref, nan=-7.0))
    host = synth_block(1, ny, nx, y0=y0, ny_total=4096, nx_total=nx, nan_permille=1, border=102)
    want = oracle_cube(host, BENCH_WCS, use_dask=True).spatial_smooth(oconv.Gaussian2DKernel(8 / 2.3548200450309493))._data
    assert_maps_close(got.cpu().numpy(), want, rtol=RTOL, what='two shards across the frame edge')


def test_spatial_smooth_elliptical_and_nonseparable_match_oracle():
    scb = _kernels()
    data = _random_cube((2, 40, 64), seed=77, nan_frac=0.02)
    sc, oc = gpu_cube(data, BENCH_WCS, use_dask=True), oracle_cube(data, BENCH_WCS, use_dask=True)
    for kg, ko in [(scb.Gaussian2DKernel(1.0, 2.0), oconv.Gaussian2DKernel(1.0, 2.0)),            # separable, hy != hx
                   (scb.Gaussian2DKernel(1.0, 2.0, theta=0.5), oconv.Gaussian2DKernel(1.0, 2.0, theta=0.5)),
                   (scb.Tophat2DKernel(3), oconv.Tophat2DKernel(3))]:
        got = sc.spatial_smooth(kg).unmasked_data[:]
        want = oc.spatial_smooth(ko)._data
        assert_maps_close(got, want, rtol=RTOL, what=str(kg.shape))


def test_spatial_smooth_under_mask_and_fill_value():
    scb = _kernels()
    data = _random_cube((2, 33, 140), seed=5, nan_frac=0.0)
    for fill in (np.nan, 0.0):
        sc, oc = gpu_cube(data, BENCH_WCS, use_dask=True), oracle_cube(data, BENCH_WCS, use_dask=True)
        sc, oc = sc.with_mask(sc > 0.0).with_fill_value(fill), oc.with_mask(oc > 0.0).with_fill_value(fill)
        got = sc.spatial_smooth(scb.Gaussian2DKernel(1.5)).unmasked_data[:]
        want = oc.spatial_smooth(oconv.Gaussian2DKernel(1.5))._data
        assert_maps_close(got, want, rtol=RTOL, atol=1e-7, what='fill=%r' % fill)


def test_spatial_smooth_row_shards_with_halos_equal_the_whole_image():
    """Row sharding: smoothing two row blocks with each other's filled edge rows as
    halos reproduces the unsharded result bit for bit."""
    import torch
    from spectral_cube_b200 import _lib
    scb = _kernels()
    lib = _lib.load()
    data = _random_cube((4, 64, 256), seed=31, nan_frac=0.01)
    k = scb.Gaussian2DKernel(8 / 2.3548200450309493)
    whole = gpu_cube(data, BENCH_WCS, use_dask=True)
    ref = whole.spatial_smooth(k)._data
    h = k.shape[0] // 2
    top, bot = gpu_cube(data[:, :40], BENCH_WCS, use_dask=True), gpu_cube(data[:, 40:], BENCH_WCS, use_dask=True)

    def pack(cube, row0, nrows):
        out = torch.empty((cube.shape[0], nrows, cube.shape[2]), dtype=torch.float32, device='cuda')
        desc, keep = cube._mask_desc()
        d = cube._data
        _lib.check(lib.sc_pack_filled_rows(d.data_ptr(), d.shape[0], d.shape[1], d.shape[2], d.stride(0), d.stride(1),
                                           desc, float('nan'), row0, nrows, out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
        return out
    halo_for_top = pack(bot, 0, h)                       # the rows just below the top shard
    halo_for_bot = pack(top, 40 - h, h)                  # the rows just above the bottom shard
    # one denominator strategy for all shards: the job-wide sample (here: the shards' counts summed)
    counts = top._spatial_strategy_counts() + bot._spatial_strategy_counts()
    assert torch.equal(counts, whole._spatial_strategy_counts())          # 40 rows: the shards' 8-row sample blocks are the whole cube's
    a = top._run_spatial_smooth(k.array, _lib.F32, halo_top=None, halo_bot=halo_for_top, halo_rows=h, strategy_counts=counts)
    b = bot._run_spatial_smooth(k.array, _lib.F32, halo_top=halo_for_bot, halo_bot=None, halo_rows=h, strategy_counts=counts)
    got = torch.cat([a, b], dim=1)
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref, nan=-7.0))
