"""
GPU parity tests for `convolve_to`: SpectralCube / DaskSpectralCube with one beam and
VaryingResolutionSpectralCube with per-channel beams, through the C ABI (sc_spatial_smooth_sep_ex /
sc_spatial_smooth_2d + sc_scale) against the oracle (oracle/cube.py `convolve_to`, oracle/beam.py).  Goldens:
spectral_cube/tests/test_regrid.py:33-96, tests/test_spectral_cube.py:2150-2225 with conftest.py:590-660.
Tolerance 1e-5 relative (float32); where the oracle's default is `convolve_fft` an absolute floor of 1e-7 of the
data scale covers the transform's rounding noise, and outputs whose interpolation weight is below 1e-6 -- where
`convolve_fft` returns that noise divided by the weight -- are compared against the direct convolution instead.

First run on a B200: 35 passed (profiles/r01_convolve_to_gpu_tests.log), after a CPU dry run of these tests against a
numpy emulation of the entry points (tools/dryrun/emu_plugin.py) had corrected three expectations.
"""
import warnings

import numpy as np
import pytest

import oracle.convolve as oconv
from oracle.beam import OBeam, SIGMA_TO_FWHM
from tests.golden import reference_goldens as G
from tests.helpers import oracle_cube, gpu_cube, assert_maps_close, RTOL
from tests.test_moments_gpu import _random_cube

pytestmark = pytest.mark.gpu
use_dask = pytest.mark.parametrize('use_dask', [False, True])

BEAMS4 = dict(major=[0.4, 0.3, 0.3, 0.4], minor=[0.1, 0.2, 0.2, 0.1], pa=[0, 45, 60, 30])        # conftest.py:61-79
BEAMS5_PIX = dict(major=[3.5, 3, 3, 3, 3], minor=[2, 2.5, 3, 2.5, 2], pa=[0, 45, 60, 30, 0])   # conftest.py:565-577


def scb():
    import spectral_cube_b200
    return spectral_cube_b200


def pair(data, use_dask, beam=None, unit='K', **kw):
    """(device cube, oracle cube) with one beam given in arcsec (major, minor, pa)."""
    pb = scb().Beam.from_arcsec(*beam) if beam is not None else None
    ob = OBeam.arcsec(*beam) if beam is not None else None
    return (gpu_cube(data, G.ADV_WCS, unit=unit, use_dask=use_dask, beam=pb, **kw),
            oracle_cube(data, G.ADV_WCS, unit=unit, use_dask=use_dask, beam=ob, **kw))


def vr_pair(data, use_dask, beams, unit='K'):
    """(device VaryingResolutionSpectralCube, oracle cube) with per-channel beams in arcsec."""
    S = scb()
    from spectral_cube_b200.masks import LazyMask
    pb = S.Beams.from_arcsec(beams['major'], beams['minor'], beams['pa'])
    ob = [OBeam.arcsec(a, b, p) for a, b, p in zip(beams['major'], beams['minor'], beams['pa'])]
    cube = S.VaryingResolutionSpectralCube(np.asarray(data, dtype=np.float32), S.CubeWCS(**G.ADV_WCS), unit=unit,
                                           beams=pb, use_dask=use_dask)
    cube._mask = LazyMask(np.isfinite, cube=cube)
    return cube, oracle_cube(data, G.ADV_WCS, unit=unit, use_dask=use_dask, beams=ob)


def assert_close_where_input_valid(got, want, data, what):
    """Parity on every voxel whose own input is valid (interpolation weight >= the kernel's centre tap).  At blank
    inputs the two astropy functions already disagree with each other: where the weight left by the neighbours is
    below 10 eps (sub-pixel kernels: e^-50 one pixel away) `convolve_fft` returns 0.0 -- or rounding noise over the
    weight a little above that -- while `convolve` and the device return the exact weighted mean.  Those voxels are
    excluded by the cube's mask either way."""
    ok = np.isfinite(data)
    assert got.shape == want.shape and ok.any()
    assert_maps_close(np.where(ok, got, 0.0), np.where(ok, want, 0.0), rtol=RTOL, atol=1e-6, what=what)


def point_source(beams):
    """conftest.py:590-660"""
    d = np.zeros((5, 11, 11))
    d[:, 5, 5] = 1.
    for i, bm in enumerate(beams):
        d[i] = oconv.convolve_fft(d[i], bm.as_kernel(2. / 3600.)) * bm.sr / np.deg2rad(2. / 3600.) ** 2
    return d


# ---- reference goldens ---------------------------------------------------------------------------------
@use_dask
def test_convolution(use_dask):
    """tests/test_regrid.py:33-58"""
    d = np.zeros([2, 5, 5])
    d[0, 2, 2] = 1.0
    cube, _ = pair(d, use_dask, beam=(1.0, 1.0, 0.0))
    target = scb().Beam.from_arcsec(1.802775637731995, 1.802775637731995, 0.0)
    conv = cube.convolve_to(target)
    expected = oconv.Gaussian2DKernel(1.5 / 3600. / SIGMA_TO_FWHM / 5.555555555555e-4, x_size=5, y_size=5).array
    np.testing.assert_almost_equal(expected / expected.sum(), conv.filled_data[0, :, :])
    assert np.all(conv.filled_data[1, :, :] == 0.0)
    assert conv.beam == target and type(conv) is type(cube)
    assert conv.header['BMAJ'] == target.major and conv.meta['beam'] == target


@use_dask
def test_beams_convolution(use_dask):
    """tests/test_regrid.py:61-82"""
    d = np.zeros([4, 5, 5])
    d[:, 2, 2] = 1.0
    cube, oc = vr_pair(d, use_dask, BEAMS4)
    S = scb()
    target = S.Beam.from_arcsec(1.802775637731995, 1.802775637731995, 0.0)
    conv = cube.convolve_to(target)
    assert type(conv) is (S.DaskSpectralCube if use_dask else S.SpectralCube) and conv.beam == target
    for ii, bm in enumerate(cube.beams):
        expected = target.deconvolve(bm).as_kernel(cube._pixscale_deg(), x_size=5, y_size=5).array
        np.testing.assert_almost_equal(expected / expected.sum(), conv.filled_data[ii, :, :])


@use_dask
def test_beams_convolution_equal(use_dask):
    """tests/test_regrid.py:85-101"""
    d = np.zeros([5, 2, 2])
    d[2] = 1.0
    beams = dict(major=[1.0, 0.4, 0.3, 0.4, 0.5], minor=[1.0, 0.2, 0.3, 0.2, 0.1], pa=[0, 45, 60, 30, 0])
    cube, _ = vr_pair(d, use_dask, beams)
    conv = cube.convolve_to(scb().Beam.from_arcsec(1.0, 1.0, 0.0))
    np.testing.assert_almost_equal(cube.filled_data[0], conv.filled_data[0])


@use_dask
def test_convolve_to_equal(use_dask):
    """tests/test_spectral_cube.py:2150-2157"""
    data = _random_cube((3, 12, 16), seed=3, nan_frac=0.0)
    cube, _ = pair(data, use_dask, beam=(1.0, 1.0, 0.0))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        convolved = cube.convolve_to(cube.beam)
    assert convolved is cube and any("identical to the current beam" in str(x.message) for x in w)
    no_beam = gpu_cube(data, G.ADV_WCS, use_dask=use_dask)
    with pytest.raises(scb().NoBeamError):
        no_beam.convolve_to(scb().Beam.from_arcsec(2.0))
    with pytest.raises(NotImplementedError):
        cube.convolve_to(scb().Beam.from_arcsec(2.0), nan_treatment='fill')


@use_dask
def test_convolve_to_jybeam_onebeam(use_dask):
    """tests/test_spectral_cube.py:2181-2189: the peak of the point source stays constant in Jy/beam."""
    d = point_source([OBeam.arcsec(6.0)] * 5)
    cube, oc = pair(d, use_dask, beam=(6.0, 6.0, 0.0), unit='Jy/beam')
    convolved = cube.convolve_to(scb().Beam.from_arcsec(10.0))
    np.testing.assert_allclose(convolved.unmasked_data[:][:, 5, 5], d[:, 5, 5], atol=1e-5, rtol=1e-5)
    assert convolved.unit == 'Jy/beam'
    want = oc.convolve_to(OBeam.arcsec(10.0))._data
    assert_maps_close(convolved.unmasked_data[:], want, rtol=RTOL, atol=1e-7, what='one beam, Jy/beam')


@use_dask
def test_convolve_to_jybeam_multibeams(use_dask):
    """tests/test_spectral_cube.py:2192-2201"""
    beams = dict(major=[2 * a for a in BEAMS5_PIX['major']], minor=[2 * b for b in BEAMS5_PIX['minor']], pa=BEAMS5_PIX['pa'])
    d = point_source([OBeam.arcsec(a, b, p) for a, b, p in zip(beams['major'], beams['minor'], beams['pa'])])
    cube, oc = vr_pair(d, use_dask, beams, unit='Jy/beam')
    convolved = cube.convolve_to(scb().Beam.from_arcsec(10.0))
    np.testing.assert_allclose(convolved.unmasked_data[:][:, 5, 5], d[:, 5, 5], atol=1e-5, rtol=1e-5)
    want = oc.convolve_to(OBeam.arcsec(10.0))._data
    assert_maps_close(convolved.unmasked_data[:], want, rtol=RTOL, atol=1e-7, what='five beams, Jy/beam')


@use_dask
def test_convolve_to_with_bad_beams(use_dask):
    """tests/test_spectral_cube.py:2204-2225"""
    data = _random_cube((4, 20, 24), seed=9, nan_frac=0.0)
    data[:, 0, :] = data[:, 1, :]                               # no blank row here: the reference's cube has none
    data[0, 7, 9] = data[2, 11, 4] = np.nan
    S = scb()
    cube, oc = vr_pair(data, use_dask, BEAMS4)
    got = cube.convolve_to(S.Beam.from_arcsec(0.5)).unmasked_data[:]
    assert_close_where_input_valid(got, oc.convolve_to(OBeam.arcsec(0.5))._data, data, '0.5 arcsec')
    if use_dask:             # the dask class's `convolve` is what the device computes, blank inputs included
        assert_maps_close(got, oc.convolve_to(OBeam.arcsec(0.5))._data, rtol=RTOL, atol=1e-6, what='0.5 arcsec, all voxels')
    with pytest.raises(ValueError, match="Beam could not be deconvolved"):
        cube.convolve_to(S.Beam.from_arcsec(0.35))              # the biggest beam is 0.4 arcsec
    masked, omasked = cube.mask_channels([False, True, True, False]), oc.mask_channels([False, True, True, False])
    assert list(masked.goodbeams_mask) == [False, True, True, False] and len(masked.beams) == 2
    convolved = masked.convolve_to(S.Beam.from_arcsec(0.35))
    assert np.all(np.isfinite(convolved.filled_data[1:3][np.isfinite(data[1:3])]))
    want = omasked.convolve_to(OBeam.arcsec(0.35))
    assert_close_where_input_valid(convolved.unmasked_data[:], want._data, data, '0.35 arcsec, two channels')
    assert np.all(np.isnan(convolved.unmasked_data[:][[0, 3]])) and np.all(np.isnan(want._data[[0, 3]]))   # filled copies
    # allow_smaller: channels that cannot be deconvolved are copied through (filled)
    got = cube.convolve_to(S.Beam.from_arcsec(0.35), allow_smaller=True).unmasked_data[:]
    want = oc.convolve_to(OBeam.arcsec(0.35), allow_smaller=True)._data
    assert_close_where_input_valid(got, want, data, 'allow_smaller')
    assert np.array_equal(np.isnan(got[[0, 3]]), np.isnan(data[[0, 3]]))          # channels copied as they are
    with pytest.raises(AttributeError):
        cube.spectral_smooth(S.Gaussian1DKernel(1.0))


# ---- against the oracle on random cubes --------------------------------------------------------------------
@use_dask
@pytest.mark.parametrize('beam,target,shape', [
    ((3.0, 3.0, 0.0), (5.0, 5.0, 0.0), (3, 40, 64)),          # round -> round: separable kernels, 16-byte rows
    ((3.0, 3.0, 0.0), (5.0, 5.0, 0.0), (2, 37, 51)),          # ... unaligned rows
    ((3.0, 2.0, 60.0), (7.0, 4.0, 25.0), (2, 48, 56)),        # rotated ellipses: the direct 2-D kernel (21 x 21)
    ((4.0, 2.5, -30.0), (6.0, 6.0, 0.0), (2, 40, 44)),        # ellipse -> round
])
def test_convolve_to_matches_oracle(beam, target, shape, use_dask):
    data = _random_cube(shape, seed=shape[1], nan_frac=0.02)
    for unit in ('K', 'Jy/beam'):
        sc, oc = pair(data, use_dask, beam=beam, unit=unit)
        got = sc.convolve_to(scb().Beam.from_arcsec(*target))
        want = oc.convolve_to(OBeam.arcsec(*target))
        assert_maps_close(got.unmasked_data[:], want._data, rtol=RTOL, atol=1e-6, what='%s %s' % (unit, (beam, target)))
        assert got.unmasked_data[:].dtype == (np.float32 if use_dask else np.float64)      # dask:829 / :2953
        assert got.unit == unit and got.beam == scb().Beam.from_arcsec(*target)
        # the mask object is unchanged: filled data carries the source's NaN pattern
        assert np.array_equal(np.isnan(got.filled_data[:]), np.isnan(data))


def test_convolve_choice_decides_what_an_empty_window_gives():
    """`convolve_fft` (the numpy class's default) returns 0.0 where the kernel window holds no valid input,
    `convolve` (the dask class's) keeps the NaN; both agree wherever the interpolation weight is solid."""
    data = _random_cube((2, 64, 72), seed=21, nan_frac=0.02)
    data[:, 8:56, 10:62] = np.nan                          # an interior hole wider than the 21 x 21 kernel
    target, otarget = scb().Beam.from_arcsec(7.0, 4.0, 25.0), OBeam.arcsec(7.0, 4.0, 25.0)
    kernel = otarget.deconvolve(OBeam.arcsec(3.0, 2.0, 60.0)).as_kernel(abs(G.ADV_WCS['cdelt'][1]))
    assert kernel.shape == (21, 21)
    weight = np.stack([oconv.convolve(np.where(np.isnan(p), -1.0, 0.0), kernel) + 1.0 for p in data.astype(float)])

    def fft_like(array, kernel, **kw):
        raise AssertionError("never called: only its name is read")
    fft_like.__name__ = 'convolve_fft'

    def direct_like(array, kernel, **kw):
        raise AssertionError("never called: only its name is read")
    direct_like.__name__ = 'convolve'

    for use_dask in (False, True):
        sc, oc = pair(data, use_dask, beam=(3.0, 2.0, 60.0), unit='Jy/beam')
        direct = oc.convolve_to(otarget, convolve=oconv.convolve)._data
        empty = np.isnan(direct)
        assert empty.any() and np.all(weight[empty] < 1e-12)
        for conv_arg, zeros in ((None, not use_dask), (fft_like, True), (direct_like, False)):
            got = sc.convolve_to(target, convolve=conv_arg).unmasked_data[:]
            if zeros:
                assert np.all(got[empty] == 0.0)
            else:
                assert np.all(np.isnan(got[empty]))
            # everywhere else the device result is the exact ratio the direct convolution gives
            np.testing.assert_allclose(got[~empty], direct[~empty], rtol=RTOL, atol=1e-6)
        # and the oracle's convolve_fft agrees with both where its weight is not lost in rounding noise
        solid = weight > 1e-6
        fft = oc.convolve_to(otarget, convolve=oconv.convolve_fft)._data
        got = sc.convolve_to(target, convolve=fft_like).unmasked_data[:]
        np.testing.assert_allclose(got[solid], fft[solid], rtol=RTOL, atol=1e-6)
        assert np.all(fft[empty] == 0.0)
    with pytest.raises(NotImplementedError):
        sc.convolve_to(target, convolve=lambda a, k: a)


@use_dask
def test_round_beams_keep_the_interpolation_weight_deep_inside_blank_regions(use_dask):
    """Round beams take the separable kernels.  Their 16 sigma wide factors have outer taps down to e^-32: outputs
    several sigma inside a blank region hang on those taps alone, and must still be the exact weighted mean."""
    data = _random_cube((2, 64, 80), seed=17, nan_frac=0.01)
    data[:, 12:52, 14:66] = np.nan                         # wider than the 15 x 15 kernel
    sc, oc = pair(data, use_dask, beam=(3.0, 3.0, 0.0))
    got = sc.convolve_to(scb().Beam.from_arcsec(5.0), convolve=None if use_dask else _named('convolve')).unmasked_data[:]
    want = oc.convolve_to(OBeam.arcsec(5.0), convolve=oconv.convolve)._data
    kernel = OBeam.arcsec(5.0).deconvolve(OBeam.arcsec(3.0)).as_kernel(abs(G.ADV_WCS['cdelt'][1]))
    assert kernel.shape == (15, 15)
    weight = np.stack([oconv.convolve(np.where(np.isnan(p), -1.0, 0.0), kernel) + 1.0 for p in data.astype(float)])
    assert np.isnan(want).any() and ((weight < 1e-9) & ~np.isnan(want)).any()      # the fringe exists
    assert_maps_close(got, want, rtol=RTOL, atol=1e-6, what='round beam, blank block')


def _named(name):
    def fn(array, kernel, **kw):
        raise AssertionError("never called: only its name is read")
    fn.__name__ = name
    return fn


def test_planes_copied_through_are_neither_rescaled_nor_zeroed():
    """numpy class: a channel with nothing included by the mask is returned as it is, filled
    (spectral_cube.py:161-172), without the Jy/beam factor; the dask class convolves it like any other."""
    data = np.abs(_random_cube((3, 24, 32), seed=13, nan_frac=0.0)) + 1.0
    data[1] = -data[1]                                      # channel 1 fails `> 0` everywhere
    for use_dask in (False, True):
        for fill in (np.nan, 5.0):
            sc, oc = pair(data, use_dask, beam=(3.0, 3.0, 0.0), unit='Jy/beam')
            sc, oc = sc.with_mask(sc > 0.0).with_fill_value(fill), oc.with_mask(oc > 0.0).with_fill_value(fill)
            got = sc.convolve_to(scb().Beam.from_arcsec(5.0)).unmasked_data[:]
            want = oc.convolve_to(OBeam.arcsec(5.0), convolve=oconv.convolve)._data
            if not use_dask:
                if np.isnan(fill):
                    assert np.all(np.isnan(got[1]))
                else:
                    assert np.all(got[1] == 5.0)
                assert_maps_close(got, want, rtol=RTOL, atol=1e-6, what='numpy class, fill=%r' % fill)
            else:
                # the blank plane is convolved too: NaN deep inside, 0.0 where the (valid) zero padding is in reach
                if np.isnan(fill):
                    assert np.isnan(got[1, 12, 16]) and got[1, 0, 0] == 0.0
                assert_maps_close(got, want, rtol=RTOL, atol=1e-6, what='dask class, fill=%r' % fill)


def test_sc_scale_through_the_c_abi():
    """The epilogue kernel on its own: vector and scalar paths, float32 and float64, strided views, flags."""
    import torch
    from spectral_cube_b200 import _lib
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(2)
    for shape, dtype in (((3, 16, 64), np.float32), ((3, 9, 37), np.float32), ((2, 7, 20), np.float64)):
        host = rng.normal(size=shape).astype(dtype)
        host[rng.random(shape) < 0.1] = np.nan
        skip = np.array([0, 1, 0][:shape[0]], dtype=np.uint8)
        for nan_to_zero in (0, 1):
            for use_skip in (False, True):
                dev = torch.from_numpy(host.copy()).cuda()
                flags = torch.from_numpy(skip).cuda() if use_skip else None
                _lib.check(lib.sc_scale(dev.data_ptr(), _lib.F64 if dtype == np.float64 else _lib.F32, *shape,
                                        dev.stride(0), dev.stride(1), 2.5, nan_to_zero,
                                        flags.data_ptr() if use_skip else None, stream))
                want = (host.astype(np.float64) * 2.5).astype(dtype)
                if nan_to_zero:
                    want[np.isnan(host)] = 0.0
                if use_skip:
                    want[skip.astype(bool)] = host[skip.astype(bool)]
                np.testing.assert_array_equal(dev.cpu().numpy(), want)
    # a strided view: the middle columns of a wider block, one channel
    host = rng.normal(size=(2, 6, 40)).astype(np.float32)
    dev = torch.from_numpy(host.copy()).cuda()
    view = dev[1:2, :, 8:24]
    _lib.check(lib.sc_scale(view.data_ptr(), _lib.F32, 1, 6, 16, view.stride(0), view.stride(1), -3.0, 0, None, stream))
    want = host.copy()
    want[1, :, 8:24] = (want[1, :, 8:24].astype(np.float64) * -3.0).astype(np.float32)
    np.testing.assert_array_equal(dev.cpu().numpy(), want)
    assert lib.sc_scale(None, _lib.F32, 1, 1, 1, 1, 1, 2.0, 0, None, stream) != 0


# ---- DaskSpectralCube.statistics (dask_spectral_cube.py:769-814) ---------------------------
@pytest.mark.parametrize('maskname', ['isfinite', 'gt'])
def test_statistics(maskname):
    data = (1.5 + _random_cube((24, 40, 52), seed=8, nan_frac=0.05)).astype(np.float32)
    sc, oc = gpu_cube(data, G.ADV_WCS, use_dask=True), oracle_cube(data, G.ADV_WCS, use_dask=True)
    if maskname == 'gt':
        sc, oc = sc.with_mask(sc > 1.0), oc.with_mask(oc > 1.0)
    got, want = sc.statistics(), oc.statistics()
    assert set(got) == {'npts', 'min', 'max', 'sum', 'sumsq', 'mean', 'sigma', 'rms'} == set(want)
    assert got['npts'] == int(want['npts']) and got['min'] == float(want['min']) and got['max'] == float(want['max'])
    # the reference sums in the data's float32 (pairwise); exact values in float64 here
    filled = np.where(oc._mask_include(), data, np.nan).astype(np.float64)
    for key, exact in (('sum', np.nansum(filled)), ('sumsq', np.nansum(filled * filled))):
        assert np.isclose(got[key], exact, rtol=1e-12), key
    for key in ('sum', 'sumsq', 'mean', 'sigma', 'rms'):
        assert np.isclose(got[key], float(want[key]), rtol=RTOL), (key, got[key], want[key])
    assert not hasattr(gpu_cube(data, G.ADV_WCS, use_dask=False), 'statistics')      # the dask class only


# ---- Projection.convolve_to (lower_dimensional_structures.py:450-494) ------------------------------------------
def test_projection_convolve_to():
    data = _random_cube((6, 40, 48), seed=12, nan_frac=0.02)
    S = scb()
    sc, oc = pair(data, False, beam=(3.0, 2.0, 60.0), unit='Jy/beam')
    m0 = sc.moment0()
    assert m0.beam == sc.beam
    target = S.Beam.from_arcsec(7.0, 4.0, 25.0)
    got = m0.convolve_to(target)
    kernel = OBeam.arcsec(7.0, 4.0, 25.0).deconvolve(OBeam.arcsec(3.0, 2.0, 60.0)).as_kernel(oc._pixscale())
    image = np.asarray(m0.value, dtype=np.float64)
    scale = float(np.nanmax(np.abs(image)))
    assert_maps_close(got.value, oconv.convolve_fft(image, kernel), rtol=RTOL, atol=1e-6 * scale, what='convolve_fft')
    direct = m0.convolve_to(target, convolve=_named('convolve'))
    assert_maps_close(direct.value, oconv.convolve(image, kernel), rtol=RTOL, atol=1e-6 * scale, what='convolve')
    assert isinstance(got, S.Projection) and got.value.dtype == np.float64 and got.unit == m0.unit       # no Jy/beam rescale
    assert got.beam == target and got.header['BMAJ'] == target.major and got.wcs is m0.wcs
    with pytest.raises(ValueError, match="No beam is contained in Projection.meta."):
        gpu_cube(data, G.ADV_WCS).moment0().convolve_to(target)


# ---- the opt-in tiled direct 2-D kernel (SC_DIRECT2D=1) ------------------------------------------------------------
@pytest.mark.parametrize('shape,kshape', [((2, 70, 40), (5, 7)), ((1, 30, 33), (9, 3)), ((3, 130, 97), (21, 21)),
                                          ((1, 64, 64), (45, 45)),
                                          ((1, 80, 72), (77, 77))])            # 77 x 77: the 4-warp, 32-row tile
def test_tiled_direct_kernel_equals_the_direct_kernel(shape, kshape, monkeypatch):
    S = scb()
    rng = np.random.default_rng(kshape[0])
    data = _random_cube(shape, seed=4, nan_frac=0.05)
    data[:, 3:3 + kshape[0] + 4, 5:5 + kshape[1] + 6] = np.nan        # windows with nothing valid
    if kshape[0] == kshape[1]:
        sigma = kshape[0] / 16.0
        kernel = S.EllipticalGaussian2DKernel(sigma, 0.6 * sigma, 0.7, x_size=kshape[1], y_size=kshape[0]).array
    else:
        kernel = rng.random(kshape) + 0.01                            # asymmetric: a missing flip would show
    assert S.BaseSpectralCube._separable_factors(kernel) is None
    sc, oc = gpu_cube(data, G.ADV_WCS, use_dask=True), oracle_cube(data, G.ADV_WCS, use_dask=True)
    sc, oc = sc.with_mask(sc > -1.5), oc.with_mask(oc > -1.5)
    monkeypatch.setenv('SC_DIRECT2D', '0')
    direct = sc.spatial_smooth(S.CustomKernel(kernel)).unmasked_data[:]
    monkeypatch.setenv('SC_DIRECT2D', '1')
    tiled = sc.spatial_smooth(S.CustomKernel(kernel)).unmasked_data[:]
    assert np.isnan(direct).any() or kshape[0] > 45          # (the widest kernel reaches the zero padding from everywhere)
    assert_maps_close(tiled, direct, rtol=2e-6, atol=1e-7, what='tiled vs direct %r' % (kshape,))
    want = oc.spatial_smooth(oconv.Kernel(kernel))._data
    assert_maps_close(tiled, want, rtol=RTOL, atol=1e-6, what='tiled vs oracle %r' % (kshape,))
