"""
Host plumbing of `convolve_to` without a device: tensors stay on the host, the REAL libsc_b200.so is called, so
ctypes marshals every argument and the library's own argument validation (SC_CHECK_ARG) runs; each call then stops
at its first CUDA API call ("no device"), which is tolerated here and only here (the `host` fixture of
tests/conftest.py).  What this pins: class
construction and beam propagation, the per-channel plan of VaryingResolutionSpectralCube, which C-ABI entry
points a call reaches and in what number, and that the library accepts the shapes / strides / flags it is given.
Numerical parity is the GPU tests' business (tests/test_convolve_to_gpu.py).
"""
import warnings

import numpy as np
import pytest

from tests.golden import reference_goldens as G


def make(S, data, use_dask, beam=None, unit='K'):
    from spectral_cube_b200.masks import LazyMask
    cls = S.DaskSpectralCube if use_dask else S.SpectralCube
    cube = cls(np.asarray(data, dtype=np.float32), S.CubeWCS(**G.ADV_WCS), unit=unit, beam=beam)
    cube._mask = LazyMask(np.isfinite, cube=cube)
    return cube


@pytest.mark.parametrize('use_dask', [False, True])
@pytest.mark.parametrize('shape,beam,target', [((3, 40, 64), (3., 3., 0.), (5., 5., 0.)),
                                               ((2, 37, 51), (3., 2., 60.), (7., 4., 25.))])
def test_single_beam_cube_reaches_the_library(host, use_dask, shape, beam, target):
    S, calls = host
    rng = np.random.default_rng(0)
    for unit in ('K', 'Jy/beam'):
        cube = make(S, rng.normal(size=shape), use_dask, S.Beam.from_arcsec(*beam), unit)
        cube = cube.with_mask(cube > -10.0).with_fill_value(5.0)
        del calls[:]
        out = cube.convolve_to(S.Beam.from_arcsec(*target))
        # one smoothing call; the epilogue (sc_scale) runs for Jy/beam and for the numpy class's convolve_fft rule
        assert len(calls) == (2 if (unit == 'Jy/beam' or not use_dask) else 1), calls
        assert all('CUDA error' in msg for rc, msg in calls)
        assert type(out) is type(cube) and out.shape == cube.shape and out.unit == unit
        assert out.beam == S.Beam.from_arcsec(*target) and out.meta['beam'] == out.beam
        assert out.header['BMAJ'] == out.beam.major and out.header['BPA'] == out.beam.pa
        assert cube.beam == S.Beam.from_arcsec(*beam)                       # the source keeps its own beam
        assert out.mask is cube.mask
        assert out.with_fill_value(0.0).beam == out.beam                     # carried by _new_cube_with
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        assert cube.convolve_to(cube.beam) is cube
    assert any("identical to the current beam" in str(x.message) for x in w)
    with pytest.raises(NotImplementedError):
        cube.convolve_to(S.Beam.from_arcsec(*target), nan_treatment='fill')
    with pytest.raises(TypeError):
        cube.convolve_to(S.Beam.from_arcsec(*target), no_such_keyword=1)
    with pytest.raises(S.NoBeamError):
        make(S, rng.normal(size=(2, 5, 5)), use_dask).convolve_to(S.Beam.from_arcsec(2.0))
    with pytest.raises(S.BeamUnitsError):
        make(S, rng.normal(size=(2, 5, 5)), use_dask, unit='Jy/beam').with_beam(S.Beam.from_arcsec(2.0))
    hdr_cube = S.SpectralCube(np.zeros((2, 5, 5), dtype=np.float32), S.CubeWCS(**G.ADV_WCS),
                              header={'BMAJ': 1 / 3600., 'BMIN': 1 / 3600., 'BPA': 0.0}, use_dask=use_dask)
    assert hdr_cube.beam == S.Beam.from_arcsec(1.0)                          # cube_utils.try_load_beam


@pytest.mark.parametrize('use_dask', [False, True])
def test_varying_resolution_cube_reaches_the_library(host, use_dask):
    S, calls = host
    from spectral_cube_b200.masks import LazyMask
    rng = np.random.default_rng(1)
    beams = S.Beams.from_arcsec([0.4, 0.3, 0.3, 0.4], [0.1, 0.2, 0.2, 0.1], [0, 45, 60, 30])
    data = rng.normal(size=(4, 20, 24)).astype(np.float32)
    vr = S.VaryingResolutionSpectralCube(data, S.CubeWCS(**G.ADV_WCS), unit='Jy/beam', beams=beams, use_dask=use_dask)
    assert type(vr) is (S.DaskVaryingResolutionSpectralCube if use_dask else S.VaryingResolutionSpectralCube)
    vr._mask = LazyMask(np.isfinite, cube=vr)
    del calls[:]
    out = vr.convolve_to(S.Beam.from_arcsec(0.5))
    assert len(calls) == 8                                       # per channel: one smoothing call + the Jy/beam rescale
    assert type(out) is (S.DaskSpectralCube if use_dask else S.SpectralCube)
    assert out.beam == S.Beam.from_arcsec(0.5) and out.shape == vr.shape and out.mask is vr.mask
    with pytest.raises(ValueError, match="Beam could not be deconvolved"):
        vr.convolve_to(S.Beam.from_arcsec(0.35))
    with pytest.raises(S.NoBeamError):
        vr.beam
    masked = vr.mask_channels([False, True, True, False])
    assert type(masked) is type(vr) and list(masked.goodbeams_mask) == [False, True, True, False]
    assert len(masked.beams) == 2 and len(masked.unmasked_beams) == 4
    del calls[:]
    masked.convolve_to(S.Beam.from_arcsec(0.35))
    assert len(calls) == 2 + 2 * 2                               # two filled copies, two channels convolved + rescaled
    del calls[:]
    vr.convolve_to(S.Beam.from_arcsec(0.35), allow_smaller=True)
    assert len(calls) == 2 + 2 * 2
    sub = vr[1:3]
    assert sub.shape == (2, 20, 24) and len(sub.unmasked_beams) == 2 and sub.unmasked_beams[0] == beams[1]
    for method in (vr.spectral_smooth, vr.spectral_interpolate):
        with pytest.raises(AttributeError):
            method(None)
    with pytest.raises(ValueError, match="Beam list must have same size"):
        S.VaryingResolutionSpectralCube(data, S.CubeWCS(**G.ADV_WCS), beams=beams[:3], use_dask=use_dask)
    with pytest.raises(ValueError, match="Must give either a beam table"):
        S.VaryingResolutionSpectralCube(data, S.CubeWCS(**G.ADV_WCS), use_dask=use_dask)
    # a CASA-style beam table (arcsec) with a non-finite row: that channel is masked out and skipped
    table = dict(BMAJ=[0.4, np.nan, 0.3, 0.4], BMIN=[0.1, 0.2, 0.2, 0.1], BPA=[0, 45, 60, 30])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        vt = S.VaryingResolutionSpectralCube(data, S.CubeWCS(**G.ADV_WCS), beam_table=table, use_dask=use_dask)
    assert any("non-finite beams" in str(x.message) for x in w)
    assert list(vt.goodbeams_mask) == [True, False, True, True] and vt.mask is not None
    del calls[:]
    vt.convolve_to(S.Beam.from_arcsec(0.5))
    assert len(calls) == 1 + 3 * (1 if use_dask else 2)        # unit K: the epilogue only applies convolve_fft's rule
    # rotated pixel axes: the reference warns that the kernels ignore it (:4173-4183)
    w2 = dict(G.ADV_WCS)
    w2['pc'] = [[0.8, -0.6, 0.0], [0.6, 0.8, 0.0], [0.0, 0.0, 1.0]]
    rot = S.VaryingResolutionSpectralCube(data, S.CubeWCS(**w2), beams=beams, use_dask=use_dask)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        rot.convolve_to(S.Beam.from_arcsec(0.5))
    assert any(issubclass(x.category, S.BeamWarning) for x in w)
    assert abs(rot._pixscale_deg() - abs(G.ADV_WCS['cdelt'][1])) < 1e-15


def test_projection_convolve_to_reaches_the_library(host):
    """lower_dimensional_structures.py:450-494"""
    S, calls = host
    rng = np.random.default_rng(2)
    cube = make(S, rng.normal(size=(6, 24, 32)), False, S.Beam.from_arcsec(3.0, 2.0, 60.0), 'Jy/beam')
    m0 = cube.moment0()
    assert m0.beam == cube.beam and hasattr(m0, 'beam')
    del calls[:]
    out = m0.convolve_to(S.Beam.from_arcsec(7.0, 4.0, 25.0))
    assert len(calls) == 2 and isinstance(out, S.Projection) and out.shape == m0.shape and out.dtype == np.float64
    assert out.beam == S.Beam.from_arcsec(7.0, 4.0, 25.0) and out.unit == m0.unit and out.wcs is m0.wcs
    assert out.header['BMAJ'] == out.beam.major and m0.beam == cube.beam
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        assert m0.convolve_to(m0.beam) is m0
    assert any("identical to the current beam" in str(x.message) for x in w)
    nobeam = make(S, rng.normal(size=(6, 24, 32)), False).moment0()
    assert not hasattr(nobeam, 'beam')
    with pytest.raises(ValueError, match="No beam is contained in Projection.meta."):
        nobeam.convolve_to(S.Beam.from_arcsec(7.0))
    with pytest.raises(ValueError, match="two spatial axes"):
        cube.moment(order=0, axis=1).convolve_to(S.Beam.from_arcsec(7.0))


@pytest.mark.parametrize('use_dask', [False, True])
def test_row_sharded_convolve_to_reaches_the_library(host, use_dask):
    """One process (no process group): the sharded wrapper degenerates to the whole image; the exchange functions
    themselves are covered by the gloo tests in tests/test_distributed_cpu.py."""
    S, calls = host
    from spectral_cube_b200 import distributed as D
    rng = np.random.default_rng(3)
    cube = make(S, rng.normal(size=(3, 40, 64)), use_dask, S.Beam.from_arcsec(3.0), 'Jy/beam')
    cube = cube.with_mask(cube > -10.0)
    sh = D.RowShardedCube(cube, 40, 0)
    del calls[:]
    out = sh.convolve_to(S.Beam.from_arcsec(5.0))
    # one smoothing call and the epilogue (no neighbours: no halo rows packed)
    assert len(calls) == 2 and isinstance(out, D.RowShardedCube) and out.local.beam == S.Beam.from_arcsec(5.0)
    assert type(out.local) is type(cube) and out.shape == (3, 40, 64)
    del calls[:]
    sm = sh.spatial_smooth(S.Gaussian2DKernel(1.0), raise_error_jybm=False)
    assert len(calls) == 2 and sm.local.beam == cube.beam            # the strategy sample + the smoothing call
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        assert sh.convolve_to(cube.beam) is sh
    assert w
