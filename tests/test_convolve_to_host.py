"""
`convolve_to`, CPU side: the oracle's restatement of radio_beam / convolve_fft against
what the reference's own tests hold (spectral_cube/tests/test_regrid.py:33-96, conftest.py:590-660 with
tests/test_spectral_cube.py:2150-2225), an independent covariance-matrix derivation of the deconvolution, and
the product's host-side beam arithmetic (spectral_cube_b200/beam.py) against the oracle's.  No GPU.
"""
import warnings

import numpy as np
import pytest

from oracle import beam as obeam
from oracle import convolve as oconv
from oracle.beam import OBeam
from oracle.cube import OracleCube
from oracle.wcs import OWCS
from spectral_cube_b200 import beam as pbeam
from tests.golden import reference_goldens as G

PIX = 5.55555561268E-04        # header_jybeam.hdr: 2 arcsec pixels, BMAJ = BMIN = 1 arcsec
use_dask = pytest.mark.parametrize('use_dask', [False, True])

# conftest.py:482-497 / :565-577
BEAMS5 = dict(major=[0.5, 0.4, 0.3, 0.4, 0.5], minor=[0.1, 0.2, 0.3, 0.2, 0.1], pa=[0, 45, 60, 30, 0])
BEAMS5_PIX = dict(major=[3.5, 3, 3, 3, 3], minor=[2, 2.5, 3, 2.5, 2], pa=[0, 45, 60, 30, 0])
# conftest.py `prepare_4_beams`
BEAMS4 = dict(major=[0.4, 0.3, 0.3, 0.4], minor=[0.1, 0.2, 0.2, 0.1], pa=[0, 45, 60, 30])


def ocube(data, unit='K', use_dask=False, **kw):
    return OracleCube(np.asarray(data, dtype=float), OWCS(**G.ADV_WCS), unit=unit, use_dask=use_dask, **kw)


def covariance(major, minor, pa_deg):
    """Covariance (x = east-ish pixel axis after as_kernel's 90 deg shift is irrelevant here) of a Gaussian
    with the given FWHM axes whose major axis makes the angle pa with the +y axis, counter-clockwise."""
    t = np.deg2rad(pa_deg + 90.0)
    r = np.array([[np.cos(t), -np.sin(t)], [np.sin(t), np.cos(t)]])
    return r @ np.diag([major ** 2, minor ** 2]) @ r.T / (8 * np.log(2))


# ---- the deconvolution formula against linear algebra ------------------------------------------------
@pytest.mark.parametrize('target,other', [((10., 10., 0.), (7., 4., 0.)), ((10., 10., 0.), (6., 5., 45.)),
                                          ((10., 8., 30.), (6., 5., 60.)), ((9., 7., 100.), (6.9, 3., 95.)),
                                          ((12., 6., -20.), (5.5, 5.5, 0.))])
def test_deconvolved_beam_convolved_with_the_other_gives_the_target(target, other):
    d = OBeam(*target).deconvolve(OBeam(*other))
    assert d.major >= d.minor
    np.testing.assert_allclose(covariance(d.major, d.minor, d.pa) + covariance(*other), covariance(*target),
                               rtol=1e-10, atol=1e-12)


def test_which_beams_cannot_be_deconvolved():
    """tests/test_spectral_cube.py:2204-2225 (vda_beams: the largest beam is 0.4 arcsec)."""
    beams = [OBeam.arcsec(a, b, p) for a, b, p in zip(BEAMS4['major'], BEAMS4['minor'], BEAMS4['pa'])]
    for bm in beams:
        OBeam.arcsec(0.5).deconvolve(bm)
    with pytest.raises(ValueError, match="Beam could not be deconvolved"):
        for bm in beams:
            OBeam.arcsec(0.35).deconvolve(bm)
    for bm in beams[1:3]:                       # the two middle beams are smaller than 0.35 arcsec
        OBeam.arcsec(0.35).deconvolve(bm)
    with pytest.raises(ValueError, match="Beam could not be deconvolved"):
        OBeam.arcsec(1.0).deconvolve(OBeam.arcsec(1.0))


# ---- reference goldens -------------------------------------------------------------------------------
@use_dask
def test_convolution(use_dask):
    """tests/test_regrid.py:33-58: 1 arcsec convolved with 1.5 arcsec -> 1.8027..."""
    d = np.zeros([2, 5, 5])
    d[0, 2, 2] = 1.0
    cube = ocube(d, use_dask=use_dask, beam=OBeam.arcsec(1.0))
    conv = cube.convolve_to(OBeam.arcsec(1.802775637731995))
    expected = oconv.Gaussian2DKernel(1.5 / 3600. / obeam.SIGMA_TO_FWHM / 5.555555555555e-4, x_size=5, y_size=5).array
    np.testing.assert_almost_equal(expected / expected.sum(), conv._get_filled_data(fill=np.nan)[0])
    assert np.all(conv._get_filled_data(fill=np.nan)[1] == 0.0)            # 2nd layer is all zeros
    assert conv.beam == OBeam.arcsec(1.802775637731995)


@use_dask
def test_beams_convolution(use_dask):
    """tests/test_regrid.py:61-82: each channel equals its own deconvolved kernel (5x5, normalised)."""
    d = np.zeros([4, 5, 5])
    d[:, 2, 2] = 1.0
    beams = [OBeam.arcsec(a, b, p) for a, b, p in zip(BEAMS4['major'], BEAMS4['minor'], BEAMS4['pa'])]
    cube = ocube(d, use_dask=use_dask, beams=beams)
    target = OBeam.arcsec(1.802775637731995)
    conv = cube.convolve_to(target)
    for ii, bm in enumerate(beams):
        expected = target.deconvolve(bm).as_kernel(cube._pixscale(), x_size=5, y_size=5).array
        np.testing.assert_almost_equal(expected / expected.sum(), conv._get_filled_data(fill=np.nan)[ii])


@use_dask
def test_beams_convolution_equal(use_dask):
    """tests/test_regrid.py:85-101: a channel already at the target beam is not convolved."""
    d = np.zeros([5, 2, 2])
    d[2] = 1.0
    beams = [OBeam.arcsec(a, b, p) for a, b, p in zip(BEAMS5['major'], BEAMS5['minor'], BEAMS5['pa'])]
    beams[0] = OBeam.arcsec(1.0, 1.0, 0.0)
    cube = ocube(d, use_dask=use_dask, beams=beams)
    conv = cube.convolve_to(OBeam.arcsec(1.0, 1.0, 0.0))
    np.testing.assert_almost_equal(cube._get_filled_data(fill=np.nan)[0], conv._get_filled_data(fill=np.nan)[0])


def point_source_fixture(beams):
    """conftest.py:590-660: a point source convolved to each channel's beam, in Jy/beam (peak = 1)."""
    d = np.zeros((5, 11, 11))
    d[:, 5, 5] = 1.
    pix = 2. / 3600.
    for i, bm in enumerate(beams):
        d[i] = oconv.convolve_fft(d[i], bm.as_kernel(pix))
        d[i] *= bm.sr / np.deg2rad(pix) ** 2
    np.testing.assert_allclose(d[:, 5, 5], 1., atol=1e-5)                 # the fixture's own check (:611, :653)
    return d


@use_dask
def test_convolve_to_jybeam_onebeam(use_dask):
    """tests/test_spectral_cube.py:2181-2189"""
    bm = OBeam.arcsec(6.0)
    d = point_source_fixture([bm] * 5)
    cube = ocube(d, unit='Jy/beam', use_dask=use_dask, beam=bm)
    conv = cube.convolve_to(OBeam.arcsec(10.0))
    np.testing.assert_allclose(conv._data[:, 5, 5], d[:, 5, 5], atol=1e-5, rtol=1e-5)


@use_dask
def test_convolve_to_jybeam_multibeams(use_dask):
    """tests/test_spectral_cube.py:2192-2201: five rotated elliptical beams to one round 10 arcsec beam.  The peak
    stays 1 Jy/beam only if the deconvolved position angle and `as_kernel`'s angle convention agree."""
    beams = [OBeam.arcsec(2 * a, 2 * b, p) for a, b, p in zip(BEAMS5_PIX['major'], BEAMS5_PIX['minor'], BEAMS5_PIX['pa'])]
    d = point_source_fixture(beams)
    cube = ocube(d, unit='Jy/beam', use_dask=use_dask, beams=beams)
    conv = cube.convolve_to(OBeam.arcsec(10.0))
    np.testing.assert_allclose(conv._data[:, 5, 5], d[:, 5, 5], atol=1e-5, rtol=1e-5)
    # and the result is the target beam itself, sampled: a round Gaussian of FWHM 10 arcsec = 5 pixels
    yy, xx = np.mgrid[-5:6, -5:6]
    want = np.exp(-0.5 * (xx ** 2 + yy ** 2) / (5 / obeam.SIGMA_TO_FWHM) ** 2)
    np.testing.assert_allclose(conv._data[1][3:8, 3:8], want[3:8, 3:8], atol=2e-3)


@use_dask
def test_convolve_to_equal_and_with_bad_beams(use_dask):
    """tests/test_spectral_cube.py:2150-2157, 2204-2225"""
    rng = np.random.default_rng(5)
    d = rng.normal(size=(4, 6, 7))
    cube = ocube(d, use_dask=use_dask, beam=OBeam.arcsec(1.0))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        assert cube.convolve_to(OBeam.arcsec(1.0)) is cube
    beams = [OBeam.arcsec(a, b, p) for a, b, p in zip(BEAMS4['major'], BEAMS4['minor'], BEAMS4['pa'])]
    vr = ocube(d, use_dask=use_dask, beams=beams)
    vr.convolve_to(OBeam.arcsec(0.5))
    with pytest.raises(ValueError, match="Beam could not be deconvolved"):
        vr.convolve_to(OBeam.arcsec(0.35))
    masked = vr.mask_channels([False, True, True, False])
    conv = masked.convolve_to(OBeam.arcsec(0.35))
    assert np.all(np.isfinite(conv._data[1:3]))


# ---- convolve_fft against the direct convolution -----------------------------------------------------------
def test_convolve_fft_equals_convolve_except_where_nothing_is_valid():
    rng = np.random.default_rng(11)
    img = rng.normal(size=(60, 72))
    img[rng.random(img.shape) < 0.1] = np.nan
    img[10:50, 12:60] = np.nan                                # an interior hole wider than the kernel
    k = OBeam.arcsec(7.0, 4.0, 25.0).as_kernel(2. / 3600.)
    assert k.shape == (23, 23)                                # 16 sigma of the larger axis-aligned extent, odd
    direct = oconv.convolve(img, k, normalize_kernel=True)
    fft = oconv.convolve_fft(img, k, normalize_kernel=True)
    hole = np.isnan(direct)
    assert hole.any() and hole.sum() < img.size // 4
    assert np.all(fft[hole] == 0.0)
    # interpolation weight of every output: the normalised kernel over the valid inputs (the zero padding is valid)
    weight = oconv.convolve(np.where(np.isnan(img), -1.0, 0.0), k, normalize_kernel=True) + 1.0
    solid = weight > 1e-6
    assert solid.sum() > img.size // 2 and not solid[hole].any()
    np.testing.assert_allclose(fft[solid], direct[solid], rtol=1e-9, atol=1e-9)
    # deeper inside the hole only kernel tails beyond ~5 sigma reach valid data: the FFT result there is rounding
    # noise over a vanishing weight (or 0.0 below 10 eps), the direct ratio stays exact -- they are not comparable
    fringe = ~solid & ~hole
    assert fringe.any() and np.all(np.isfinite(direct[fringe]))


# ---- the product's host arithmetic against the oracle's ------------------------------------------------
@pytest.mark.parametrize('target,other', [((10., 10., 0.), (7., 4., 0.)), ((10., 10., 0.), (6., 5., 45.)),
                                          ((10., 8., 30.), (6., 5., 60.)), ((9., 7., 100.), (6.9, 3., 95.)),
                                          ((1.802775637731995, 1.802775637731995, 0.), (1., 1., 0.))])
def test_product_beam_matches_the_oracle(target, other):
    pt, po = pbeam.Beam.from_arcsec(*target), pbeam.Beam.from_arcsec(*other)
    ot, oo = OBeam.arcsec(*target), OBeam.arcsec(*other)
    pd, od = pt.deconvolve(po), ot.deconvolve(oo)
    np.testing.assert_allclose([pd.major, pd.minor, pd.pa], [od.major, od.minor, od.pa], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(pt.sr, ot.sr, rtol=1e-14)
    for pix in (PIX, 0.7 / 3600.):
        pk, ok = pd.as_kernel(pix).array, od.as_kernel(pix).array
        assert pk.shape == ok.shape and pk.shape[0] % 2 == 1
        np.testing.assert_allclose(pk, ok, rtol=1e-12, atol=1e-300)
    pk5 = pd.as_kernel(PIX, x_size=5, y_size=7).array
    assert pk5.shape == (7, 5)
    np.testing.assert_allclose(pk5, od.as_kernel(PIX, x_size=5, y_size=7).array, rtol=1e-12, atol=1e-300)


def test_product_beam_equality_header_and_errors():
    B = pbeam.Beam
    assert B.from_arcsec(1.0) == B(1.0 / 3600.) and B.from_arcsec(2.0, 1.0, 10.) == B.from_arcsec(2.0, 1.0, 190.)
    assert B.from_arcsec(2.0, 1.0, 10.) != B.from_arcsec(2.0, 1.0, 20.)
    assert B.from_arcsec(2.0, 2.0, 10.) == B.from_arcsec(2.0, 2.0, 77.)            # round: pa is ignored
    with pytest.raises(ValueError, match="Minor axis greater than major axis."):
        B(1.0, 2.0)
    with pytest.raises(pbeam.BeamError, match="Beam could not be deconvolved"):
        B.from_arcsec(0.35).deconvolve(B.from_arcsec(0.4, 0.1, 0.))
    assert issubclass(pbeam.BeamError, ValueError)
    hdr = B.from_arcsec(3.0, 2.0, 30.).to_header_keywords()
    assert B.from_fits_header(hdr) == B.from_arcsec(3.0, 2.0, 30.)
    with pytest.raises(pbeam.NoBeamError):
        B.from_fits_header({})
    beams = pbeam.Beams.from_arcsec([0.5, np.nan, 0.3], [0.1, 0.2, 0.3], [0, 45, 60])
    assert list(beams.isfinite) == [True, False, True] and len(beams[beams.isfinite]) == 2
    assert beams[0] == B.from_arcsec(0.5, 0.1, 0.) and isinstance(beams[1:], pbeam.Beams)

    class Quantity(float):                       # duck-typed astropy Quantity (a real radio_beam.Beam's attributes)
        def to_value(self, unit):
            assert unit == 'deg'
            return float(self)

    class RadioBeam(object):
        major, minor, pa = Quantity(2 / 3600.), Quantity(1 / 3600.), Quantity(15.)

    assert B.coerce(RadioBeam()) == B.from_arcsec(2.0, 1.0, 15.)
    with pytest.raises(TypeError, match="beam must be a radio_beam.Beam object."):
        B.coerce(3.0)
