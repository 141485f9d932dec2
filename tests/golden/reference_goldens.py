"""
Golden vectors re-typed from the reference's own tests (values only; the reference
cannot be imported here).  Each block cites where it comes from in
/root/reference/spectral_cube/tests.
"""
import numpy as np

# --- test_moments.py:12-48 --------------------------------------------------------------------
# cube = arange(27).reshape(3,3,3); WCS RA---TAN/DEC--TAN/VELO, cdelt = float32([-1,2,3])/1e5,
# crpix = 1, crval = float32([0, 1e-3, 2e-3]), cunit = deg, deg, km/s (test_moments.py:56-70).
# Units after WCS normalisation: velocity in m/s, angles in deg.
DV = 3e-2      # m/s
DY = 2e-5      # deg
DX = 1e-5      # deg

M0V = np.array([[27, 30, 33], [36, 39, 42], [45, 48, 51]]) * DV
M0Y = np.array([[9, 12, 15], [36, 39, 42], [63, 66, 69]]) * DY
M0X = np.array([[3, 12, 21], [30, 39, 48], [57, 66, 75]]) * DX
M1V = np.array([[1.66666667, 1.6, 1.54545455],
                [1.5, 1.46153846, 1.42857143],
                [1.4, 1.375, 1.35294118]]) * DV + 2.0
M1Y = np.array([[1.66666667, 1.5, 1.4],
                [1.16666667, 1.15384615, 1.14285714],
                [1.0952381, 1.09090909, 1.08695652]]) * DY
M1X = np.array([[1.66666667, 1.16666667, 1.0952381],
                [1.06666667, 1.05128205, 1.04166667],
                [1.03508772, 1.03030303, 1.02666667]]) * DX
M2V = np.array([[0.22222222, 0.30666667, 0.36914601],
                [0.41666667, 0.45364892, 0.4829932],
                [0.50666667, 0.52604167, 0.54209919]]) * DV ** 2
M2Y = np.array([[0.22222222, 0.41666667, 0.50666667],
                [0.63888889, 0.64299803, 0.6462585],
                [0.65759637, 0.6584022, 0.65910523]]) * DY ** 2
M2X = np.array([[0.22222222, 0.63888889, 0.65759637],
                [0.66222222, 0.66403682, 0.66493056],
                [0.66543552, 0.66574839, 0.66595556]]) * DX ** 2
MOMENTS = [[M0V, M0Y, M0X], [M1V, M1Y, M1X], [M2V, M2Y, M2X]]
MOMENT_UNITS = [['K m/s', 'K deg', 'K deg'], ['m/s', 'deg', 'deg'], ['m/s2', 'deg2', 'deg2']]

VARIANCE_WARNING_TEXT = ("Note that the second moment returned will be a "
                         "variance map. To get a linewidth map, use the "
                         "SpectralCube.linewidth_fwhm() or "
                         "SpectralCube.linewidth_sigma() methods instead.")


def moment_cube_data():
    return np.arange(27).reshape([3, 3, 3]).astype(float)


MOMENT_WCS = dict(
    ctype=['RA---TAN', 'DEC--TAN', 'VELO'],
    cdelt=(np.array([-1, 2, 3], dtype='float32') / 1e5).astype(np.float64),
    crpix=np.array([1, 1, 1], dtype='float32').astype(np.float64),
    crval=np.array([0, 1e-3, 2e-3], dtype='float32').astype(np.float64),
    cunit=['deg', 'deg', 'km/s'],
)

# --- conftest.py:259-271 (prepare_adv_data) + tests/data/header_jybeam.hdr ----------------------
# d = np.random.seed(96); np.random.random((4, 3, 2)); BUNIT 'K'


def adv_data():
    rs = np.random.RandomState(96)
    return rs.random_sample((4, 3, 2))


ADV_WCS = dict(
    ctype=['RA---SIN', 'DEC--SIN', 'VOPT'],
    cdelt=[-5.55555561268E-04, 5.55555561268E-04, 1.28821496879E+00],
    crpix=[1.37300000000E+03, 1.15200000000E+03, 1.0],
    crval=[2.31837500515E+01, 3.05765277962E+01, -3.21214698632E+02],
    cunit=['deg', 'deg', 'km/s'],
)

# test_spectral_cube.py:2363-2383  cube.spatial_smooth(Gaussian2DKernel(3)), 7 decimals
G2D_RESULT0 = np.array([[0.0585795, 0.0588712],
                        [0.0612525, 0.0614312],
                        [0.0576757, 0.057723]])
G2D_RESULT2 = np.array([[0.027322, 0.027257],
                        [0.0280423, 0.02803],
                        [0.0259688, 0.0260123]])
# test_spectral_cube.py:2401-2421  cube.spatial_smooth(Tophat2DKernel(3)), 7 decimals
T2D_RESULT0 = np.full((3, 2), 0.1265607)
T2D_RESULT2 = np.full((3, 2), 0.0585135)

# --- conftest.py:473-479 (data_522_delta) ------------------------------------------------------


def delta_522():
    d = np.zeros([5, 2, 2], dtype='float')
    d[2, :, :] = 1.0
    return d


def delta_255():
    d = np.zeros([2, 5, 5], dtype='float')
    d[0, 2, 2] = 1.0
    return d


# test_regrid.py:138-172: result[:,0,0] == Gaussian1DKernel(1.0).array[2:-2] to 4 decimals, size 9
# test_regrid.py:234-248: midpoints -> [0.0, 0.5, 0.5, 0.0]
INTERP_MIDPOINTS = np.array([0.0, 0.5, 0.5, 0.0])
# test_regrid.py:292-303: grid stepping out of bounds with fill_value=42 -> 42 everywhere
# test_regrid.py:318-345: CDELT3 negated, mask[:2]=False, grid = midpoints reversed
INTERP_WITH_MASK = np.array([0.0, 0.5, np.nan, np.nan])
# test_regrid.py:251-270 (dask, 2x5x5 delta at [0,2,2]): midpoint -> [0.5] at [:,2,2]
