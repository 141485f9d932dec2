"""
oracle/mosaic.py -- TEST INFRASTRUCTURE.  ``mosaic_cubes`` / ``combine_headers`` restated in numpy
(spectral_cube/cube_utils.py:744-856).

``reproject.mosaicking.find_optimal_celestial_wcs`` (reproject >= 0.9.1, not installable here) is restated from its
published algorithm: frame of the first input, TAN, unrotated (the reference passes ``auto_rotate=False``,
cube_utils.py:775), reference position = mean of the inputs' reference positions on the unit sphere, resolution = the
finest input pixel scale, CRPIX / NAXIS from the extreme corners of all inputs (corner pixels at -0.5 and n - 0.5;
``crpix = (1 - min) + 0.5``, ``naxis = round(max - min)``).  Pinned by the reference's own test of the function,
``spectral_cube/tests/test_regrid.py:602-634`` (two overlapping parts of one cube mosaic back to the cube's shape and
values, nearest-neighbour, 3 decimals) -- restated in ``tests/test_oracle_goldens.py``; this is also the one
value-carrying pin the reference holds on ``reproject``.
"""
import numpy as np
import scipy.ndimage

from .wcs import OWCS
from . import reproject as _reproj


def optimal_celestial_wcs(inputs):
    """inputs: list of ((ny, nx), OWCS).  Returns (OWCS of the common grid with the first input's spectral axis, (ny, nx))."""
    lons, lats, refs, scales = [], [], [], []
    for (ny, nx), w in inputs:
        xc = np.array([-0.5, nx - 0.5, nx - 0.5, -0.5])
        yc = np.array([-0.5, -0.5, ny - 0.5, ny - 0.5])
        lon, lat = w.celestial_pix2world(xc, yc, origin=0)
        lons.append(lon)
        lats.append(lat)
        refs.append(w.celestial_pix2world(w.crpix[0], w.crpix[1], origin=1))
        m = np.asarray(w.pixel_scale_matrix)[:2, :2]
        scales.append(np.sqrt((m ** 2).sum(axis=0)).min())
    rl = np.radians([float(r[0]) for r in refs])
    rb = np.radians([float(r[1]) for r in refs])
    v = np.array([(np.cos(rb) * np.cos(rl)).mean(), (np.cos(rb) * np.sin(rl)).mean(), np.sin(rb).mean()])
    ref = (np.degrees(np.arctan2(v[1], v[0])) % 360.0, np.degrees(np.arctan2(v[2], np.hypot(v[0], v[1]))))
    res = float(min(scales))
    w0 = inputs[0][1]
    out = OWCS(ctype=[w0.ctype[0][:5] + 'TAN', w0.ctype[1][:5] + 'TAN', w0.ctype[2]],
               crval=[ref[0], ref[1], w0.crval[2]], crpix=[1.0, 1.0, w0.crpix[2]], cdelt=[-res, res, w0.cdelt[2]],
               cunit=['deg', 'deg', w0.cunit[2]])
    xp, yp = out.celestial_world2pix(np.concatenate(lons), np.concatenate(lats), origin=1)
    out.crpix[0] = (1.0 - xp.min()) + 0.5
    out.crpix[1] = (1.0 - yp.min()) + 0.5
    return out, (int(round(yp.max() - yp.min())), int(round(xp.max() - xp.min())))


def sample_nearest(image, yin, xin):
    ny, nx = image.shape
    padded = np.pad(image.astype(np.float64), 1, mode='edge')
    bad = ~np.isfinite(yin) | ~np.isfinite(xin)
    yc, xc = np.where(bad, -10.0, yin), np.where(bad, -10.0, xin)
    vals = scipy.ndimage.map_coordinates(padded, [yc + 1.0, xc + 1.0], order=0, mode='constant', cval=np.nan)
    vals[bad | (yc < -0.5) | (yc > ny - 0.5) | (xc < -0.5) | (xc > nx - 0.5)] = np.nan
    return vals


def mosaic_cubes(cubes, order='bilinear'):
    """cubes: OracleCube objects.  Returns (float64 array (nchan, ny, nx), OWCS of the mosaic)."""
    grid, shape = None, None
    inputs = [(c.shape[1:], c._wcs) for c in cubes]
    # the reference combines pairwise, left to right (cube_utils.py:813-816)
    grid, shape = inputs[0][1], inputs[0][0]
    for nxt in inputs[1:]:
        grid, shape = optimal_celestial_wcs([(shape, grid), nxt])
    nchan = cubes[0].shape[0]
    final = np.zeros((nchan,) + shape)
    weight = np.zeros(shape)
    for c in cubes:
        data = c.unitless_filled_data
        yin, xin = _reproj.input_pixel_coords(c._wcs, grid, shape)
        sample = _reproj.sample_bilinear if order == 'bilinear' else sample_nearest
        rep = np.stack([sample(data[k], yin, xin) for k in range(nchan)])
        weight += (~np.isnan(rep[0])).astype(float)
        final += np.nan_to_num(rep)
    with np.errstate(divide='ignore', invalid='ignore'):
        final = final / weight
    return final, grid
