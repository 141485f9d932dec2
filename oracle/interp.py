"""
oracle/interp.py -- TEST INFRASTRUCTURE.  Spectral resampling, both reference variants.

``spectral_interpolate_numpy`` follows ``spectral_cube/spectral_cube.py:3224-3332``
(per-spaxel ``numpy.interp`` with ``left=right=fill_value``; new mask =
``interp(mask) > 0`` unless the whole ray is included; rays with nothing included
become NaN / False; reversed input axis and reversed output grid handling
:3259-3269, :3299-3310).
``spectral_interpolate_dask`` follows ``spectral_cube/dask_spectral_cube.py:1250-1373``
(``scipy.interpolate.interp1d(..., fill_value=fill_value, bounds_error=False)`` per
block, NaN outside the input range unless ``fill_value`` is given; new mask =
``~isnan(newcube)`` taken *before* the output is flipped back, :1364-1367).
numpy and scipy are importable here, so the real ``np.interp`` / ``interp1d`` are
called -- nothing about them is restated.

Both take plain arrays: ``data`` = the cube's filled data (mask -> fill value),
``include`` = the boolean include mask, ``inaxis``/``grid`` = spectral coordinates
in one common unit.  They return ``(newdata, newmask, outdiff_sign)``.
"""
import numpy as np
import scipy.interpolate


def _orient(inaxis, grid):
    inaxis = np.asarray(inaxis, dtype=np.float64)
    grid = np.asarray(grid, dtype=np.float64)
    reverse_out = np.mean(np.diff(grid)) < 0
    reverse_in = np.mean(np.diff(inaxis)) < 0
    if reverse_out:
        grid = grid[::-1]
    if reverse_in:
        inaxis = inaxis[::-1]
    assert np.all(np.diff(grid) > 0)
    assert np.all(np.diff(inaxis) > 0)
    np.testing.assert_allclose(np.diff(grid), np.mean(np.diff(grid)),
                               err_msg="Output grid must be linear")
    return inaxis, grid, reverse_in, reverse_out


def spectral_interpolate_numpy(data, include, inaxis, grid, fill_value=None):
    inaxis, grid, reverse_in, reverse_out = _orient(inaxis, grid)
    specslice = slice(None, None, -1) if reverse_in else slice(None)
    outslice = slice(None, None, -1) if reverse_out else slice(None)
    nout = grid.size
    _, ny, nx = data.shape
    newcube = np.empty((nout, ny, nx), dtype=data.dtype)
    newmask = np.empty((nout, ny, nx), dtype=bool)
    for iy in range(ny):
        for ix in range(nx):
            m = include[specslice, iy, ix]
            if m.any():
                newcube[outslice, iy, ix] = np.interp(grid, inaxis, data[specslice, iy, ix],
                                                      left=fill_value, right=fill_value)
                if m.all():
                    newmask[:, iy, ix] = True
                else:
                    newmask[outslice, iy, ix] = np.interp(grid, inaxis, m) > 0
            else:
                newmask[:, iy, ix] = False
                newcube[:, iy, ix] = np.nan
    return newcube, newmask, reverse_out


def spectral_interpolate_dask(data_nanfilled, inaxis, grid, fill_value=None):
    inaxis, grid, reverse_in, reverse_out = _orient(inaxis, grid)
    y = data_nanfilled[::-1] if reverse_in else data_nanfilled
    f = scipy.interpolate.interp1d(inaxis, y.T, fill_value=fill_value, bounds_error=False)
    newcube = f(grid).T
    newmask = ~np.isnan(newcube)          # taken before the flip (dask:1364)
    if reverse_out:
        newcube = newcube[::-1]
    return newcube, newmask, reverse_out
