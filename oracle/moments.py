"""
oracle/moments.py -- TEST INFRASTRUCTURE.  numpy restatement of the reference moment engine.

What is restated (reference = spectral_cube/_moments.py unless noted):
  * ``how='cube'``  (:170-193)  whole-cube nansum on ``filled(NaN) * pix_size``; order 0
    goes through ``allbadtonan`` (np_compat.py:3-27) so an all-NaN ray gives NaN.
  * ``how='slice'`` (:30-125)   plane-by-plane accumulation; order 0 fills with NaN and
    tracks a ``valid`` plane (:48-54), order >= 1 fills with 0 (:81, :118) and makes a
    second pass around the order-1 map for higher orders (:108-123).
  * ``how='ray'``   (:128-168)  one spectrum at a time over included voxels only;
    rays with nothing included stay NaN.
  * ``how='auto'``  (:196-202)  + ``iterator_strategy`` (cube_utils.py:277-301).
  * the dask class (dask_spectral_cube.py:1083-1101): explicit float64 cast and
    ``nansum_allbadtonan`` (:54-59) for numerator *and* denominator.
All return a float64 array of shape ``cube.shape`` minus ``axis``, before units and
before the ``+ world[0]`` shift that ``moment()`` applies (spectral_cube.py:1709-1710).

``cube`` is an ``oracle.cube.OracleCube``.
"""
import numpy as np

MEMORY_THRESHOLD = 1e8          # cube_utils.py:266-268


def nansum_allbad_nan(a, axis):
    """nansum along ``axis`` where fully-NaN lines give NaN (np_compat.py:20-24)."""
    out = np.nansum(a, axis=axis)
    out[np.isnan(a).all(axis=axis)] = np.nan
    return out


def strategy_for(cube):
    """cube_utils.py:299-301."""
    return 'cube' if cube.size < MEMORY_THRESHOLD else 'slice'


def _planes(cube, axis, fill):
    """Yield (index-tuple, filled plane) along ``axis`` -- the loop at :46-53 / :79-85."""
    for i in range(cube.shape[axis]):
        sel = tuple(i if a == axis else slice(None) for a in range(3))
        yield sel, cube._get_filled_data(fill=fill, view=sel)


def _weighted_plane_sums(cube, axis, power, centre=0.0):
    """sum_i plane_i * (x_i - centre)**power * dx  and  sum_i plane_i * dx, fill=0."""
    dx = cube._pix_size_slice(axis)
    x = cube._pix_cen()[axis]
    shape = cube.shape[:axis] + cube.shape[axis + 1:]
    num = np.zeros(shape)
    den = np.zeros(shape)
    for sel, plane in _planes(cube, axis, 0):
        if power == 1 and np.isscalar(centre):
            num += plane * x[sel] * dx                      # :82-84
        else:
            num += plane * (x[sel] - centre) ** power * dx  # :119-121
        den += plane * dx
    return num, den


def moment_slicewise(cube, order, axis):
    with np.errstate(invalid='ignore', divide='ignore'):
        if order == 0:
            dx = cube._pix_size_slice(axis)
            shape = cube.shape[:axis] + cube.shape[axis + 1:]
            total = np.zeros(shape)
            seen = np.zeros(shape, dtype=bool)
            for _, plane in _planes(cube, axis, np.nan):
                seen |= np.isfinite(plane)
                total += np.nan_to_num(plane) * dx
            total[~seen] = np.nan
            return total
        num, den = _weighted_plane_sums(cube, axis, 1)
        first = num / den
        if order == 1:
            return first
        num, den = _weighted_plane_sums(cube, axis, order, centre=first)
        return num / den


def moment_raywise(cube, order, axis):
    shape = cube.shape[:axis] + cube.shape[axis + 1:]
    out = np.full(shape, np.nan)
    x = cube._pix_cen()[axis]
    dx = cube._pix_size_slice(axis)
    with np.errstate(invalid='ignore', divide='ignore'):
        for pos in np.ndindex(*shape):
            ray = list(pos)
            ray.insert(axis, slice(None))
            ray = tuple(ray)
            keep = cube._mask_include(view=ray)
            if not keep.any():
                continue
            w = cube._data[ray][keep] * dx                  # cube.flattened(slc).value * pix_size
            if order == 0:
                out[pos] = w.sum()
                continue
            xr = x[ray][keep]
            m1 = (w * xr).sum() / w.sum()
            out[pos] = m1 if order == 1 else (w * (xr - m1) ** order).sum() / w.sum()
    return out


def moment_cubewise(cube, order, axis):
    x = cube._pix_cen()[axis]
    w = cube._get_filled_data() * cube._pix_size_slice(axis)          # :176
    with np.errstate(invalid='ignore', divide='ignore'):
        if order == 0:
            return nansum_allbad_nan(w, axis)
        den = np.nansum(w, axis=axis)
        m1 = np.nansum(w * x, axis=axis) / den
        if order == 1:
            return m1
        return np.nansum(w * (x - np.expand_dims(m1, axis)) ** order, axis=axis) / den


def moment_auto(cube, order, axis):
    return DISPATCH[strategy_for(cube)](cube, order, axis)


def moment_dask(cube, order, axis):
    x = cube._pix_cen()[axis]
    dx = cube._pix_size_slice(axis)
    d = cube._get_filled_data(fill=np.nan).astype(np.float64)         # dask:1083
    with np.errstate(invalid='ignore', divide='ignore'):
        den = nansum_allbad_nan(d * dx, axis)
        if order == 0:
            return den
        m1 = nansum_allbad_nan(d * dx * x, axis) / den
        if order == 1:
            return m1
        return nansum_allbad_nan(d * dx * (x - np.expand_dims(m1, axis)) ** order, axis) / den


DISPATCH = dict(slice=moment_slicewise, cube=moment_cubewise, ray=moment_raywise, auto=moment_auto)
