"""
oracle/reproject.py -- TEST INFRASTRUCTURE.  Bilinear spatial reprojection.  PARITY UNPINNED.

The reference delegates to ``reproject.reproject_interp((data, header), wcs_out,
shape_out=..., order='bilinear')`` (``spectral_cube/spectral_cube.py:2726-2732``) and
turns the returned footprint into a ``BooleanArrayMask`` (:2741-2746).  ``reproject``
(>=0.9.1, ``pyproject.toml:41``) is not vendored and not installable here, and the
reference's own tests pin only shape and WCS for this call
(``spectral_cube/tests/test_regrid.py:99-135``).  So this file restates the package's
published algorithm from memory and says so: **parity for reproject is unpinned**.

Algorithm restated (reproject ``interpolation/core.py`` + ``array_utils.map_coordinates``):
  1. for every output pixel, map pixel -> world with the output WCS and world ->
     pixel with the input WCS (float64);
  2. sample with ``scipy.ndimage.map_coordinates(order=1, mode='constant',
     cval=nan)`` on a copy of the image padded by one edge-replicated pixel, so that
     samples up to half a pixel outside the outermost pixel centres are kept;
     anything further out is NaN;
  3. footprint = ``~isnan(result)``.
The spectral axis is treated as an identity map (same spectral WCS in and out, the
case ``SpectralCube.reproject`` documents: "If you want to reproject a cube both
spatially and spectrally, you need to use spectral_interpolate as well", :2656-2657),
i.e. each channel image is resampled independently.
The real ``scipy.ndimage.map_coordinates`` is called, so zero-weight NaN neighbours
poison the sample exactly as they do in the reference stack.
"""
import numpy as np
import scipy.ndimage


def input_pixel_coords(wcs_in, wcs_out, shape_out_yx):
    """(yin, xin) float64 planes: where each output pixel centre falls in the input image."""
    ny, nx = shape_out_yx
    yo, xo = np.meshgrid(np.arange(ny, dtype=np.float64), np.arange(nx, dtype=np.float64),
                         indexing='ij')
    lon, lat = wcs_out.celestial_pix2world(xo, yo, origin=0)
    xin, yin = wcs_in.celestial_world2pix(lon, lat, origin=0)
    return yin, xin


def sample_bilinear(image, yin, xin):
    ny, nx = image.shape
    padded = np.pad(image.astype(np.float64), 1, mode='edge')
    bad = ~np.isfinite(yin) | ~np.isfinite(xin)
    yc = np.where(bad, -10.0, yin)
    xc = np.where(bad, -10.0, xin)
    vals = scipy.ndimage.map_coordinates(padded, [yc + 1.0, xc + 1.0], order=1,
                                         mode='constant', cval=np.nan)
    outside = bad | (yc < -0.5) | (yc > ny - 0.5) | (xc < -0.5) | (xc > nx - 0.5)
    vals[outside] = np.nan
    return vals


def reproject_cube(data, wcs_in, wcs_out, shape_out):
    """data (nchan, ny_in, nx_in) -> (newdata float64 (nchan, ny_out, nx_out), footprint bool)."""
    nchan = data.shape[0]
    assert shape_out[0] == nchan
    yin, xin = input_pixel_coords(wcs_in, wcs_out, shape_out[1:])
    out = np.empty(shape_out, dtype=np.float64)
    for c in range(nchan):
        out[c] = sample_bilinear(data[c], yin, xin)
    return out, ~np.isnan(out)
