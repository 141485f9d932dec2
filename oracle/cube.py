"""
oracle/cube.py -- TEST INFRASTRUCTURE.  A small CPU cube object that strings the oracle
pieces together behind the reference's method names, so the parity tests read like
the reference's own tests.

Restates, for the hot path only:
  BaseSpectralCube.__init__           spectral_cube/spectral_cube.py:179-234
  _get_filled_data                    spectral_cube/base_class.py:389-417
  world / _pix_cen / _pix_size_slice  base_class.py:178-241, spectral_cube.py:1455-1535
  moment (+ units, + world[0] shift)  spectral_cube.py:1614-1720 ; dask :1031-1132
  linewidth_sigma / linewidth_fwhm    spectral_cube.py:1746-1763
  comparison operators, with_mask     spectral_cube.py:2263-2296, 1259-1306
  spectral_smooth / spatial_smooth    spectral_cube.py:3186-3222, 2808-2842,
                                      _apply_function_parallel_base :2900-3047,
                                      _apply_spectral/spatial_function :147-172 ;
                                      dask :880-917, :962-993, :501-552, :816-844
  spectral_interpolate                spectral_cube.py:3224-3332 ; dask :1250-1373
  reproject                           spectral_cube.py:2649-2746
``use_dask`` selects which of the two reference classes is mirrored where they differ
(SURVEY.md section 7 "Semantics that differ between the two reference classes").
Units are carried as plain strings; results are (ndarray, unit-string) pairs.
"""
import operator
import warnings
import numpy as np

from . import moments as _mom
from . import masks as _masks
from . import convolve as _conv
from . import interp as _interp
from . import reproject as _reproj
from .wcs import OWCS, unit_scale, angular_separation

SIGMA2FWHM = 2. * np.sqrt(2. * np.log(2.))        # spectral_cube.py:82


class VarianceWarning(UserWarning):
    pass


class SmoothingWarning(UserWarning):
    pass


class BeamUnitsError(Exception):
    pass


class OracleCube(object):
    def __init__(self, data, wcs, mask='isfinite', unit='K', fill_value=np.nan,
                 spectral_unit=None, use_dask=False, meta=None, beam=None, beams=None, goodbeams_mask=None):
        # `beam`: one OBeam (SpectralCube); `beams`: one per channel (VaryingResolutionSpectralCube)
        self.beam, self.beams = beam, beams
        self.goodbeams_mask = (np.array([b.isfinite for b in beams]) if beams is not None and goodbeams_mask is None
                               else goodbeams_mask)
        self._data = data
        self._wcs = wcs
        # io/fits.py:214 attaches LazyMask(np.isfinite) on read; direct construction may pass None
        if isinstance(mask, str) and mask == 'isfinite':
            mask = _masks.LazyMask(np.isfinite, data=data)
        self._mask = mask
        self.unit = unit
        self._fill_value = fill_value
        self._spectral_unit = spectral_unit if spectral_unit is not None else wcs.cunit[2]
        self._spectral_scale = unit_scale(wcs.cunit[2], self._spectral_unit)
        self.use_dask = use_dask
        self.meta = dict(meta or {})
        self._cache = {}

    # -- basic properties ------------------------------------------------------------------
    @property
    def shape(self):
        return self._data.shape

    @property
    def size(self):
        return self._data.size

    @property
    def mask(self):
        return self._mask

    @property
    def fill_value(self):
        return self._fill_value

    def _new_cube_with(self, **kw):
        args = dict(data=self._data, wcs=self._wcs, mask=self._mask, unit=self.unit,
                    fill_value=self._fill_value, spectral_unit=self._spectral_unit,
                    use_dask=self.use_dask, meta=self.meta, beam=self.beam, beams=self.beams,
                    goodbeams_mask=self.goodbeams_mask)
        args.update(kw)
        return OracleCube(**args)

    def with_mask(self, mask, inherit_mask=True):
        if isinstance(mask, np.ndarray):
            mask = _masks.BooleanArrayMask(mask, shape=self.shape)
        if self._mask is not None and inherit_mask:
            mask = self._mask & mask
        return self._new_cube_with(mask=mask)

    def with_fill_value(self, fill_value):
        return self._new_cube_with(fill_value=fill_value)

    def with_spectral_unit(self, unit):
        return self._new_cube_with(spectral_unit=unit)

    def _cmp(self, op, value):
        # spectral_cube.py:2237-2296: the threshold reaches numpy as an np.float64
        if not hasattr(value, 'shape'):
            value = np.float64(value)
        return _masks.LazyComparisonMask(op, value, data=self._data)

    def __gt__(self, v):
        return self._cmp(operator.gt, v)

    def __ge__(self, v):
        return self._cmp(operator.ge, v)

    def __lt__(self, v):
        return self._cmp(operator.lt, v)

    def __le__(self, v):
        return self._cmp(operator.le, v)

    def __eq__(self, v):
        return self._cmp(operator.eq, v)

    def __ne__(self, v):
        return self._cmp(operator.ne, v)

    def __hash__(self):
        return id(self)

    # -- masked data access ----------------------------------------------------------------
    def _get_filled_data(self, view=(), fill=np.nan):
        if self._mask is None:
            return self._data[view]
        return self._mask._filled(data=self._data, fill=fill, view=view)

    @property
    def unitless_filled_data(self):
        return self._get_filled_data(fill=self._fill_value)

    filled_data = unitless_filled_data

    def _mask_include(self, view=()):
        if self._mask is None:
            return np.ones(self._data[view].shape, dtype=bool)
        return self._mask.include(data=self._data, view=view)

    def flattened(self, view=()):
        return self._data[view][self._mask_include(view)]

    # -- coordinates -------------------------------------------------------------------------
    @property
    def spectral_axis(self):
        """Channel centres in ``_spectral_unit`` (spectral_cube.py:1765-1771)."""
        return self._wcs.spectral_pix2world(np.arange(self.shape[0])) * self._spectral_scale

    def world_plane0(self):
        """``self.world[0, :, :]`` -> (spectral, lat, lon) planes for channel 0."""
        ny, nx = self.shape[1:]
        yy, xx = np.meshgrid(np.arange(ny, dtype=float), np.arange(nx, dtype=float), indexing='ij')
        lon, lat = self._wcs.celestial_pix2world(xx, yy)
        spec = np.full((ny, nx), float(self._wcs.spectral_pix2world(0.0))) * self._spectral_scale
        return spec, lat, lon

    def _pix_cen(self):
        if 'pix_cen' in self._cache:
            return self._cache['pix_cen']
        _, lat, lon = self.world_plane0()
        spectral = self.spectral_axis.copy()
        spectral -= spectral[0]
        lon = np.radians(lon)
        lat = np.radians(lat)
        dx = angular_separation(lon[:, :-1], lat[:, :-1], lon[:, 1:], lat[:, :-1])
        dy = angular_separation(lon[:-1, :], lat[:-1, :], lon[1:, :], lat[1:, :])
        x = np.zeros(self.shape[1:])
        y = np.zeros(self.shape[1:])
        x[:, 1:] = np.cumsum(np.degrees(dx), axis=1)
        y[1:, :] = np.cumsum(np.degrees(dy), axis=0)
        x, y, spectral = np.broadcast_arrays(x[None, :, :], y[None, :, :], spectral[:, None, None])
        self._cache['pix_cen'] = (spectral, y, x)
        return self._cache['pix_cen']

    def _pix_size_slice(self, axis):
        psm = self._wcs.pixel_scale_matrix
        if axis == 0:
            return np.abs(psm[2, 2]) * self._spectral_scale
        elif axis in (1, 2):
            return np.sum(psm[2 - axis, :] ** 2) ** 0.5
        raise ValueError("Cubes have 3 axes.")

    # -- reductions (spectral_cube.py:361-470 `apply_numpy_function`, how='cube'; 578-826) -------------
    def _apply_numpy_function(self, function, fill=np.nan, axis=None):
        """spectral_cube.py:446-454: the function runs on the filled data of the whole cube."""
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            return function(self._get_filled_data(fill=fill), axis=axis)

    def sum(self, axis=None):
        if axis is None:                                                               # np_compat.allbadtonan, whole cube
            filled = self._get_filled_data(fill=np.nan)
            return np.float32(np.nan) if np.isnan(filled).all() else np.nansum(filled)
        return self._apply_numpy_function(_mom.nansum_allbad_nan, axis=axis)            # :578-588 (np_compat.allbadtonan)

    def mean(self, axis=None):
        return self._apply_numpy_function(np.nanmean, axis=axis)                       # :650-652

    def std(self, axis=None, ddof=0):
        return self._apply_numpy_function(lambda a, axis: np.nanstd(a, axis=axis, ddof=ddof), axis=axis)   # :721-724

    def max(self, axis=None):
        return self._apply_numpy_function(np.nanmax, axis=axis)                        # :770-781

    def min(self, axis=None):
        return self._apply_numpy_function(np.nanmin, axis=axis)                        # :785-796

    def argmax(self, axis=None):
        return self._apply_numpy_function(np.nanargmax, fill=-np.inf, axis=axis)       # :800-811

    def argmin(self, axis=None):
        return self._apply_numpy_function(np.nanargmin, fill=np.inf, axis=axis)        # :815-826

    def statistics(self):
        """dask_spectral_cube.py:769-814 (one chunk = the whole cube; the dtype of the data is kept like there)."""
        chunk = self._get_filled_data(fill=np.nan)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            stats = {'npts': np.sum(~np.isnan(chunk)), 'min': np.nanmin(chunk), 'max': np.nanmax(chunk),
                     'sum': np.nansum(chunk), 'sumsq': np.nansum(chunk * chunk)}
            stats['mean'] = stats['sum'] / stats['npts']
            stats['sigma'] = ((stats['sumsq'] - stats['sum'] ** 2 / stats['npts']) / (stats['npts'] - 1)) ** 0.5
            stats['rms'] = np.sqrt(stats['sumsq'] / stats['npts'])
        return stats

    # -- moments -------------------------------------------------------------------------------
    def moment(self, order=0, axis=0, how='auto'):
        if axis == 0 and order == 2:
            warnings.warn("Note that the second moment returned will be a "
                          "variance map. To get a linewidth map, use the "
                          "SpectralCube.linewidth_fwhm() or "
                          "SpectralCube.linewidth_sigma() methods instead.",
                          VarianceWarning)
        if self.use_dask:
            out = _mom.moment_dask(self, order, axis)
        else:
            if how not in _mom.DISPATCH:
                return ValueError("Invalid how. Must be in %s" % sorted(list(_mom.DISPATCH.keys())))
            out = _mom.DISPATCH[how](self, order, axis)
        axunit = self._spectral_unit if axis == 0 else 'deg'
        if order == 0:
            unit = '%s %s' % (self.unit, axunit)
        else:
            unit = axunit if max(order, 1) == 1 else '%s%d' % (axunit, max(order, 1))
        if order == 1 and axis == 0:
            out = out + self.world_plane0()[0]
        return out, unit

    def moment0(self, axis=0, how='auto'):
        return self.moment(order=0, axis=axis, how=how)

    def moment1(self, axis=0, how='auto'):
        return self.moment(order=1, axis=axis, how=how)

    def moment2(self, axis=0, how='auto'):
        return self.moment(order=2, axis=axis, how=how)

    def linewidth_sigma(self, how='auto'):
        with np.errstate(invalid='ignore'):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", VarianceWarning)
                m2, unit = self.moment2(how=how)
                return np.sqrt(m2), self._spectral_unit

    def linewidth_fwhm(self, how='auto'):
        s, unit = self.linewidth_sigma()
        return s * SIGMA2FWHM, unit

    # -- smoothing -----------------------------------------------------------------------------
    def spectral_smooth(self, kernel):
        nchan, ny, nx = self.shape
        if self.use_dask:
            # dask:816-844 (fill=NaN hard-coded), :912-914 one 3-d convolution, dtype kept
            data = self._get_filled_data(fill=np.nan)
            out = _conv.convolve(data, kernel.array.reshape(-1, 1, 1), normalize_kernel=True)
        else:
            data = self.unitless_filled_data
            out = np.empty(self.shape, dtype=float)           # :2953/:2963 float64 buffer
            for jj in range(ny):
                for ii in range(nx):
                    spec = data[:, jj, ii]
                    if np.any(self._mask_include((slice(None), jj, ii))):
                        out[:, jj, ii] = _conv.convolve(spec, kernel, normalize_kernel=True)
                    else:
                        out[:, jj, ii] = spec                  # :155-158
        return self._new_cube_with(data=out)                    # mask object unchanged :3043-3045

    def spatial_smooth(self, kernel, raise_error_jybm=True):
        if raise_error_jybm and self.unit.replace(' ', '').lower() in ('jy/beam',):
            raise BeamUnitsError("Attempting to change the spatial resolution of a cube with "
                                 "Jy/beam units. To ignore this error, set "
                                 "`raise_error_jybm=False`.")
        data = self.unitless_filled_data                        # fill = cube fill value (dask:552)
        if self.use_dask:
            out = np.zeros_like(data)                           # dask:540-547
            for c in range(self.shape[0]):
                out[c] = _conv.convolve(data[c], kernel.array, normalize_kernel=True)
        else:
            out = np.empty(self.shape, dtype=float)
            for c in range(self.shape[0]):
                img = data[c]
                if np.any(self._mask_include((c, slice(None), slice(None)))):
                    out[c] = _conv.convolve(img, kernel, normalize_kernel=True)
                else:
                    out[c] = img                                 # :169-172
        return self._new_cube_with(data=out)

    # -- convolution to a common beam -----------------------------------------------------------
    def _pixscale(self):
        """proj_plane_pixel_area(wcs.celestial)**0.5 in degrees (spectral_cube.py:3369)."""
        m = np.asarray(self._wcs.pixel_scale_matrix)[:2, :2]
        return float(np.sqrt(abs(np.linalg.det(m))))

    def mask_channels(self, goodchannels):
        goodchannels = np.asarray(goodchannels, dtype=bool)                  # :4289-4300
        cube = self.with_mask(np.broadcast_to(goodchannels[:, None, None], self.shape).copy())
        if self.beams is not None:
            cube.goodbeams_mask = np.logical_and(goodchannels, self.goodbeams_mask)
        return cube

    def convolve_to(self, beam, allow_smaller=False, convolve=None):
        """spectral_cube.py:3335-3392 (one beam; per image through ``_apply_spatial_function``) and :4127-4240
        (per-channel beams); dask_spectral_cube.py:1412-1464, 1512-1630.  ``convolve`` defaults to the class's
        default: ``convolve_fft`` for the numpy classes, ``convolve`` for the dask ones."""
        if convolve is None:
            convolve = _conv.convolve if self.use_dask else _conv.convolve_fft
        jybeam = self.unit.replace(' ', '').lower() == 'jy/beam'
        pixscale = self._pixscale()
        data = self.unitless_filled_data
        if self.beams is None:
            if beam == self.beam:
                warnings.warn("The given beam is identical to the current beam. Skipping convolution.")
                return self
            kernel = beam.deconvolve(self.beam).as_kernel(pixscale)
            factor = beam.sr / self.beam.sr if jybeam else 1.
            out = np.empty(self.shape, dtype=data.dtype if self.use_dask else float)
            for c in range(self.shape[0]):
                if self.use_dask or np.any(self._mask_include((c, slice(None), slice(None)))):
                    out[c] = convolve(data[c], kernel, normalize_kernel=True) * factor
                else:
                    out[c] = data[c]                                          # :169-172
            return self._new_cube_with(data=out, beam=beam)
        plan = []
        for bm, valid in zip(self.beams, self.goodbeams_mask):
            if not valid or beam == bm:
                plan.append((None, 1.))
                continue
            try:
                plan.append((beam.deconvolve(bm).as_kernel(pixscale), beam.sr / bm.sr if jybeam else 1.))
            except ValueError:
                if not allow_smaller:
                    raise
                plan.append((None, 1.))
        out = np.empty(self.shape, dtype=data.dtype if self.use_dask else float)
        for c, (kernel, factor) in enumerate(plan):
            out[c] = data[c] if kernel is None else convolve(data[c], kernel, normalize_kernel=True) * factor
        return self._new_cube_with(data=out, beam=beam, beams=None, goodbeams_mask=None)

    # -- resampling ----------------------------------------------------------------------------
    def spectral_interpolate(self, spectral_grid, suppress_smooth_warning=False, fill_value=None):
        """``spectral_grid`` is in ``_spectral_unit``."""
        inaxis = self.spectral_axis
        grid = np.asarray(spectral_grid, dtype=np.float64)
        indiff = abs(np.mean(np.diff(inaxis)))
        outdiff = abs(np.mean(np.diff(grid)))
        if outdiff > 2 * indiff and not suppress_smooth_warning:
            warnings.warn("Input grid has too small a spacing. The data should "
                          "be smoothed prior to resampling.", SmoothingWarning)
        if self.use_dask:
            newdata, newmask, rev = _interp.spectral_interpolate_dask(
                self._get_filled_data(fill=np.nan), inaxis, grid, fill_value)
        else:
            newdata, newmask, rev = _interp.spectral_interpolate_numpy(
                self.unitless_filled_data, self._mask_include(), inaxis, grid, fill_value)
        # new spectral WCS: crpix=1, crval=grid[0] (ascending order), cdelt=+-mean diff (:3317-3324)
        asc = grid[::-1] if rev else grid
        w = self._wcs.copy()
        scale = 1.0 / self._spectral_scale          # back to the WCS's SI unit
        w.crpix[2] = 1.0
        w.crval[2] = (asc[-1] if rev else asc[0]) * scale
        w.cdelt[2] = (-1.0 if rev else 1.0) * np.mean(np.diff(asc)) * scale
        w.pc[2, 2] = 1.0
        return self._new_cube_with(data=newdata, wcs=w,
                                   mask=_masks.BooleanArrayMask(newmask, shape=newmask.shape))

    def reproject(self, wcs_out, shape_out, filled=True):
        data = self.unitless_filled_data if filled else self._data
        newdata, valid = _reproj.reproject_cube(data, self._wcs, wcs_out, tuple(shape_out))
        if np.all(np.isnan(newdata)):
            raise ValueError("All values in reprojected cube are nan.")
        return self._new_cube_with(data=newdata, wcs=wcs_out,
                                   mask=_masks.BooleanArrayMask(valid, shape=valid.shape))
