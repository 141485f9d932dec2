"""
``SpectralCube`` / ``DaskSpectralCube`` with the reference's method surface for the hot path
(spectral_cube/spectral_cube.py, spectral_cube/dask_spectral_cube.py), backed by
libsc_b200 (hand-written CUDA for sm_100a, called through the C ABI in include/sc_b200.h).

The cube's float32 voxels live in HBM (a torch tensor is used purely as the memory holder);
masks are lazy expression trees evaluated inside the kernels; host work per call is
O(nchan).  Where the two reference classes differ (output dtype of smoothing, behaviour
outside the input range when interpolating, meta keys) ``SpectralCube`` mirrors the numpy
class and ``DaskSpectralCube`` the dask class.  Nothing here falls back to the CPU.
"""
import operator
import warnings

import numpy as np

from . import _lib
from . import masks as _masks
from .masks import (BooleanArrayMask, LazyMask, LazyComparisonMask, MaskBase, lower_mask)
from .projection import Projection, _unit_mul, _unit_pow
from .wcs import as_cube_wcs, spectral_unit_scale
from .beam import Beam, Beams, BeamError, NoBeamError

SIGMA2FWHM = 2. * np.sqrt(2. * np.log(2.))        # spectral_cube.py:82
MEMORY_THRESHOLD = 1e8                            # cube_utils.py:266-268
BIGDATAURL = "https://spectral-cube.readthedocs.io/en/latest/big_data.html"      # utils.py:11


class SpectralCubeWarning(Warning):
    pass


class VarianceWarning(SpectralCubeWarning):      # utils.py
    pass


class SmoothingWarning(SpectralCubeWarning):
    pass


class BeamUnitsError(Exception):
    pass


class UnitsError(ValueError):
    """Stands in for astropy.units.UnitsError when astropy is absent."""


def _torch():
    return _lib.require_cuda()


def _stream():
    return _torch().cuda.current_stream().cuda_stream


# ---- whole-cube statistics from the per-spaxel partials of one device pass (device-agnostic torch code:
#      tests/test_reduce_host.py runs it on CPU tensors) ------------------------------------------------------
def whole_sum(torch, sum_map, count_map):
    return float(torch.nansum(sum_map).item()) if int(count_map.sum().item()) > 0 else float('nan')


def whole_mean(torch, sum_map, count_map):
    n = int(count_map.sum().item())
    return float(torch.nansum(sum_map).item()) / n if n > 0 else float('nan')


def whole_std(torch, sum_map, count_map, m2_map, ddof=0):
    """Chan et al. combination: M2 = sum m2_i + sum n_i (mean_i - mean)^2."""
    n_i = count_map.to(torch.float64)
    n = float(n_i.sum().item())
    if n - ddof <= 0:
        return float('nan')
    ok = n_i > 0
    s_i = torch.where(ok, sum_map, torch.zeros_like(sum_map))
    mean = float(s_i.sum().item()) / n
    mean_i = s_i / torch.clamp(n_i, min=1.0)
    m2 = torch.where(ok, m2_map + n_i * (mean_i - mean) ** 2, torch.zeros_like(n_i)).sum().item()
    return float(np.sqrt(m2 / (n - ddof)))


def whole_statistics(torch, sum_map, count_map, m2_map, min_map, max_map):
    """The dictionary of ``DaskSpectralCube.statistics`` (dask_spectral_cube.py:769-814, CASA's ia.statistics names)
    from the per-spaxel partials.  ``sumsq`` is rebuilt without cancellation as sum_i (m2_i + s_i^2 / n_i); ``sigma``
    follows the reference's textbook formula on these sums."""
    n_i = count_map.to(torch.float64)
    ok = n_i > 0
    zero = torch.zeros_like(n_i)
    npts = int(count_map.sum().item())
    s_i = torch.where(ok, sum_map, zero)
    total = float(s_i.sum().item())
    sumsq = float(torch.where(ok, m2_map + s_i * s_i / torch.clamp(n_i, min=1.0), zero).sum().item())
    stats = {'npts': npts, 'min': whole_extremum(torch, min_map, 'min'), 'max': whole_extremum(torch, max_map, 'max'),
             'sum': total, 'sumsq': sumsq}
    with np.errstate(invalid='ignore', divide='ignore'):
        stats['mean'] = float(np.float64(total) / npts)
        stats['sigma'] = float(np.sqrt((np.float64(sumsq) - np.float64(total) ** 2 / npts) / (npts - 1)))
        stats['rms'] = float(np.sqrt(np.float64(sumsq) / npts))
    return stats


def whole_extremum(torch, ext_map, which):
    ok = ~torch.isnan(ext_map)
    if not bool(ok.any()):
        return float('nan')
    return float((ext_map[ok].max() if which == 'max' else ext_map[ok].min()).item())


def whole_arg_extremum(torch, ext_map, idx_map, which):
    """Flat C-order index of the FIRST extremum like np.nanarg*: the smallest (channel, spaxel) among the ties;
    0 ("arbitrary" in the reference) when nothing takes part."""
    ok = ~torch.isnan(ext_map)
    if not bool(ok.any()):
        return 0
    best = ext_map[ok].max() if which == 'max' else ext_map[ok].min()
    ny, nx = ext_map.shape
    spaxel = torch.arange(ny * nx, device=ext_map.device, dtype=torch.int64).view(ny, nx)
    flat = idx_map.to(torch.int64) * (ny * nx) + spaxel
    return int(flat[ok & (ext_map == best)].min().item())


class BaseSpectralCube(object):
    _mirrors_dask = False

    def __init__(self, data, wcs, mask=None, meta=None, fill_value=np.nan, header=None,
                 allow_huge_operations=False, unit=None, spectral_unit=None, device=None, **kwargs):
        torch = _torch()
        deferred = data if isinstance(data, _PendingFloat32Copy) else None
        if deferred is not None:
            t = deferred.hi                          # (only its shape is looked at below)
        elif isinstance(data, torch.Tensor):
            t = data
            if t.dtype != torch.float32:
                t = t.to(torch.float32)
            if not t.is_cuda:
                t = t.cuda(device)
        else:
            arr = np.asarray(data)
            if arr.ndim != 3:
                raise ValueError("data should be a 3-d array")
            t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).cuda(device)
        if t.dim() != 3:
            raise ValueError("data should be a 3-d array")
        if t.stride(2) != 1:
            t = t.contiguous()
        # a length-1 axis may carry any stride (numpy's `a[None]` gives 0): give it the one the kernels expect
        sy = t.stride(1) if t.shape[1] > 1 and t.stride(1) >= t.shape[2] else t.shape[2]
        sc = t.stride(0) if t.shape[0] > 1 and t.stride(0) >= t.shape[1] * sy else t.shape[1] * sy
        if (t.shape[1] == 1 and t.stride(1) != sy) or (t.shape[0] == 1 and t.stride(0) != sc):
            t = t.as_strided(tuple(t.shape), (sc, sy, 1), t.storage_offset())
        self._data_t = t if deferred is None else None
        self._hi = None if deferred is None else deferred.hi    # float64 tensor when the numpy-class semantics produce one ...
        self._hi_is_widened_f32 = False # ... or a flag: the float64 the reference returns is the float32 copy widened
        self._pending = deferred        # lazy op recorded by DaskSpectralCube (see spectral_smooth) / float32 copy not made yet
        self._wcs = as_cube_wcs(wcs)
        if mask is not None and not isinstance(mask, MaskBase):
            mask = BooleanArrayMask(np.asarray(mask, dtype=bool), self._wcs, shape=tuple(t.shape))
        self._mask = mask
        self._meta = dict(meta or {})
        self._header = dict(header or {})
        if unit is None:
            unit = self._meta.get('BUNIT', self._header.get('BUNIT', ''))
        self._unit = unit
        self._fill_value = fill_value
        self.allow_huge_operations = allow_huge_operations
        self._spectral_unit = spectral_unit if spectral_unit is not None else self._wcs.cunit[2]
        self._spectral_scale = spectral_unit_scale(self._wcs.cunit[2], self._spectral_unit)
        self._workspace = None
        self._beam = None
        self._passthrough_flags = None

    @property
    def _data_hi(self):
        """The float64 data the numpy class of the reference would hold, or None when it holds float32."""
        if self._hi is not None:
            return self._hi
        if self._hi_is_widened_f32:
            return self._data.to(_torch().float64)
        return None

    @_data_hi.setter
    def _data_hi(self, value):
        self._hi = value

    @property
    def _data(self):
        """float32 device tensor of the voxels; a pending lazy op is materialised on first use."""
        if self._data_t is None:
            self._data_t = self._pending.materialize()
        return self._data_t

    @property
    def _shape(self):
        return tuple(self._data_t.shape) if self._data_t is not None else self._pending.shape

    # -- construction ------------------------------------------------------------------------
    def _new_cube_with(self, data=None, wcs=None, mask=None, meta=None, fill_value=None,
                       spectral_unit=None, unit=None, cls=None, **kw):
        """spectral_cube.py:244-289"""
        cls = cls or type(self)
        cube = cls.__new__(cls)
        still_lazy = data is None and self._data_t is None and isinstance(self._pending, _PendingSpectralSmooth)
        if still_lazy:
            data = self._pending.source._data            # placeholder; replaced below
        BaseSpectralCube.__init__(
            cube,
            data=self._data if data is None else data,
            wcs=self._wcs if wcs is None else wcs,
            mask=self._mask if mask is None else mask,
            meta=self._meta if meta is None else meta,
            fill_value=self._fill_value if fill_value is None else fill_value,
            header=self._header, allow_huge_operations=self.allow_huge_operations,
            unit=self._unit if unit is None else unit,
            spectral_unit=self._spectral_unit if spectral_unit is None else spectral_unit)
        if still_lazy:
            cube._data_t, cube._pending = None, self._pending
        beam = kw.get('beam', None)
        if beam is None:
            beam = getattr(self, '_beam', None)               # spectral_cube.py:3733-3738
        cube._beam = None
        if beam is not None and hasattr(cube, '_attach_beam'):
            cube._meta = dict(cube._meta)
            cube._header = dict(cube._header)
            cube._attach_beam(beam)
        return cube

    # -- basic properties ------------------------------------------------------------------------
    @property
    def shape(self):
        return self._shape

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def ndim(self):
        return 3

    @property
    def _is_huge(self):
        """cube_utils.py:270-274"""
        return not (self.size < MEMORY_THRESHOLD)

    def _refuse_huge(self, name, how=None, accepts_how=False):
        """The `warn_slow` contract (utils.py:41-75): operations the reference can only do with the whole cube in host
        memory are refused for cubes of >= 1e8 voxels unless ``allow_huge_operations`` is set; the text is the
        reference's (tests/test_spectral_cube.py:104-133 match on it).  On the device nothing is loaded, but a script
        written against the reference relies on the refusal as much as on the result."""
        if how in ('slice', 'ray') or not self._is_huge or self.allow_huge_operations:
            return
        msg = ("This function ({0}) requires loading the entire "
               "cube into memory, and the cube is large ({1} "
               "pixels), so by default we disable this operation. "
               "To enable the operation, set "
               "`cube.allow_huge_operations=True` and try again.  ").format(
                   "<function %s.%s>" % (type(self).__name__, name), self.size)
        if accepts_how:
            msg += ("Alternatively, you may want to consider using an "
                    "approach that does not load the whole cube into "
                    "memory by specifying how='slice' or how='ray'.  ")
        msg += "See {bigdataurl} for details.".format(bigdataurl=BIGDATAURL)
        raise ValueError(msg)

    @property
    def unit(self):
        return self._unit

    @property
    def wcs(self):
        return self._wcs

    @property
    def mask(self):
        return self._mask

    @property
    def meta(self):
        return self._meta

    @property
    def fill_value(self):
        return self._fill_value

    @property
    def header(self):
        hdr = dict(self._header)
        hdr.update(self._wcs.to_header())
        hdr['BUNIT'] = str(self._unit)
        hdr['NAXIS'] = 3
        hdr['NAXIS1'], hdr['NAXIS2'], hdr['NAXIS3'] = self.shape[2], self.shape[1], self.shape[0]
        return hdr

    @property
    def spectral_axis(self):
        """Channel centres in the cube's spectral unit (spectral_cube.py:1765-1771)."""
        return self._wcs.spectral_pix2world(np.arange(self.shape[0])) * self._spectral_scale

    @property
    def device_data(self):
        """The float32 device tensor holding the voxels (memory holder only)."""
        return self._data

    def with_mask(self, mask, inherit_mask=True, wcs_tolerance=None):
        """spectral_cube.py:1259-1306"""
        if isinstance(mask, np.ndarray):
            if not _masks.is_broadcastable_and_smaller(self.shape, mask.shape):
                raise ValueError("Mask shape is not broadcastable to data shape: "
                                 "%s vs %s" % (mask.shape, self.shape))
            mask = BooleanArrayMask(mask, self._wcs, shape=self.shape)
        if self._mask is not None and inherit_mask:
            new_mask = self._mask & mask
        else:
            new_mask = mask
        cube = self._new_cube_with()
        cube._mask = new_mask
        return cube

    def with_fill_value(self, fill_value):
        return self._new_cube_with(fill_value=fill_value)

    def mask_channels(self, goodchannels):
        """Mask out whole channels with a 1-D boolean array (spectral_cube.py:3394-3418)."""
        goodchannels = np.asarray(goodchannels, dtype='bool')
        if goodchannels.ndim != 1:
            raise ValueError("goodchannels mask must be one-dimensional")
        if goodchannels.size != self.shape[0]:
            raise ValueError("goodchannels must have a length equal to the "
                             "cube's spectral dimension.")
        return self.with_mask(goodchannels[:, None, None])

    def with_spectral_unit(self, unit, **kwargs):
        return self._new_cube_with(spectral_unit=unit)

    # -- comparison operators -> lazy masks (spectral_cube.py:2263-2296) ---------------------------
    def _val_to_own_unit(self, value):
        if hasattr(value, 'unit') and hasattr(value, 'to') and not isinstance(self._unit, str):
            return value.to(self._unit).value
        if hasattr(value, 'value') and hasattr(value, 'unit'):
            return value.value
        return value

    def _cmp(self, op, value):
        value = self._val_to_own_unit(value)
        return LazyComparisonMask(op, value, data=self._data, wcs=self._wcs)

    def __gt__(self, value):
        return self._cmp(operator.gt, value)

    def __ge__(self, value):
        return self._cmp(operator.ge, value)

    def __lt__(self, value):
        return self._cmp(operator.lt, value)

    def __le__(self, value):
        return self._cmp(operator.le, value)

    def __eq__(self, value):
        return self._cmp(operator.eq, value)

    def __ne__(self, value):
        return self._cmp(operator.ne, value)

    def __hash__(self):
        return id(self)

    # -- masked data access (base_class.py:389-450) ------------------------------------------------
    def _mask_desc(self):
        return lower_mask(self._mask, self._data)

    def _filled_tensor(self, fill=np.nan):
        torch = _torch()
        lib = _lib.load()
        if self._mask is None:
            return self._data
        if getattr(self, '_nan_filled_already', False) and fill != fill:
            # the data of a freshly interpolated cube are NaN wherever its new mask excludes them (both bracketing samples
            # were masked: spectral_cube.py:3305-3313, dask :1364): filling with NaN changes nothing, no pass needed
            return self._data
        desc, keep = self._mask_desc()
        out = torch.empty(self.shape, dtype=torch.float32, device=self._data.device)
        nchan, ny, nx = self.shape
        _lib.check(lib.sc_fill_masked(self._data.data_ptr(), nchan, ny, nx, self._data.stride(0),
                                      self._data.stride(1), desc, float(fill), out.data_ptr(), _stream()))
        return out

    def _get_filled_data(self, view=(), fill=np.nan, **kwargs):
        if self._data_hi is not None:
            torch = _torch()
            if self._mask is None:
                return self._data_hi.cpu().numpy()[view]
            if isinstance(self._mask, BooleanArrayMask) and self._data_t is None:
                # a freshly reprojected cube: the mask is the footprint array itself, no need for the float32 copy
                inc = self._mask._mask.bool() if self._mask._mask_type == 'include' else ~self._mask._mask.bool()
                inc = inc.expand(self.shape) if tuple(inc.shape) != tuple(self.shape) else inc
            else:
                inc = self._mask._include_tensor(self._data).bool()
            return torch.where(inc, self._data_hi, torch.full((), float(fill), dtype=torch.float64,
                                                              device=self._data_hi.device)).cpu().numpy()[view]
        return self._filled_tensor(fill).cpu().numpy()[view]

    @property
    def filled_data(self):
        return _Sliceable(lambda view: self._get_filled_data(view=view, fill=self._fill_value))

    @property
    def unitless_filled_data(self):
        return self.filled_data

    @property
    def unmasked_data(self):
        return _Sliceable(lambda view: (self._data_hi if self._data_hi is not None else self._data).cpu().numpy()[view])

    def flattened(self, slice=(), weights=None):
        return self._mask._flattened(self._data, view=slice) if self._mask is not None \
            else self._data.cpu().numpy()[slice].ravel()

    # -- coordinates (O(nchan) on the host) ----------------------------------------------------------
    def _spectral_offsets(self):
        """``_pix_cen()[0]`` reduced to its nchan-long content (spectral_cube.py:1473-1475)."""
        spectral = self.spectral_axis.copy()
        spectral -= spectral[0]
        return spectral

    def _pix_size_slice(self, axis):
        """spectral_cube.py:1510-1535"""
        psm = self._wcs.pixel_scale_matrix
        if axis == 0:
            return np.abs(psm[2, 2]) * self._spectral_scale
        elif axis in (1, 2):
            return np.sum(psm[2 - axis, :] ** 2) ** 0.5
        raise ValueError("Cubes have 3 axes.")

    def _world0_spectral(self):
        """Spectral part of ``self.world[0, :, :]`` -- a constant (base_class.py:236-239)."""
        return float(self._wcs.spectral_pix2world(0.0)) * self._spectral_scale

    def _get_workspace(self, nbytes):
        torch = _torch()
        if self._workspace is None or self._workspace.numel() < nbytes:
            dev = self._data_t.device if self._data_t is not None else self._pending.source._data.device
            self._workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        return self._workspace

    # -- views (spectral_cube.py:1290-1380 `__getitem__`, :1822-1876 `spectral_slab`) --------------------
    def __getitem__(self, view):
        """Sub-cube for a view made of slices (a VIEW of the same device memory, like the reference's);
        the WCS follows `wcs_utils.slice_wcs`, the mask is sliced alongside.  Integer indices (which give
        spectra and slices in the reference) are not part of the device path."""
        from .masks import _slice3
        view = _slice3(view)
        if any(v.step is not None and v.step < 1 for v in view):
            raise NotImplementedError("negative steps are not supported")
        data = self._data[view]
        if 0 in data.shape:
            raise ValueError("the view selects no voxel")
        wcs = self._wcs.copy()
        for np_axis, sl in enumerate(view):
            w = 2 - np_axis                                       # numpy axis -> FITS axis index
            start, stop, step = sl.indices(self.shape[np_axis])
            # wcs_utils.slice_wcs: crpix follows the first kept pixel; a step stretches cdelt about its centre
            wcs.crpix[w] = (wcs.crpix[w] - start - 0.5) / step + 0.5
            wcs.cdelt[w] = wcs.cdelt[w] * step
        mask = self._mask[view] if self._mask is not None else None
        cube = self._new_cube_with(data=data, wcs=wcs, mask=mask)
        if self._mask is None:
            cube._mask = None
        return cube

    def closest_spectral_channel(self, value):
        """spectral_cube.py:1800-1820 (value in the cube's spectral unit)."""
        value = float(getattr(value, 'value', value))
        return int(np.argmin(np.abs(np.asarray(self.spectral_axis, dtype=np.float64) - value)))

    def spectral_slab(self, lo, hi):
        """Extract a new cube between two spectral coordinates (spectral_cube.py:1822-1876): a view."""
        ilo, ihi = self.closest_spectral_channel(lo), self.closest_spectral_channel(hi)
        if ilo == ihi:
            warnings.warn("The maxmimum and minimum spectral channel in the spectral"
                          "slab are identical; this indicates that one or both are "
                          "likely incorrect and/or out of range.", SpectralCubeWarning)
        if ilo > ihi:
            ilo, ihi = ihi, ilo
        return self[ilo:ihi + 1, :, :]

    # -- smoothing (spectral_cube.py:3186-3222, 2808-2842; dask_spectral_cube.py:880-917, 962-993) ----
    @staticmethod
    def _kernel_array(kernel, ndim):
        arr = kernel.array if hasattr(kernel, 'array') else kernel
        if hasattr(arr, 'unit'):
            # spectral_cube.py:3212-3214
            raise UnitsError("The convolution kernel should be defined without a unit.")
        arr = np.asarray(arr, dtype=np.float64)
        if arr.ndim != ndim:
            raise Exception("array and kernel have differing number of dimensions.")
        if any(n % 2 == 0 for n in arr.shape):
            raise Exception("Kernel size must be odd in all axes.")
        return arr

    def _smooth_fill(self):
        """Fill value masked voxels take before convolving (numpy class: the cube's fill value,
        spectral_cube.py:3085/3139; the dask spectral path hard-codes NaN, dask:816-823)."""
        return self._fill_value

    def _run_spectral_smooth(self, taps, out_dtype, out=None):
        """`out` may be the cube's own float32 tensor: the kernel then smooths in place (a 137 GB
        cube has no room for a second copy in 180 GB of HBM)."""
        torch = _torch()
        lib = _lib.load()
        src = self._data
        nchan, ny, nx = self.shape
        if out is None:
            out = torch.empty((nchan, ny, nx), dtype=torch.float64 if out_dtype == _lib.F64 else torch.float32,
                              device=src.device)
        desc, keep = self._mask_desc()
        ws = self._get_workspace(lib.sc_workspace_bytes(_lib.OP_SPECTRAL_SMOOTH, nchan, ny, nx, len(taps)))
        tarr, tptr = _lib.as_double_array(taps)
        _lib.check(lib.sc_spectral_smooth(
            src.data_ptr(), out.data_ptr(), out_dtype, nchan, ny, nx, src.stride(0), src.stride(1),
            out.stride(0), out.stride(1), desc, float(self._smooth_fill()), tptr, len(taps),
            0 if self._mirrors_dask else 1,      # numpy class: _apply_spectral_function pass-through (:147-158)
            ws.data_ptr(), ws.numel(), _stream()))
        return out

    def _new_cube_reporting_f64(self, out32):
        """The numpy class stores the convolution in a float64 buffer (spectral_cube.py:2953/:2963), but astropy's
        `convolve` has already cast every value back to the input's float32 -- the float64 cube IS the float32
        result widened.  It is therefore kept as float32 (8 instead of 12 + 12 B/voxel of traffic and a third of
        the memory) and widened on demand."""
        cube = self._new_cube_with(data=out32)
        cube._hi_is_widened_f32 = True
        return cube

    def spectral_smooth(self, kernel, convolve=None, verbose=0, use_memmap=True, num_cores=None, **kwargs):
        """Smooth the cube along the spectral dimension; the mask is left unchanged.  ``convolve`` may be astropy's
        ``convolve`` (the default) or ``convolve_fft``; any other callable is refused -- the convolution runs on the
        device and a user function cannot be called there (spectral_cube.py:3188, 3216-3222)."""
        self._check_convolve_kwargs(kwargs)
        fft = self._fft_semantics(convolve, default=False)
        taps = self._kernel_array(kernel, 1)
        out = self._run_spectral_smooth(taps, _lib.F32)
        self._convolve_epilogue(out, 1.0, fft)
        return self._new_cube_reporting_f64(out)

    def check_jybeam_smoothing(self, raise_error_jybm=True):
        """base_class.py:116-140"""
        if str(self._unit).replace(' ', '').lower() == 'jy/beam' and raise_error_jybm:
            raise BeamUnitsError("Attempting to change the spatial resolution of a cube with Jy/beam units."
                                 " To ignore this error, set `raise_error_jybm=False`.")

    @staticmethod
    def _separable_factors(k2d, every_tap=False):
        """(ky, kx) if the 2-D kernel is an outer product to float64 rounding, else None.  ``every_tap``: the
        product has to reproduce each tap to 1e-12 of ITSELF, not of the largest one -- beam kernels are 16 sigma
        wide and outputs deep inside blank regions hang on taps of 1e-20 and less (a rotated sub-pixel ellipse is an
        outer product to 1e-14 of its peak and nothing like one out there)."""
        cy, cx = k2d.shape[0] // 2, k2d.shape[1] // 2
        piv = k2d[cy, cx]
        if piv == 0:
            return None
        ky, kx = k2d[:, cx].copy(), k2d[cy, :] / piv
        err = np.abs(np.outer(ky, kx) - k2d)
        if (np.all(err <= 1e-12 * np.abs(k2d)) if every_tap else np.max(err) <= 1e-14 * np.max(np.abs(k2d))):
            return ky, kx
        return None

    def _spatial_strategy_counts(self):
        """Device uint32[2] {crowded blocks, blocks sampled} of this cube's rows: what the separable spatial
        kernels use to pick their denominator strategy.  Row-sharded jobs sum it over the ranks (distributed.py)."""
        torch = _torch()
        lib = _lib.load()
        src = self._data
        counts = torch.zeros((2,), dtype=torch.int32, device=src.device)
        desc, keep = self._mask_desc()
        _lib.check(lib.sc_spatial_missing_sample(src.data_ptr(), *self.shape, src.stride(0), src.stride(1), desc,
                                                 counts.data_ptr(), _stream()))
        return counts

    def _run_spatial_smooth(self, k2d, out_dtype, halo_top=None, halo_bot=None, halo_rows=0, strategy_counts=None,
                            out=None, passthrough=None, every_tap=False):
        """`out` may be a (nchan, ny, nx) view of a larger tensor (x stride 1): `convolve_to` on per-channel
        beams writes each channel plane of one result cube with its own kernel."""
        torch = _torch()
        lib = _lib.load()
        src = self._data
        nchan, ny, nx = self.shape
        if out is None:
            out = torch.empty((nchan, ny, nx), dtype=torch.float64 if out_dtype == _lib.F64 else torch.float32,
                              device=src.device)
        desc, keep = self._mask_desc()
        ws = self._get_workspace(lib.sc_workspace_bytes(_lib.OP_SPATIAL_SMOOTH, nchan, ny, nx, k2d.size))
        ht = halo_top.data_ptr() if halo_top is not None else None
        hb = halo_bot.data_ptr() if halo_bot is not None else None
        common = (src.data_ptr(), out.data_ptr(), out_dtype, nchan, ny, nx, src.stride(0), src.stride(1),
                  out.stride(0), out.stride(1), desc, float(self._fill_value))
        if passthrough is None:
            passthrough = 0 if self._mirrors_dask else 1     # _apply_spatial_function (:161-172)
        # the library leaves the per-channel "copied through" flags in the workspace behind the taps
        self._passthrough_flags = (ws[k2d.size * 8 + 512: k2d.size * 8 + 512 + nchan]
                                   if passthrough and desc.n_nodes > 0 else None)
        sep = self._separable_factors(k2d, every_tap=every_tap)
        if sep is not None:
            (ya, yp), (xa, xp) = _lib.as_double_array(sep[0]), _lib.as_double_array(sep[1])
            sc = strategy_counts.data_ptr() if strategy_counts is not None else None
            _lib.check(lib.sc_spatial_smooth_sep_ex(*common, yp, len(ya), xp, len(xa), ht, hb, int(halo_rows), passthrough,
                                                    sc, ws.data_ptr(), ws.numel(), _stream()))
        else:
            ka, kp = _lib.as_double_array(k2d.ravel())
            _lib.check(lib.sc_spatial_smooth_2d(*common, kp, k2d.shape[0], k2d.shape[1], ht, hb, int(halo_rows), passthrough,
                                                ws.data_ptr(), ws.numel(), _stream()))
        return out

    def spatial_smooth(self, kernel, convolve=None, raise_error_jybm=True, **kwargs):
        """Smooth the image in each spatial-spatial plane of the cube (``convolve``: see ``spectral_smooth``)."""
        self.check_jybeam_smoothing(raise_error_jybm=raise_error_jybm)
        self._check_convolve_kwargs(kwargs)
        fft = self._fft_semantics(convolve, default=False)
        k2d = self._kernel_array(kernel, 2)
        out = self._run_spatial_smooth(k2d, _lib.F32)
        # planes copied through (:169-172) never met the convolution function
        self._convolve_epilogue(out, 1.0, fft, skip=self._passthrough_flags)
        if self._mirrors_dask:
            return self._new_cube_with(data=out)
        return self._new_cube_reporting_f64(out)

    # -- user-function seams (spectral_cube.py:3049-3159; dask_spectral_cube.py:501-638) ------------------------
    def _apply_function_on_device(self, function, what, **kwargs):
        """The reference maps a numpy callable over spectra / images on host cores.  Here the cube lives in HBM, so the
        callable is handed the WHOLE filled cube once, as a float32 CUDA tensor of shape (nchan, ny, nx) -- the dask
        class's ``accepts_chunks=True`` contract (dask_spectral_cube.py:516-520, 569-574) with the device as the one
        chunk -- and must return a tensor of the same shape on the same device.  A function that needs numpy arrays
        cannot run on this path and is refused with that explanation rather than silently pulled to the host."""
        torch = _torch()
        for k in ('num_cores', 'verbose', 'use_memmap', 'parallel', 'accepts_chunks', 'return_new_cube', 'save_to_tmp_dir',
                  'update_function', 'memmap_dir'):
            kwargs.pop(k, None)
        data = self._filled_tensor(self._fill_value)
        try:
            out = function(data, **kwargs)
        except TypeError as exc:
            raise NotImplementedError(
                "apply_function_parallel_%s: `function` is called once with the filled cube as a float32 CUDA tensor "
                "(nchan, ny, nx) and must return a tensor of that shape (no CPU fallback): %r" % (what, exc))
        if not isinstance(out, torch.Tensor) or tuple(out.shape) != tuple(data.shape) or out.device != data.device:
            raise ValueError("apply_function_parallel_%s: `function` must return a CUDA tensor of shape %s on %s"
                             % (what, tuple(data.shape), data.device))
        return self._new_cube_with(data=out.to(torch.float32))

    def apply_function_parallel_spectral(self, function, **kwargs):
        """Apply ``function`` along the spectral dimension of the filled data; the mask is left unchanged
        (spectral_cube.py:3103-3159, :3043-3045).  See ``_apply_function_on_device`` for the callable's contract."""
        return self._apply_function_on_device(function, 'spectral', **kwargs)

    def apply_function_parallel_spatial(self, function, **kwargs):
        """Apply ``function`` to the channel images of the filled data (spectral_cube.py:3049-3101)."""
        return self._apply_function_on_device(function, 'spatial', **kwargs)

    # -- convolution to a common beam (spectral_cube.py:3335-3392; dask_spectral_cube.py:1412-1464) ------
    @property
    def beam(self):
        """base_class.py:830-838"""
        if getattr(self, '_beam', None) is None:
            raise NoBeamError("No beam is defined for this SpectralCube or the"
                              " beam information could not be parsed from the"
                              " header. A `~radio_beam.Beam` object can be"
                              " added using `cube.with_beam`.")
        return self._beam

    @beam.setter
    def beam(self, obj):
        self._beam = None if obj is None else Beam.coerce(obj)

    def with_beam(self, beam, raise_error_jybm=True):
        """Attach a beam object (spectral_cube.py:3741-3765)."""
        beam = Beam.coerce(beam)
        self.check_jybeam_smoothing(raise_error_jybm=raise_error_jybm)
        return self._new_cube_with(beam=beam)

    def _attach_beam(self, beam):
        self._beam = beam
        if beam is not None:
            self._meta['beam'] = beam
            self._header.update(beam.to_header_keywords())

    def _is_jybeam(self):
        return str(self._unit).replace(' ', '').lower() == 'jy/beam'

    def _pixscale_deg(self):
        """``proj_plane_pixel_area(self.wcs.celestial)**0.5`` in degrees (:3369)."""
        m = self._wcs.pixel_scale_matrix[:2, :2]
        return float(np.sqrt(abs(m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0])))

    @staticmethod
    def _check_convolve_kwargs(kwargs):
        """The device convolution implements astropy's defaults (zero-filled boundary, NaNs interpolated, kernel
        normalised), under which ``convolve`` and ``convolve_fft`` agree; anything else is refused, not ignored."""
        supported = {'nan_treatment': 'interpolate', 'boundary': 'fill', 'fill_value': 0.0, 'preserve_nan': False,
                     'normalize_kernel': True}
        for key, val in kwargs.items():
            if key in supported:
                if val != supported[key]:
                    raise NotImplementedError("convolution keyword %s=%r is not available on the device path "
                                              "(only %r)" % (key, val, supported[key]))
            elif key not in ('allow_huge', 'num_cores', 'use_memmap', 'parallel', 'verbose', 'psf_pad', 'fft_pad',
                             'save_to_tmp_dir', 'update_function', 'memmap_dir'):
                raise TypeError("unexpected keyword argument %r" % key)

    @staticmethod
    def _fft_semantics(convolve, default):
        """Which astropy function the caller asked for: they differ only where a kernel window holds no valid
        input (``convolve`` keeps the NaN, ``convolve_fft`` returns 0.0)."""
        if convolve is None:
            return default
        name = getattr(convolve, '__name__', '')
        if name == 'convolve_fft':
            return True
        if name == 'convolve':
            return False
        raise NotImplementedError("`convolve` must be astropy.convolution.convolve or convolve_fft (got %r): "
                                  "the convolution runs on the device" % (convolve,))

    def _convolved_denominator_counts(self):
        """Device {missing, sampled} = {1, 1}: makes the separable path take the kernel whose denominator is the
        validity map convolved with the float32 factors.  Beam kernels are 16 sigma wide: their outer taps fall
        below the 2^-31 quantum of the sparse-deficit kernel's integer denominator, which would lose the
        interpolation weight of outputs that only such taps connect to valid data (deep inside blank regions)."""
        torch = _torch()
        return torch.ones((2,), dtype=torch.int32, device=self._data.device)

    def _convolve_epilogue(self, out, factor, nan_to_zero, skip=None):
        """In place: ``out *= factor`` (Jy/beam rescale, :3376-3378 / :4230-4233), NaN -> 0 for ``convolve_fft``."""
        if factor == 1.0 and not nan_to_zero:
            return
        lib = _lib.load()
        dtype = _lib.F64 if out.dtype == _torch().float64 else _lib.F32
        nchan, ny, nx = out.shape
        _lib.check(lib.sc_scale(out.data_ptr(), dtype, nchan, ny, nx, out.stride(0), out.stride(1), float(factor),
                                1 if nan_to_zero else 0, skip.data_ptr() if skip is not None else None, _stream()))

    def convolve_to(self, beam, convolve=None, update_function=None, **kwargs):
        """Convolve each channel in the cube to a specified beam; returns a cube with that ``beam``.

        The kernel is ``beam.deconvolve(self.beam).as_kernel(pixscale)``; values in Jy/beam are rescaled by
        ``beam.sr / self.beam.sr``.  ``convolve`` is accepted for drop-in compatibility: with the defaults both
        astropy functions compute the same NaN-interpolating, zero-padded, normalised convolution the device
        kernels do."""
        if not self._mirrors_dask:
            self._refuse_huge('convolve_to')                          # @warn_slow, spectral_cube.py:3334
        self._check_convolve_kwargs(kwargs)
        beam = Beam.coerce(beam)
        if beam == self.beam:
            warnings.warn("The given beam is identical to the current beam. "
                          "Skipping convolution.")
            return self
        kernel = beam.deconvolve(self.beam).as_kernel(self._pixscale_deg())
        factor = beam.sr / self.beam.sr if self._is_jybeam() else 1.0
        fft = self._fft_semantics(convolve, default=not self._mirrors_dask)      # :3336 / dask:1412 defaults
        out = self._run_spatial_smooth(self._kernel_array(kernel, 2), _lib.F32, every_tap=True,
                                       strategy_counts=self._convolved_denominator_counts())
        # planes copied through (:169-172) are neither rescaled nor zeroed
        self._convolve_epilogue(out, factor, fft, skip=self._passthrough_flags)
        cube = self._new_cube_with(data=out) if self._mirrors_dask else self._new_cube_reporting_f64(out)
        cube._attach_beam(beam)
        return cube

    # -- spectral resampling (spectral_cube.py:3224-3332; dask_spectral_cube.py:1250-1373) --------------
    def _interp_axes(self, spectral_grid, suppress_smooth_warning):
        """The checks and flips of spectral_cube.py:3240-3290: (grid ascending, input axis ascending, reverse_in,
        reverse_out, mean output spacing)."""
        grid = np.asarray(getattr(spectral_grid, 'value', spectral_grid), dtype=np.float64)
        inaxis = self.spectral_axis
        indiff = np.mean(np.diff(inaxis))
        outdiff = np.mean(np.diff(grid))
        reverse_out = outdiff < 0
        reverse_in = indiff < 0
        if reverse_out:
            grid = grid[::-1]
            outdiff = np.mean(np.diff(grid))
        if reverse_in:
            inaxis = inaxis[::-1]
            indiff = np.mean(np.diff(inaxis))
        if indiff < 0 or outdiff < 0:
            raise ValueError("impossible.")
        assert np.all(np.diff(grid) > 0)
        assert np.all(np.diff(inaxis) > 0)
        np.testing.assert_allclose(np.diff(grid), outdiff, err_msg="Output grid must be linear")
        if outdiff > 2 * indiff and not suppress_smooth_warning:
            warnings.warn("Input grid has too small a spacing. The data should "
                          "be smoothed prior to resampling.", SmoothingWarning)
        return grid, inaxis, reverse_in, reverse_out, outdiff

    def _interp_wcs(self, grid, reverse_out, outdiff):
        """new spectral WCS: crpix=1, crval = first grid value as given, cdelt = +-mean spacing (:3317-3324)"""
        newwcs = self._wcs.copy()
        inv = 1.0 / self._spectral_scale
        newwcs.crpix[2] = 1.0
        newwcs.crval[2] = (grid[-1] if reverse_out else grid[0]) * inv
        newwcs.cdelt[2] = (-outdiff if reverse_out else outdiff) * inv
        newwcs.pc[2, :] = [0.0, 0.0, 1.0]
        return newwcs

    def _interp_nan_filled(self, fill_value):
        """Is the interpolated data already NaN wherever its new mask excludes it?"""
        return (fill_value is None or fill_value != fill_value) and \
            (self._mirrors_dask or float(self._fill_value) != float(self._fill_value))

    def spectral_interpolate(self, spectral_grid, suppress_smooth_warning=False, fill_value=None,
                             update_function=None, force_rechunk=True, **kwargs):
        """Resample the cube spectrally onto ``spectral_grid`` (values in the cube's spectral unit,
        or a Quantity-like with ``.value``); linear interpolation per spaxel."""
        torch = _torch()
        lib = _lib.load()
        grid, inaxis, reverse_in, reverse_out, outdiff = self._interp_axes(spectral_grid, suppress_smooth_warning)
        src = self._data
        nchan, ny, nx = self.shape
        nout = grid.size
        dask = self._mirrors_dask
        out = torch.empty((nout, ny, nx), dtype=torch.float64 if dask else torch.float32, device=src.device)
        omask = torch.empty((nout, ny, nx), dtype=torch.uint8, device=src.device)
        desc, keep = self._mask_desc()
        ws = self._get_workspace(lib.sc_workspace_bytes(_lib.OP_SPECTRAL_INTERP, nchan, ny, nx, nout))
        (ia, ip), (ga, gp) = _lib.as_double_array(inaxis), _lib.as_double_array(grid)
        _lib.check(lib.sc_spectral_interp(
            src.data_ptr(), out.data_ptr(), _lib.F64 if dask else _lib.F32, omask.data_ptr(),
            nchan, ny, nx, src.stride(0), src.stride(1), nout, desc,
            float('nan') if dask else float(self._fill_value),           # dask:1308 fills with NaN
            ip, gp, 0 if fill_value is None else 1, 0.0 if fill_value is None else float(fill_value),
            1 if reverse_in else 0, 1 if reverse_out else 0, 1 if dask else 0,
            ws.data_ptr(), ws.numel(), _stream()))
        newwcs = self._interp_wcs(grid, reverse_out, outdiff)
        newmask = BooleanArrayMask(omask, wcs=newwcs)
        if dask:
            cube = self._new_cube_with(data=out.to(torch.float32), wcs=newwcs, mask=newmask)
            cube._data_hi = out
        else:
            cube = self._new_cube_with(data=out, wcs=newwcs, mask=newmask)
        cube._mask = newmask
        cube._nan_filled_already = self._interp_nan_filled(fill_value)
        return cube

    def _spectral_interpolate_scatter(self, spectral_grid, chan_ptrs, suppress_smooth_warning=False, fill_value=None,
                                      phase=0, nphases=1):
        """`spectral_interpolate` of these rows with every output channel stored where ``chan_ptrs`` (device int64
        tensor, one address per output channel IN OUTPUT ORDER) says -- the channel owners' buffers of a row-sharded job
        (`distributed.RowShardedCube.spectral_interpolate_to_channels`).  float32 NaN-filled data only: what `reproject`
        consumes.  ``phase`` / ``nphases`` = rank / world: staggers the ranks' marches over the spectrum.  Returns the new
        spectral WCS."""
        torch = _torch()
        lib = _lib.load()
        if not self._interp_nan_filled(fill_value):
            raise ValueError("the scattered form carries no mask: it needs a NaN fill value (the data must say what is masked)")
        grid, inaxis, reverse_in, reverse_out, outdiff = self._interp_axes(spectral_grid, suppress_smooth_warning)
        src = self._data
        nchan, ny, nx = self.shape
        nout = grid.size
        if chan_ptrs.numel() != nout or chan_ptrs.dtype != torch.int64 or chan_ptrs.device != src.device:
            raise ValueError("chan_ptrs: one int64 address per output channel, on the cube's device")
        desc, keep = self._mask_desc()
        ws = self._get_workspace(lib.sc_workspace_bytes(_lib.OP_SPECTRAL_INTERP, nchan, ny, nx, nout))
        (ia, ip), (ga, gp) = _lib.as_double_array(inaxis), _lib.as_double_array(grid)
        _lib.check(lib.sc_spectral_interp_scatter(
            src.data_ptr(), chan_ptrs.data_ptr(), _lib.F32, None,
            nchan, ny, nx, src.stride(0), src.stride(1), nout, desc,
            float('nan') if self._mirrors_dask else float(self._fill_value),
            ip, gp, 0 if fill_value is None else 1, 0.0 if fill_value is None else float(fill_value),
            1 if reverse_in else 0, 1 if reverse_out else 0, 1 if self._mirrors_dask else 0,
            int(phase), int(nphases), ws.data_ptr(), ws.numel(), _stream()))
        return self._interp_wcs(grid, reverse_out, outdiff)

    # -- reprojection (spectral_cube.py:2649-2746) -------------------------------------------------------
    _ORDERS = {'nearest-neighbor': 0, 'bilinear': 1, 0: 0, 1: 1}

    def _pixel_map(self, wcs_out, ny_out, nx_out):
        """float64 device planes (yin, xin): where every output pixel falls in this cube's image."""
        torch = _torch()
        lib = _lib.load()
        dev = self._data.device
        yin = torch.empty((ny_out, nx_out), dtype=torch.float64, device=dev)
        xin = torch.empty((ny_out, nx_out), dtype=torch.float64, device=dev)
        (oa, op), (ia, ip) = _lib.as_double_array(wcs_out.celestial_params()), _lib.as_double_array(self._wcs.celestial_params())
        _lib.check(lib.sc_wcs_pixel_map(op, ip, ny_out, nx_out, yin.data_ptr(), xin.data_ptr(), _stream()))
        return yin, xin

    def _run_reproject(self, yin, xin, order, filled=True, out_dtype=None, want_f32=False):
        """One pass: the float64 result `reproject_interp` returns, its float32 working copy, the
        footprint and the "anything valid at all" flag (spectral_cube.py:2726-2746)."""
        torch = _torch()
        lib = _lib.load()
        src = self._data
        nchan, ny, nx = self.shape
        ny_out, nx_out = yin.shape
        out_dtype = _lib.F64 if out_dtype is None else out_dtype
        out = torch.empty((nchan, ny_out, nx_out), dtype=torch.float64 if out_dtype == _lib.F64 else torch.float32,
                          device=src.device)
        out32 = (torch.empty((nchan, ny_out, nx_out), dtype=torch.float32, device=src.device)
                 if out_dtype == _lib.F64 and want_f32 else None)
        foot = torch.empty((nchan, ny_out, nx_out), dtype=torch.uint8, device=src.device)
        flag = torch.empty((1,), dtype=torch.int32, device=src.device)
        desc, keep = self._mask_desc() if filled else lower_mask(None, src)
        _lib.check(lib.sc_reproject_ex(src.data_ptr(), out.data_ptr(), out_dtype,
                                       out32.data_ptr() if out32 is not None else None, foot.data_ptr(), flag.data_ptr(),
                                       nchan, ny, nx, src.stride(0), src.stride(1), ny_out, nx_out, desc,
                                       float(self._fill_value), yin.data_ptr(), xin.data_ptr(), order, _stream()))
        return out, (out32 if out32 is not None else (out if out_dtype != _lib.F64 else None)), foot, flag

    def reproject(self, header, order='bilinear', use_memmap=False, filled=True, **kwargs):
        """Spatially reproject the cube into a new header (a FITS-header-like mapping with NAXISn and
        celestial WCS keywords, or a ``CubeWCS`` plus ``shape_out=``)."""
        torch = _torch()
        if order not in self._ORDERS:
            raise NotImplementedError("order=%r: only 'nearest-neighbor' and 'bilinear' are implemented" % (order,))
        if hasattr(header, 'celestial_params'):
            newwcs = header
            shape_out = tuple(kwargs.pop('shape_out'))
        else:
            newwcs = as_cube_wcs(header)
            shape_out = tuple(int(header['NAXIS%d' % (i + 1)]) for i in range(int(header['NAXIS'])))[::-1]
        if shape_out[0] != self.shape[0]:
            raise ValueError("reproject() resamples the celestial axes only; use spectral_interpolate for the "
                             "spectral axis (spectral_cube.py:2656-2657)")
        self._refuse_huge('reproject')                                # @warn_slow, spectral_cube.py:2649
        yin, xin = self._pixel_map(newwcs, shape_out[1], shape_out[2])
        out, out32, foot, flag = self._run_reproject(yin, xin, self._ORDERS[order], filled=filled)
        if int(flag.item()) == 0:
            raise ValueError("All values in reprojected cube are nan.  This can be caused"
                             " by an error in which coordinates do not 'round-trip'.  Try "
                             "setting ``roundtrip_coords=False``.  You might also check "
                             "whether the WCS transformation produces valid pixel->world "
                             "and world->pixel coordinates in each axis.")
        newmask = BooleanArrayMask(foot, wcs=newwcs)
        # reproject_interp returns float64: that is the cube's data; its float32 working copy is made on first use
        cube = self._new_cube_with(data=_PendingFloat32Copy(out), wcs=newwcs, mask=newmask)
        cube._mask = newmask
        return cube

    # -- reductions along the spectral axis (spectral_cube.py:361-470, 578-826; dask:641-767) ------------
    def _reduce_axis0_raw(self, want):
        """One pass over the cube for every requested statistic.  `want` is a set out of
        {'sum', 'count', 'm2', 'min', 'max', 'argmin', 'argmax'}; returns name -> device tensor (ny, nx)."""
        torch = _torch()
        lib = _lib.load()
        src = self._materialized()._data
        nchan, ny, nx = self.shape
        dev = src.device
        dt = {'sum': torch.float64, 'count': torch.int32, 'm2': torch.float64, 'min': torch.float32,
              'max': torch.float32, 'argmin': torch.int32, 'argmax': torch.int32}
        outs = dict((k, torch.empty((ny, nx), dtype=dt[k], device=dev)) for k in want)
        ptr = lambda k: outs[k].data_ptr() if k in outs else None
        desc, keep = self._mask_desc()
        _lib.check(lib.sc_reduce_axis0(src.data_ptr(), nchan, ny, nx, src.stride(0), src.stride(1), desc,
                                       ptr('sum'), ptr('count'), ptr('m2'), ptr('min'), ptr('max'),
                                       ptr('argmin'), ptr('argmax'), _stream()))
        return outs

    def _reduce_spatial_raw(self, axis, want):
        """`_reduce_axis0_raw` along numpy axis 1 or 2 (`sc_reduce_spatial`): name -> device tensor (nchan, nx | ny)."""
        torch = _torch()
        lib = _lib.load()
        src = self._materialized()._data
        nchan, ny, nx = self.shape
        shape = (nchan, nx) if axis == 1 else (nchan, ny)
        dt = {'sum': torch.float64, 'count': torch.int32, 'm2': torch.float64, 'min': torch.float32,
              'max': torch.float32, 'argmin': torch.int32, 'argmax': torch.int32}
        outs = dict((k, torch.empty(shape, dtype=dt[k], device=src.device)) for k in want)
        ptr = lambda k: outs[k].data_ptr() if k in outs else None
        desc, keep = self._mask_desc()
        _lib.check(lib.sc_reduce_spatial(src.data_ptr(), nchan, ny, nx, src.stride(0), src.stride(1), int(axis), desc,
                                         ptr('sum'), ptr('count'), ptr('m2'), ptr('min'), ptr('max'),
                                         ptr('argmin'), ptr('argmax'), _stream()))
        return outs

    def _reduce_raw(self, axis, want):
        return self._reduce_axis0_raw(want) if axis in (0, None) else self._reduce_spatial_raw(axis, want)

    def _reduction_axis(self, axis, name, how=None):
        if not self._mirrors_dask:
            self._refuse_huge(name, how=how, accepts_how=True)       # @warn_slow, spectral_cube.py:577-826
        if axis not in (0, 1, 2, None):
            raise NotImplementedError("%s(axis=%r): a cube has axes 0 (spectral), 1 and 2 (spatial)" % (name, axis))

    def _collapsed(self, values, unit, axis=0):
        """Projection of a collapsed axis (spectral_cube.py:395-414)."""
        meta = {'collapse_axis': axis or 0}
        meta.update(self._meta)
        return Projection(values, unit=unit, wcs=self._wcs.drop_axis(axis or 0), meta=meta, header=self._header,
                          copy=False)

    def _np_dtype(self):
        return np.float32          # the cube's dtype: the nan-functions of the reference keep it

    # axis=None: the per-spaxel partials of the one device pass are combined on the (ny, nx) maps -- S values
    # against the V voxels of the pass.  Scalars come back as numpy float32 (the reference wraps them in a Quantity).
    def sum(self, axis=None, how='auto', **kwargs):
        """Sum over the spectral axis (or everything); nothing included -> NaN (np_compat.allbadtonan)."""
        self._reduction_axis(axis, 'sum', how)
        r = self._reduce_raw(axis, {'sum', 'count'} if axis is None else {'sum'})
        if axis is None:
            return self._np_dtype()(whole_sum(_torch(), r['sum'], r['count']))
        return self._collapsed(r['sum'].cpu().numpy().astype(self._np_dtype()), self._unit, axis)

    def mean(self, axis=None, how='cube', **kwargs):
        self._reduction_axis(axis, 'mean', how)
        torch = _torch()
        r = self._reduce_raw(axis, {'sum', 'count'})
        if axis is None:
            return self._np_dtype()(whole_mean(torch, r['sum'], r['count']))
        out = r['sum'] / r['count'].to(torch.float64)            # 0 / 0 never happens: sum is NaN there
        return self._collapsed(out.cpu().numpy().astype(self._np_dtype()), self._unit, axis)

    def std(self, axis=None, how='cube', ddof=0, **kwargs):
        self._reduction_axis(axis, 'std', how)
        torch = _torch()
        if axis is None:
            r = self._reduce_raw(axis, {'sum', 'count', 'm2'})
            return self._np_dtype()(whole_std(torch, r['sum'], r['count'], r['m2'], ddof))
        r = self._reduce_raw(axis, {'m2', 'count'})
        n = r['count'].to(torch.float64) - float(ddof)
        out = torch.sqrt(r['m2'] / torch.where(n > 0, n, torch.full_like(n, float('nan'))))
        return self._collapsed(out.cpu().numpy().astype(self._np_dtype()), self._unit, axis)

    def max(self, axis=None, how='auto', **kwargs):
        self._reduction_axis(axis, 'max', how)
        m = self._reduce_raw(axis, {'max'})['max']
        if axis is None:
            return self._np_dtype()(whole_extremum(_torch(), m, 'max'))
        return self._collapsed(m.cpu().numpy(), self._unit, axis)

    def min(self, axis=None, how='auto', **kwargs):
        self._reduction_axis(axis, 'min', how)
        m = self._reduce_raw(axis, {'min'})['min']
        if axis is None:
            return self._np_dtype()(whole_extremum(_torch(), m, 'min'))
        return self._collapsed(m.cpu().numpy(), self._unit, axis)

    def _arg_extremum(self, axis, which, how=None):
        self._reduction_axis(axis, 'arg' + which, how)
        r = self._reduce_raw(axis, {which, 'arg' + which})
        if axis is not None:
            return r['arg' + which].cpu().numpy().astype(np.int64)
        return whole_arg_extremum(_torch(), r[which], r['arg' + which], which)

    def argmax(self, axis=None, how='auto', **kwargs):
        """Channel of the (first) maximum; arbitrary (0) where nothing is included (spectral_cube.py:800-811)."""
        return self._arg_extremum(axis, 'max', how)

    def argmin(self, axis=None, how='auto', **kwargs):
        return self._arg_extremum(axis, 'min', how)

    # -- moments (spectral_cube.py:1614-1763; dask_spectral_cube.py:1031-1132) -----------------------
    def _moments_axis0_raw(self, want_bits):
        """Run the fused kernel; returns dict order -> float64 device tensor (ny, nx), with units
        folded in the way ``moment()`` does (M1 already carries the channel-0 world offset)."""
        torch = _torch()
        lib = _lib.load()
        if self._pending is not None and self._data_t is None:
            return self._pending.moments(self, want_bits)
        nchan, ny, nx = self.shape
        dev = self._data.device
        outs = {}
        ptrs = []
        for bit, order in ((1, 0), (2, 1), (4, 2)):
            if want_bits & bit:
                outs[order] = torch.empty((ny, nx), dtype=torch.float64, device=dev)
                ptrs.append(outs[order].data_ptr())
            else:
                ptrs.append(None)
        desc, keep = self._mask_desc()
        wsb = lib.sc_workspace_bytes(_lib.OP_MOMENTS, nchan, ny, nx, 0)
        ws = self._get_workspace(wsb)
        xoff, xptr = _lib.as_double_array(self._spectral_offsets())
        _lib.check(lib.sc_moments_axis0(
            self._data.data_ptr(), nchan, ny, nx, self._data.stride(0), self._data.stride(1), desc,
            xptr, float(self._pix_size_slice(0)), self._world0_spectral(), want_bits,
            ptrs[0], ptrs[1], ptrs[2], ws.data_ptr(), ws.numel(), _stream()))
        return outs

    def _moment_unit(self, order, axis):
        axunit = self._spectral_unit if axis == 0 else 'deg'
        if order == 0:
            return _unit_mul(self._unit, axunit)
        return _unit_pow(axunit, max(order, 1))

    def _moment_meta(self, order, axis, how):
        meta = {'moment_order': order, 'moment_axis': axis}
        if not self._mirrors_dask:
            meta['moment_method'] = how                 # spectral_cube.py:1714-1716 vs dask:1127-1128
        meta.update(self._meta)
        return meta

    def moment(self, order=0, axis=0, how='auto', **kwargs):
        if axis == 0 and order == 2:
            warnings.warn("Note that the second moment returned will be a "
                          "variance map. To get a linewidth map, use the "
                          "SpectralCube.linewidth_fwhm() or "
                          "SpectralCube.linewidth_sigma() methods instead.",
                          VarianceWarning)
        if not self._mirrors_dask and how not in ('slice', 'cube', 'ray', 'auto'):
            # the reference *returns* the exception object (spectral_cube.py:1687-1689)
            return ValueError("Invalid how. Must be in %s" % sorted(['slice', 'cube', 'ray', 'auto']))
        if axis == 0:
            out = self._moment_axis0(order)
        else:
            out = self._moment_spatial(order, axis)
        return Projection(out, unit=self._moment_unit(order, axis), wcs=self._wcs.drop_axis(axis),
                          meta=self._moment_meta(order, axis, how), header=self._header, copy=False)

    def _moment_axis0(self, order):
        torch = _torch()
        lib = _lib.load()
        if order in (0, 1, 2):
            return self._moments_axis0_raw(1 << order)[order].cpu().numpy()
        # higher orders: first moment pass, then a central-moment pass (_moments.py:108-123)
        nchan, ny, nx = self.shape
        m1 = self._moments_axis0_raw(_lib.WANT_M1)[1]
        centre = m1 - self._world0_spectral()
        out = torch.empty((ny, nx), dtype=torch.float64, device=self._data.device)
        desc, keep = self._mask_desc()
        ws = self._get_workspace(nchan * 8 + 512)
        xoff, xptr = _lib.as_double_array(self._spectral_offsets())
        src = self._materialized()
        _lib.check(lib.sc_moment_central_axis0(
            src._data.data_ptr(), nchan, ny, nx, src._data.stride(0), src._data.stride(1), desc,
            xptr, centre.data_ptr(), int(order), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
        return out.cpu().numpy()

    def _pixel_offsets(self, axis):
        """``_pix_cen()[axis]`` for a spatial axis as a (ny, nx) float64 device plane (degrees)."""
        torch = _torch()
        lib = _lib.load()
        nchan, ny, nx = self.shape
        off = torch.empty((ny, nx), dtype=torch.float64, device=self._data.device)
        ws = self._get_workspace(ny * nx * 16 + 512)
        wa, wp = _lib.as_double_array(self._wcs.celestial_params())
        _lib.check(lib.sc_pixel_offsets(wp, ny, nx, axis, off.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
        return off

    def _moment_spatial(self, order, axis):
        torch = _torch()
        lib = _lib.load()
        if order not in (0, 1, 2):
            raise NotImplementedError("moments of order > 2 along spatial axes")
        nchan, ny, nx = self.shape
        src = self._data
        off = self._pixel_offsets(axis)
        out = torch.empty((nchan, nx if axis == 1 else ny), dtype=torch.float64, device=src.device)
        desc, keep = self._mask_desc()
        ptrs = [out.data_ptr() if o == order else None for o in (0, 1, 2)]
        _lib.check(lib.sc_moments_spatial(src.data_ptr(), nchan, ny, nx, src.stride(0), src.stride(1), axis, desc,
                                          off.data_ptr(), float(self._pix_size_slice(axis)), 1 << order,
                                          ptrs[0], ptrs[1], ptrs[2], _stream()))
        return out.cpu().numpy()

    def _materialized(self):
        return self

    def moment0(self, axis=0, how='auto', **kwargs):
        return self.moment(axis=axis, order=0, how=how, **kwargs)

    def moment1(self, axis=0, how='auto', **kwargs):
        return self.moment(axis=axis, order=1, how=how, **kwargs)

    def moment2(self, axis=0, how='auto', **kwargs):
        return self.moment(axis=axis, order=2, how=how, **kwargs)

    def moments012(self, how='auto'):
        """Extension: moment 0, 1 and 2 from ONE pass over the cube (the reference needs four)."""
        raw = self._moments_axis0_raw(7)
        return tuple(Projection(raw[o].cpu().numpy(), unit=self._moment_unit(o, 0),
                                wcs=self._wcs.drop_axis(0), meta=self._moment_meta(o, 0, how),
                                header=self._header, copy=False) for o in (0, 1, 2))

    def linewidth_sigma(self, how='auto', **kwargs):
        with np.errstate(invalid='ignore'):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", VarianceWarning)
                return self.moment2(how=how, **kwargs).sqrt()

    def linewidth_fwhm(self, how='auto', **kwargs):
        s = self.linewidth_sigma(**kwargs)               # the reference drops `how` here (:1763)
        return s._with(s.value * SIGMA2FWHM, s.unit)


class _PendingFloat32Copy(object):
    """The float32 working copy of a cube whose result is float64 (`reproject` returns what `reproject_interp` does),
    not made until something asks for it: `_data_hi`, `unmasked_data`, `filled_data`, `write` and `mosaic_cubes` read
    the float64 tensor, and the reproject kernel saves 4 of its 13 output bytes per voxel."""

    def __init__(self, hi):
        self.hi = hi
        self.shape = tuple(hi.shape)

    @property
    def source(self):                       # `.source._data.device` is where the cube lives.  (A property, not an attribute
        return self                         # holding `self`: that cycle kept the 17 GB result until the cyclic GC came by.)

    @property
    def _data(self):
        return self.hi

    @property
    def _data_t(self):
        return self.hi

    def materialize(self):
        return self.hi.to(_torch().float32)

    def moments(self, cube, want_bits):
        cube._data                          # makes the float32 copy; the cube is an ordinary one from here on
        cube._pending = None
        return cube._moments_axis0_raw(want_bits)


class _PendingSpectralSmooth(object):
    """A spectral_smooth recorded but not yet run (the dask class is lazy, dask:880-917).  A moment
    requested on the result runs the fused kernel and never writes the smoothed cube; any other
    access materialises it (float32, dask:829)."""

    def __init__(self, source, taps):
        self.source, self.taps = source, taps
        self.shape = source.shape

    def materialize(self):
        return self.source._run_spectral_smooth(self.taps, _lib.F32)

    def moments(self, cube, want_bits):
        torch = _torch()
        lib = _lib.load()
        src_cube = self.source
        src = src_cube._data
        nchan, ny, nx = self.shape
        outs, ptrs = {}, []
        for bit, order in ((1, 0), (2, 1), (4, 2)):
            if want_bits & bit:
                outs[order] = torch.empty((ny, nx), dtype=torch.float64, device=src.device)
                ptrs.append(outs[order].data_ptr())
            else:
                ptrs.append(None)
        # the smoothed cube keeps the OLD mask object; lowered against the source tensor it is the
        # include mask of the original data (spectral_cube.py:3043-3045)
        desc, keep = lower_mask(cube._mask, src)
        wsb = lib.sc_workspace_bytes(_lib.OP_SMOOTH_MOMENTS, nchan, ny, nx, len(self.taps))
        ws = cube._get_workspace(wsb)
        xoff, xptr = _lib.as_double_array(cube._spectral_offsets())
        tarr, tptr = _lib.as_double_array(self.taps)
        _lib.check(lib.sc_smooth_moments_axis0(
            src.data_ptr(), nchan, ny, nx, src.stride(0), src.stride(1), desc, float(src_cube._smooth_fill()),
            tptr, len(self.taps), _lib.F32, xptr, float(cube._pix_size_slice(0)), cube._world0_spectral(),
            want_bits, ptrs[0], ptrs[1], ptrs[2], ws.data_ptr(), ws.numel(), _stream()))
        return outs


class _Sliceable(object):
    def __init__(self, getter):
        self._getter = getter

    def __getitem__(self, view):
        return self._getter(view)


class SpectralCube(BaseSpectralCube):
    """Mirrors the numpy-backed reference class (spectral_cube.py:3691-3765).
    ``SpectralCube(..., use_dask=True)`` returns a ``DaskSpectralCube`` (:3697-3702)."""

    def __new__(cls, *args, **kwargs):
        if kwargs.pop('use_dask', False) and cls is SpectralCube:
            return super(SpectralCube, cls).__new__(DaskSpectralCube)
        return super(SpectralCube, cls).__new__(cls)

    def __init__(self, data, wcs, mask=None, meta=None, fill_value=np.nan, header=None,
                 allow_huge_operations=False, beam=None, wcs_tolerance=0.0, use_dask=False, **kwargs):
        super(SpectralCube, self).__init__(data=data, wcs=wcs, mask=mask, meta=meta,
                                           fill_value=fill_value, header=header,
                                           allow_huge_operations=allow_huge_operations, **kwargs)
        # beam loading happens after the WCS is read (spectral_cube.py:3716-3731; cube_utils.try_load_beam)
        if beam is None:
            if isinstance(self._meta.get('beam'), Beam):
                beam = self._meta['beam']
            else:
                try:                                          # cube_utils.try_load_beam: no beam is fine
                    beam = Beam.from_fits_header(self._header)
                except Exception:
                    beam = None
        else:
            beam = Beam.coerce(beam)
        self._attach_beam(beam)

    @classmethod
    def read(cls, filename, format='fits', use_dask=False, **kwargs):
        """``SpectralCube.read`` for FITS files (io/core.py -> io/fits.py:171-260): the data block is decoded
        on the device (`sc_fits_decode`), the cube carries ``LazyMask(np.isfinite)`` like the reference's."""
        if format not in (None, 'fits'):
            raise NotImplementedError("only FITS cubes can be read (format=%r)" % (format,))
        from .io_fits import load_fits_cube
        target = DaskSpectralCube if (use_dask or cls is DaskSpectralCube) else SpectralCube
        return load_fits_cube(filename, target, **kwargs)

    def write(self, filename, overwrite=False, format='fits'):
        """io/fits.py:262-282 -> ``cube.hdu`` = ``PrimaryHDU(self.unitless_filled_data[:], header=self.header)``
        (spectral_cube.py:2563-2570; dask :1400-1405): the FILLED data -- masked voxels are written as the fill
        value -- in the dtype the reference's cube holds (float64 after reproject / numpy-class smoothing)."""
        from .io_fits import write_fits
        hdr = dict(self._header or {})
        hdr.update(self._wcs.to_header())
        if self._unit:
            hdr['BUNIT'] = str(self._unit)
        write_fits(filename, self._get_filled_data(fill=self._fill_value), hdr, overwrite=overwrite)


class DaskSpectralCube(SpectralCube):
    """Mirrors the dask-backed reference class (dask_spectral_cube.py:1376-1650)."""
    _mirrors_dask = True

    def use_dask_scheduler(self, scheduler, num_workers=None):
        """Accepted for drop-in compatibility (dask_spectral_cube.py:278-312); scheduling is the
        GPU's business here."""
        return _NullContext()

    def rechunk(self, *args, **kwargs):
        return self

    def _smooth_fill(self):
        return np.nan                        # dask_spectral_cube.py:816, 823

    def statistics(self):
        """Global basic statistics of the data (dask_spectral_cube.py:769-814): npts, min, max, sum, sumsq, mean,
        sigma, rms -- ONE pass over the cube (`sc_reduce_axis0`), combined on the (ny, nx) partial maps.  Values
        are plain floats in the cube's unit (its square for sumsq); the reference wraps them in Quantities."""
        r = self._reduce_axis0_raw({'sum', 'count', 'm2', 'min', 'max'})
        return whole_statistics(_torch(), r['sum'], r['count'], r['m2'], r['min'], r['max'])

    def spectral_smooth(self, kernel, convolve=None, save_to_tmp_dir=False, **kwargs):
        """Lazy like the reference's dask class; ``save_to_tmp_dir=True`` computes straight away."""
        self._check_convolve_kwargs(kwargs)
        if self._fft_semantics(convolve, default=False):
            # `convolve_fft` turns empty windows into 0.0: not expressible in the lazy fused form, run it now
            out = self._run_spectral_smooth(self._kernel_array(kernel, 1), _lib.F32)
            self._convolve_epilogue(out, 1.0, True)
            return self._new_cube_with(data=out)
        taps = self._kernel_array(kernel, 1)
        if self._mask is not None and self._pending is not None and self._data_t is None:
            self._data                        # chain of lazy ops: materialise the inner one first
        cube = self._new_cube_with()
        cube._data_t = None
        cube._pending = _PendingSpectralSmooth(self, taps)
        if save_to_tmp_dir:
            cube._data
        return cube


class BeamWarning(SpectralCubeWarning):
    pass


class NonFiniteBeamsWarning(SpectralCubeWarning):
    pass


class VaryingResolutionSpectralCube(BaseSpectralCube):
    """A cube with one beam per channel (spectral_cube.py:3767-3894, base_class.py:476-542): ``beams`` is a
    ``Beams`` / list of ``Beam``, or ``beam_table`` a mapping / record array with BMAJ, BMIN (arcsec) and BPA
    (deg) columns.  ``use_dask=True`` gives the dask-mirroring variant (:3777-3782)."""

    def __new__(cls, *args, **kwargs):
        if kwargs.pop('use_dask', False) and cls is VaryingResolutionSpectralCube:
            return super(VaryingResolutionSpectralCube, cls).__new__(DaskVaryingResolutionSpectralCube)
        return super(VaryingResolutionSpectralCube, cls).__new__(cls)

    def __init__(self, *args, **kwargs):
        beam_table = kwargs.pop('beam_table', None)
        beams = kwargs.pop('beams', None)
        self.beam_threshold = kwargs.pop('beam_threshold', 0.01)
        goodbeams_mask = kwargs.pop('goodbeams_mask', None)
        kwargs.pop('use_dask', None)
        kwargs.pop('wcs_tolerance', None)
        if beam_table is None and beams is None:
            raise ValueError("Must give either a beam table or a list of beams to "
                             "initialize a VaryingResolutionSpectralCube")
        super(VaryingResolutionSpectralCube, self).__init__(*args, **kwargs)
        if beam_table is not None:
            # CASA beam tables are in arcsec (:3838-3846)
            beams = Beams.from_arcsec(np.asarray(beam_table['BMAJ'], dtype=np.float64),
                                      np.asarray(beam_table['BMIN'], dtype=np.float64),
                                      np.asarray(beam_table['BPA'], dtype=np.float64))
            good = beams.isfinite
            self._goodbeams_mask = good
            if not np.all(good):
                warnings.warn("There were {0} non-finite beams; layers with "
                              "non-finite beams will be masked out.".format(np.count_nonzero(~good)),
                              NonFiniteBeamsWarning)
            beam_mask = BooleanArrayMask(good[:, None, None], self._wcs, shape=self.shape)
            self._mask = beam_mask if self._mask is None else self._mask & beam_mask
        if not isinstance(beams, Beams):
            beams = Beams(beams)
        if len(beams) != self.shape[0]:
            raise ValueError("Beam list must have same size as spectral "
                             "dimension")
        self._beams = beams
        if goodbeams_mask is not None:
            self.goodbeams_mask = goodbeams_mask

    # -- MultiBeamMixinClass (base_class.py:476-542) ------------------------------------------------
    @property
    def beams(self):
        return self._beams[self.goodbeams_mask]

    @property
    def unmasked_beams(self):
        return self._beams

    @property
    def goodbeams_mask(self):
        if getattr(self, '_goodbeams_mask', None) is not None:
            return self._goodbeams_mask
        return self._beams.isfinite

    @goodbeams_mask.setter
    def goodbeams_mask(self, value):
        value = np.asarray(value, dtype=bool)
        if value.size != self.shape[0]:
            raise ValueError("The 'good beams' mask must have the same size "
                             "as the cube's spectral dimension")
        self._goodbeams_mask = value

    def _new_cube_with(self, **kw):
        cube = BaseSpectralCube._new_cube_with(self, **kw)
        if isinstance(cube, VaryingResolutionSpectralCube):
            cube._beams = kw.get('beams', self._beams)
            cube._goodbeams_mask = kw.get('goodbeams_mask', getattr(self, '_goodbeams_mask', None))
            cube.beam_threshold = self.beam_threshold
        return cube

    def __getitem__(self, view):
        from .masks import _slice3
        cube = BaseSpectralCube.__getitem__(self, view)
        spec = _slice3(view)[0]
        cube._beams = self._beams[spec]
        if getattr(self, '_goodbeams_mask', None) is not None:
            cube._goodbeams_mask = self._goodbeams_mask[spec]
        return cube

    def mask_channels(self, goodchannels):
        """:4270-4300 -- the beams of the masked channels are skipped by ``convolve_to``."""
        cube = BaseSpectralCube.mask_channels(self, goodchannels)
        cube.goodbeams_mask = np.logical_and(np.asarray(goodchannels, dtype='bool'), self.goodbeams_mask)
        return cube

    def spectral_interpolate(self, *args, **kwargs):
        raise AttributeError("VaryingResolutionSpectralCubes can't be "
                             "spectrally interpolated.  Convolve to a "
                             "common resolution with `convolve_to` before "
                             "attempting spectral interpolation.")

    def spectral_smooth(self, *args, **kwargs):
        raise AttributeError("VaryingResolutionSpectralCubes can't be "
                             "spectrally smoothed.  Convolve to a "
                             "common resolution with `convolve_to` before "
                             "attempting spectral smoothed.")

    # -- convolve_to (:4127-4240; dask_spectral_cube.py:1512-1630) ----------------------------------
    _result_class = None       # set below: SpectralCube / DaskSpectralCube

    def _channel_plan(self, beam, allow_smaller):
        """Per channel: (kernel array or None, Jy/beam factor), following the reference's loop (:4186-4209)."""
        pixscale = self._pixscale_deg()
        jybeam = self._is_jybeam()
        plan = []
        for bm, valid in zip(self.unmasked_beams, self.goodbeams_mask):
            if not valid or beam == bm:
                plan.append((None, 1.0))        # masked-out beams are skipped; equal beams are a point response
                continue
            try:
                kernel = beam.deconvolve(bm).as_kernel(pixscale)
                plan.append((self._kernel_array(kernel, 2), beam.sr / bm.sr if jybeam else 1.0))
            except ValueError:
                if not allow_smaller:
                    raise
                plan.append((None, 1.0))
        return plan

    def convolve_to(self, beam, allow_smaller=False, convolve=None, update_function=None, **kwargs):
        """Convolve each channel in the cube to a specified beam; returns a single-beam cube.

        Every channel has its own deconvolved kernel; each plane is one launch of the spatial kernels writing
        its slice of the result (the planes of a real cube are tens of MB: the launches fill the chip), followed
        by the in-place Jy/beam rescale of that plane."""
        if not self._mirrors_dask:
            self._refuse_huge('convolve_to')                          # @warn_slow, spectral_cube.py:4126
        self._check_convolve_kwargs(kwargs)
        beam = Beam.coerce(beam)
        pc = self._wcs.pc
        if pc[0, 1] != 0 or pc[1, 0] != 0:
            warnings.warn("The beams will produce convolution kernels "
                          "that are not aware of any misaligment "
                          "between pixel and world coordinates, "
                          "and there are off-diagonal elements of the "
                          "WCS spatial transformation matrix.  "
                          "Unexpected results are likely.", BeamWarning)
        plan = self._channel_plan(beam, allow_smaller)
        fft = self._fft_semantics(convolve, default=not self._mirrors_dask)      # :4128 / dask:1513 defaults
        torch = _torch()
        lib = _lib.load()
        nchan, ny, nx = self.shape
        out = torch.empty((nchan, ny, nx), dtype=torch.float32, device=self._data.device)
        fill = float(self._fill_value)
        counts = self._convolved_denominator_counts()
        for ii, (k2d, factor) in enumerate(plan):
            plane = BaseSpectralCube.__getitem__(self, (slice(ii, ii + 1), slice(None), slice(None)))
            dst = out[ii:ii + 1]
            if k2d is None:
                # newdata[ii] = img, the FILLED plane (:4222-4225)
                src = plane._data
                if plane._mask is None:
                    dst.copy_(src)
                else:
                    desc, keep = plane._mask_desc()
                    _lib.check(lib.sc_fill_masked(src.data_ptr(), 1, ny, nx, src.stride(0), src.stride(1), desc,
                                                  fill, dst.data_ptr(), _stream()))
            else:
                plane._run_spatial_smooth(k2d, _lib.F32, out=dst, passthrough=0, strategy_counts=counts, every_tap=True)
                self._convolve_epilogue(dst, factor, fft)
            if update_function is not None:
                update_function()
        cube = BaseSpectralCube._new_cube_with(self, data=out, cls=self._result_class, beam=beam)
        if not cube._mirrors_dask:
            cube._hi_is_widened_f32 = True       # the reference's buffer is float64 (:4214)
        return cube


class DaskVaryingResolutionSpectralCube(VaryingResolutionSpectralCube):
    """dask_spectral_cube.py:1467-1650"""
    _mirrors_dask = True

    def use_dask_scheduler(self, scheduler, num_workers=None):
        return _NullContext()

    def rechunk(self, *args, **kwargs):
        return self


def _on_cube_device(method):
    """Run `method` with the CUDA device that holds the cube's voxels as the current device: the library launches on
    the calling thread's current device and `_stream()` is that device's current stream, so a cube on cuda:1 must
    not be processed while cuda:0 is current."""
    import functools

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        t = getattr(self, '_data_t', None)
        if t is None and getattr(self, '_pending', None) is not None:
            t = self._pending.source._data_t
        with _lib.on_device_of(t):
            return method(self, *args, **kwargs)
    return wrapper


def _guard_device_of_methods(*classes):
    import types
    for cls in classes:
        for name, attr in list(vars(cls).items()):
            if isinstance(attr, types.FunctionType) and not (name.startswith('__') and name.endswith('__')):
                setattr(cls, name, _on_cube_device(attr))


VaryingResolutionSpectralCube._result_class = SpectralCube
DaskVaryingResolutionSpectralCube.statistics = DaskSpectralCube.statistics        # DaskSpectralCubeMixin (dask:769-814)
DaskVaryingResolutionSpectralCube._result_class = DaskSpectralCube


class _NullContext(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_guard_device_of_methods(BaseSpectralCube, SpectralCube, DaskSpectralCube, VaryingResolutionSpectralCube,
                         DaskVaryingResolutionSpectralCube)
