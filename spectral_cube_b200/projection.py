"""
Result containers of the hot path.

The reference returns ``Projection`` (a ``Quantity`` subclass carrying ``wcs``, ``meta`` and
``header``; ``lower_dimensional_structures.py:246-292``).  astropy is not a dependency here,
so ``Projection`` is an ``ndarray`` subclass with the same attribute names (``value``,
``unit``, ``wcs``, ``meta``, ``header``); ``unit`` is a plain string unless astropy units
were given, and ``to_quantity()`` upgrades to a real ``astropy.units.Quantity`` when astropy
is installed.
"""
import numpy as np


def _unit_mul(a, b):
    if not isinstance(a, str) or not isinstance(b, str):
        return a * b
    return ('%s %s' % (a, b)).strip()


def _unit_pow(a, n):
    if not isinstance(a, str):
        return a ** n
    return a if n == 1 else '%s%d' % (a, n)


class LowerDimensionalObject(np.ndarray):
    def __new__(cls, value, unit=None, wcs=None, meta=None, header=None, copy=True):
        obj = np.array(value, copy=copy).view(cls) if copy else np.asarray(value).view(cls)
        obj._unit = unit
        obj._wcs = wcs
        obj._meta = {} if meta is None else dict(meta)
        obj._header = header
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self._unit = getattr(obj, '_unit', None)
        self._wcs = getattr(obj, '_wcs', None)
        self._meta = getattr(obj, '_meta', {})
        self._header = getattr(obj, '_header', None)

    @property
    def value(self):
        return np.asarray(self)

    @property
    def unit(self):
        return self._unit

    @property
    def wcs(self):
        return self._wcs

    @property
    def meta(self):
        return self._meta

    @property
    def header(self):
        """The parent cube's non-WCS cards + this object's own WCS (lower_dimensional_structures.py:66-97 builds it
        from ``self.wcs.to_header()``): every WCS keyword of an axis above ``ndim`` is dropped, so a moment map of a
        cube read from FITS does not carry the collapsed axis's CTYPE3/CRVAL3/... into a NAXIS=2 file."""
        hdr = dict((k, v) for k, v in (self._header or {}).items() if not _is_wcs_card_above(k, self.ndim))
        if self._wcs is not None and hasattr(self._wcs, 'to_header'):
            hdr.update(self._wcs.to_header())
        hdr['BUNIT'] = str(self._unit) if self._unit is not None else ''
        for i, n in enumerate(self.shape[::-1]):
            hdr['NAXIS%d' % (i + 1)] = n
        hdr['NAXIS'] = self.ndim
        return hdr

    def write(self, filename, format=None, overwrite=False):
        """lower_dimensional_structures.py:115-124 -> io/fits.py:284-300: the map with its celestial WCS."""
        from .io_fits import write_fits
        hdr = dict((k, v) for k, v in self.header.items() if not isinstance(v, (dict, list, tuple)))
        write_fits(filename, np.asarray(self.value), hdr, overwrite=overwrite)

    def to_quantity(self):
        import astropy.units as u          # optional
        return u.Quantity(self.value, u.Unit(self._unit) if isinstance(self._unit, str) else self._unit)

    def _with(self, value, unit):
        return type(self)(value, unit=unit, wcs=self._wcs, meta=self._meta, header=self._header, copy=False)

    def sqrt(self):
        with np.errstate(invalid='ignore'):
            v = np.sqrt(self.value)
        unit = self._unit
        if isinstance(unit, str) and unit.endswith('2'):
            unit = unit[:-1]
        elif unit is not None and not isinstance(unit, str):
            unit = unit ** 0.5
        return self._with(v, unit)


_AXIS_CARD = None


def _is_wcs_card_above(key, ndim):
    """True for a per-axis FITS card (NAXISn, CTYPEn, CRVALn, CRPIXn, CDELTn, CUNITn, CROTAn, PCi_j, CDi_j, PVi_m)
    that refers to an axis number > ndim."""
    import re
    global _AXIS_CARD
    if _AXIS_CARD is None:
        _AXIS_CARD = (re.compile(r'^(NAXIS|CTYPE|CRVAL|CRPIX|CDELT|CUNIT|CROTA|CRDER|CSYER|CNAME)(\d+)[A-Z]?$'),
                      re.compile(r'^(PC|CD)0*(\d+)_0*(\d+)[A-Z]?$'), re.compile(r'^(PV|PS)(\d+)_\d+[A-Z]?$'))
    k = str(key).upper()
    m = _AXIS_CARD[0].match(k)
    if m:
        return int(m.group(2)) > ndim
    m = _AXIS_CARD[1].match(k)
    if m:
        return int(m.group(2)) > ndim or int(m.group(3)) > ndim
    m = _AXIS_CARD[2].match(k)
    if m:
        return int(m.group(2)) > ndim
    return k == 'WCSAXES'


class Projection(LowerDimensionalObject):
    """2-D result (moment maps).  lower_dimensional_structures.py:246-292."""

    @property
    def beam(self):
        """The beam in ``meta['beam']`` or the header's BMAJ/BMIN/BPA (lower_dimensional_structures.py:285-290);
        AttributeError when there is none, so that ``hasattr(proj, 'beam')`` works as in the reference."""
        from .beam import Beam
        bm = self._meta.get('beam') if self._meta else None
        if bm is None and self._header and 'BMAJ' in self._header:
            try:
                bm = Beam.from_fits_header(self._header)
            except Exception:
                bm = None
        if bm is None:
            raise AttributeError("this Projection has no beam")
        return Beam.coerce(bm)

    def convolve_to(self, beam, convolve=None, **kwargs):
        """Convolve the image to a specified beam (lower_dimensional_structures.py:450-494): the map goes to the
        device as a one-channel cube and through the same kernels as ``SpectralCube.convolve_to``.  No Jy/beam
        rescale here (the reference's has none); ``convolve`` defaults to ``convolve_fft`` like the reference's."""
        import warnings
        from . import _lib
        from .beam import Beam
        from .cube import BaseSpectralCube
        from .wcs import CelestialWCS, CubeWCS
        if not isinstance(self._wcs, CelestialWCS):
            raise ValueError("WCS does not contain two spatial axes.")          # utils.WCSCelestialError
        if not hasattr(self, 'beam'):
            raise ValueError("No beam is contained in Projection.meta.")
        BaseSpectralCube._check_convolve_kwargs(kwargs)
        beam = Beam.coerce(beam)
        if beam == self.beam:
            warnings.warn("The given beam is identical to the current beam. "
                          "Skipping convolution.")
            return self
        m = np.asarray(self._wcs.cdelt, dtype=np.float64)[:, None] * np.asarray(self._wcs.pc, dtype=np.float64)
        pixscale = float(np.sqrt(abs(m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0])))
        kernel = beam.deconvolve(self.beam).as_kernel(pixscale)
        fft = BaseSpectralCube._fft_semantics(convolve, default=True)
        image = np.empty((1,) + self.shape, dtype=np.float32)             # a fresh block: proper strides for the size-1 axis
        image[0] = self.value
        plane = BaseSpectralCube(image,
                                 CubeWCS(['RA---TAN', 'DEC--TAN', 'VRAD'], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [-1.0, 1.0, 1.0]))
        out = plane._run_spatial_smooth(plane._kernel_array(kernel, 2), _lib.F32, passthrough=0, every_tap=True,
                                        strategy_counts=plane._convolved_denominator_counts())
        plane._convolve_epilogue(out, 1.0, fft)
        meta = dict(self._meta)
        meta['beam'] = beam
        header = dict(self._header or {})
        header.update(beam.to_header_keywords())
        return Projection(out[0].cpu().numpy().astype(np.float64), unit=self._unit, wcs=self._wcs, meta=meta,
                          header=header, copy=False)

    def reproject(self, header, order='bilinear'):
        """Reproject the image into a new header (lower_dimensional_structures.py:496-538: `reproject_interp((value,
        header), WCS(header), shape_out=..., order=order)`): the map goes to the device as a one-channel cube and through
        `sc_wcs_pixel_map` + `sc_reproject_ex`, the kernels of `SpectralCube.reproject`.  NaN pixels poison the samples
        they touch, pixels beyond half a pixel outside the image come out NaN; the result is float64 and carries the new
        celestial WCS and the given header."""
        from . import _lib
        from .cube import BaseSpectralCube
        from .wcs import CelestialWCS, CubeWCS
        if not isinstance(self._wcs, CelestialWCS):
            raise ValueError("WCS does not contain two spatial axes.")          # utils.WCSCelestialError
        if order not in BaseSpectralCube._ORDERS:
            raise NotImplementedError("order=%r: only 'nearest-neighbor' and 'bilinear' are implemented" % (order,))
        hdr = dict(header)
        ny_out, nx_out = int(hdr['NAXIS2']), int(hdr['NAXIS1'])

        def lifted(ctype, crval, crpix, cdelt, pc, lonpole):
            pc3 = np.eye(3)
            pc3[:2, :2] = np.asarray(pc, dtype=np.float64)
            return CubeWCS(list(ctype) + ['VRAD'], list(crval) + [0.0], list(crpix) + [1.0], list(cdelt) + [1.0],
                           cunit=('deg', 'deg', 'm/s'), pc=pc3, lonpole=lonpole)
        w = self._wcs
        lift_hdr = dict(hdr)
        lift_hdr.update({'CTYPE3': 'VRAD', 'CRVAL3': 0.0, 'CRPIX3': 1.0, 'CDELT3': 1.0, 'CUNIT3': 'm/s'})
        for k in [k for k in lift_hdr if k.startswith(('PC3_', 'PC1_3', 'PC2_3', 'CD3_', 'CD1_3', 'CD2_3'))]:
            del lift_hdr[k]
        if any(k.startswith('CD') and k[2:3].isdigit() for k in lift_hdr):
            lift_hdr['CD3_3'] = 1.0
        target = CubeWCS.from_header(lift_hdr)
        image = np.empty((1,) + self.shape, dtype=np.float32)             # a fresh block: proper strides for the size-1 axis
        image[0] = self.value
        plane = BaseSpectralCube(image, lifted(w.ctype, w.crval, w.crpix, w.cdelt, w.pc, w.lonpole))
        yin, xin = plane._pixel_map(target, ny_out, nx_out)
        out, _, _, _ = plane._run_reproject(yin, xin, BaseSpectralCube._ORDERS[order], filled=False)
        return Projection(out[0].cpu().numpy(), unit=self._unit, wcs=target.celestial(), meta=self._meta,
                          header=hdr, copy=False)

    def __new__(cls, value, unit=None, wcs=None, meta=None, header=None, copy=True):
        if np.ndim(value) != 2:  # noqa
            raise ValueError("value should be a 2-d array")
        return super(Projection, cls).__new__(cls, value, unit=unit, wcs=wcs, meta=meta, header=header, copy=copy)
