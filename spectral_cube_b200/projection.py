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
        hdr = dict(self._header or {})
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


class Projection(LowerDimensionalObject):
    """2-D result (moment maps).  lower_dimensional_structures.py:246-292."""

    def __new__(cls, value, unit=None, wcs=None, meta=None, header=None, copy=True):
        if np.ndim(value) != 2:  # noqa
            raise ValueError("value should be a 2-d array")
        return super(Projection, cls).__new__(cls, value, unit=unit, wcs=wcs, meta=meta, header=header, copy=copy)
