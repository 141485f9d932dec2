"""
Synthesised-beam bookkeeping for ``convolve_to``.

The reference takes ``radio_beam.Beam`` objects (``spectral_cube.py:3335-3392`` for one beam,
``:4127-4240`` for per-channel beams) and only uses four things of them: equality, ``sr``,
``deconvolve`` and ``as_kernel(pixscale)``.  radio_beam (>=0.3.5, ``pyproject.toml:32``) and astropy
are not dependencies here, so this module carries those four with the same names and the
package's published arithmetic:

  * ``Beam.__eq__``      axes within 1e-12 deg, position angle modulo 180 deg, ignored for round beams;
  * ``Beam.sr``          2 pi sigma_maj sigma_min = pi / (4 ln 2) * major * minor;
  * ``Beam.deconvolve``  the closed-form Gaussian deconvolution (radio_beam ``utils.deconvolve_optimized``,
                         after Wild 1970 / AIPS ``DECONV``), ``BeamError("Beam could not be deconvolved")``
                         when the target is not larger than the cube's beam in every direction;
  * ``Beam.as_kernel``   ``EllipticalGaussian2DKernel(sigma_maj, sigma_min, 90 deg + pa)`` in pixels: amplitude
                         1 / (2 pi sigma_maj sigma_min), default size 8 x the larger axis-aligned full width of
                         the 1-sigma ellipse (2 x ``astropy.modeling.utils.ellipse_extent``, i.e. 16 sigma for
                         a round beam) rounded up to odd, sampled at integer pixel offsets.

Angles are plain floats in degrees (``Beam(major, minor, pa)``; ``Beam.from_arcsec`` for arcseconds); objects
with a ``.to_value`` method (astropy Quantities, a real ``radio_beam.Beam``'s attributes) are converted.  Only
host-side parameter arithmetic lives here: the convolution itself is ``sc_spatial_smooth_*`` on the device.
"""
import math

import numpy as np

from .kernels import Kernel2D, _odd_ceil, _centred_axis

SIGMA_TO_FWHM = math.sqrt(8.0 * math.log(2.0))
_ARCSEC = 1.0 / 3600.0


class BeamError(ValueError):
    """radio_beam's ``BeamError`` (a ValueError there as well: the reference catches ValueError, :4201)."""


class NoBeamError(Exception):
    """utils.py: raised by ``cube.beam`` when the cube has none."""


def _deg(value, default=None):
    if value is None:
        return default
    if hasattr(value, 'to_value'):
        return float(value.to_value('deg'))
    return float(value)


def ellipse_extent(a, b, theta):
    """Half-extent along x and y of an ellipse with semi-axes a, b rotated by theta (radians, CCW from +x)."""
    t = math.atan2(-b * math.tan(theta), a)
    dx = a * math.cos(t) * math.cos(theta) - b * math.sin(t) * math.sin(theta)
    t = math.atan2(b, a * math.tan(theta))
    dy = b * math.sin(t) * math.cos(theta) + a * math.cos(t) * math.sin(theta)
    return abs(dx), abs(dy)


class EllipticalGaussian2DKernel(Kernel2D):
    """radio_beam's kernel class: a rotated Gaussian2D model of unit integral sampled at pixel centres
    (NOT renormalised to unit sum: ``convolve(..., normalize_kernel=True)`` does that)."""

    def __init__(self, stddev_maj, stddev_min, position_angle, support_scaling=8, x_size=None, y_size=None):
        if x_size is None:
            x_size = _odd_ceil(support_scaling * 2 * max(ellipse_extent(stddev_maj, stddev_min, position_angle)))
        if y_size is None:
            y_size = x_size
        x_size, y_size = int(x_size), int(y_size)
        yy, xx = np.meshgrid(_centred_axis(y_size), _centred_axis(x_size), indexing='ij')
        ct2, st2, s2t = math.cos(position_angle) ** 2, math.sin(position_angle) ** 2, math.sin(2.0 * position_angle)
        xs2, ys2 = stddev_maj ** 2, stddev_min ** 2
        a = 0.5 * (ct2 / xs2 + st2 / ys2)
        b = 0.5 * (s2t / xs2 - s2t / ys2)
        c = 0.5 * (st2 / xs2 + ct2 / ys2)
        g = np.exp(-(a * xx * xx + b * xx * yy + c * yy * yy)) / (2.0 * math.pi * stddev_maj * stddev_min)
        super(EllipticalGaussian2DKernel, self).__init__(g)
        self.stddev_maj, self.stddev_min, self.position_angle = stddev_maj, stddev_min, position_angle


class Beam(object):
    """A 2-D Gaussian resolution element: FWHM ``major`` >= ``minor`` and position angle ``pa`` (of the major
    axis, counter-clockwise from north), all in degrees."""

    def __init__(self, major=None, minor=None, pa=None):
        major = _deg(major)
        if major is None:
            raise ValueError("Beam requires a major axis")
        minor = _deg(minor, major)
        pa = _deg(pa, 0.0)
        if minor > major:
            raise ValueError("Minor axis greater than major axis.")
        self.major, self.minor, self.pa = major, minor, pa

    @classmethod
    def from_arcsec(cls, major, minor=None, pa=0.0):
        """``Beam(major*u.arcsec, minor*u.arcsec, pa*u.deg)``"""
        return cls(major * _ARCSEC, None if minor is None else minor * _ARCSEC, pa)

    @classmethod
    def from_fits_header(cls, hdr):
        """BMAJ / BMIN / BPA in degrees (cube_utils.try_load_beam)."""
        if 'BMAJ' not in hdr:
            raise NoBeamError("No BMAJ found in the header")
        return cls(float(hdr['BMAJ']), float(hdr.get('BMIN', hdr['BMAJ'])), float(hdr.get('BPA', 0.0)))

    @classmethod
    def coerce(cls, obj):
        """This class, or a duck-typed beam (a real ``radio_beam.Beam``: .major/.minor/.pa Quantities)."""
        if isinstance(obj, cls):
            return obj
        if all(hasattr(obj, n) for n in ('major', 'minor', 'pa')):
            return cls(obj.major, obj.minor, obj.pa)
        raise TypeError("beam must be a radio_beam.Beam object.")

    def to_header_keywords(self):
        return {'BMAJ': self.major, 'BMIN': self.minor, 'BPA': self.pa}

    def __repr__(self):
        return "Beam: BMAJ=%.9g arcsec BMIN=%.9g arcsec BPA=%.9g deg" % (self.major * 3600.0, self.minor * 3600.0, self.pa)

    @property
    def isfinite(self):
        return bool(math.isfinite(self.major) and math.isfinite(self.minor) and math.isfinite(self.pa)
                    and self.major > 0 and self.minor > 0)

    def iscircular(self, rtol=1e-6):
        return (self.major - self.minor) / self.major < rtol

    def __eq__(self, other):
        if not isinstance(other, Beam):
            try:
                other = Beam.coerce(other)
            except TypeError:
                return NotImplemented
        atol = 1e-12
        same_pa = True if self.iscircular() else abs(self.pa % 180.0 - other.pa % 180.0) < atol
        return bool(abs(self.major - other.major) < atol and abs(self.minor - other.minor) < atol and same_pa)

    def __ne__(self, other):
        res = self.__eq__(other)
        return res if res is NotImplemented else not res

    __hash__ = None

    @property
    def sr(self):
        """Beam solid angle in steradians."""
        return math.pi / (4.0 * math.log(2.0)) * math.radians(self.major) * math.radians(self.minor)

    def deconvolve(self, other, failure_returns_pointlike=False):
        """The beam that, convolved with ``other``, gives this one."""
        other = Beam.coerce(other)
        maj1, min1, pa1 = self.major, self.minor, math.radians(self.pa)
        maj2, min2, pa2 = other.major, other.minor, math.radians(other.pa)
        c1, s1, c2, s2 = math.cos(pa1), math.sin(pa1), math.cos(pa2), math.sin(pa2)
        alpha = (maj1 * c1) ** 2 + (min1 * s1) ** 2 - (maj2 * c2) ** 2 - (min2 * s2) ** 2
        beta = (maj1 * s1) ** 2 + (min1 * c1) ** 2 - (maj2 * s2) ** 2 - (min2 * c2) ** 2
        gamma = 2.0 * ((min1 ** 2 - maj1 ** 2) * s1 * c1 - (min2 ** 2 - maj2 ** 2) * s2 * c2)
        s = alpha + beta
        t = math.sqrt((alpha - beta) ** 2 + gamma ** 2)
        eps = np.finfo(np.float64).eps
        # the target has to exceed the other beam along every direction (axes in degrees, squared)
        if alpha + eps < 0 or beta + eps < 0 or s < t + eps:
            if failure_returns_pointlike:
                return _PointLike()
            raise BeamError("Beam could not be deconvolved")
        new_major = math.sqrt(0.5 * (s + t))
        new_minor = math.sqrt(max(0.5 * (s - t), 0.0))
        if abs(gamma) + abs(alpha - beta) < 1e-7 / 3600.0:
            new_pa = 0.0
        else:
            new_pa = 0.5 * math.atan2(-gamma, alpha - beta)
        return Beam(new_major, new_minor, math.degrees(new_pa))

    def as_kernel(self, pixscale, **kwargs):
        """Elliptical Gaussian kernel on a grid of ``pixscale`` degrees per pixel (not aware of any rotation
        between pixel and sky axes, like the reference's)."""
        pixscale = _deg(pixscale)
        stddev_maj = self.major / (pixscale * SIGMA_TO_FWHM)
        stddev_min = self.minor / (pixscale * SIGMA_TO_FWHM)
        # pa runs counter-clockwise from north, the model angle from the +x axis
        return EllipticalGaussian2DKernel(stddev_maj, stddev_min, math.radians(90.0 + self.pa), **kwargs)


class _PointLike(Beam):
    def __init__(self):
        self.major = self.minor = self.pa = 0.0


class Beams(list):
    """Per-channel beams (radio_beam.Beams): a list of ``Beam`` with array views of the parameters."""

    def __init__(self, beams=(), major=None, minor=None, pa=None):
        if major is not None:
            major = np.atleast_1d(np.asarray(major, dtype=np.float64))
            minor = major if minor is None else np.atleast_1d(np.asarray(minor, dtype=np.float64))
            pa = np.zeros_like(major) if pa is None else np.atleast_1d(np.asarray(pa, dtype=np.float64))
            beams = [_unchecked(a, b, p) for a, b, p in zip(major, minor, pa)]
        super(Beams, self).__init__(Beam.coerce(b) for b in beams)

    @classmethod
    def from_arcsec(cls, major, minor=None, pa=None):
        major = np.asarray(major, dtype=np.float64) * _ARCSEC
        minor = None if minor is None else np.asarray(minor, dtype=np.float64) * _ARCSEC
        return cls(major=major, minor=minor, pa=pa)

    @property
    def major(self):
        return np.array([b.major for b in self])

    @property
    def minor(self):
        return np.array([b.minor for b in self])

    @property
    def pa(self):
        return np.array([b.pa for b in self])

    @property
    def sr(self):
        return np.array([b.sr for b in self])

    @property
    def isfinite(self):
        return np.array([b.isfinite for b in self], dtype=bool)

    @property
    def size(self):
        return len(self)

    def __getitem__(self, view):
        if isinstance(view, (int, np.integer)):
            return list.__getitem__(self, int(view))
        if isinstance(view, slice):
            return Beams(list.__getitem__(self, view))
        view = np.asarray(view)
        idx = np.flatnonzero(view) if view.dtype == bool else view
        return Beams([list.__getitem__(self, int(i)) for i in idx])


def _unchecked(major, minor, pa):
    """Table rows may hold NaN / non-positive placeholders for bad channels: keep them, `isfinite` flags them."""
    b = Beam.__new__(Beam)
    b.major, b.minor, b.pa = float(major), float(minor), float(pa)
    return b
