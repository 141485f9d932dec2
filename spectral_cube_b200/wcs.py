"""
Host-side coordinate metadata for the hot path.

The reference leans on ``astropy.wcs`` (``base_class.py:178-241``, ``spectral_cube.py:1455-1535``).
astropy is not a dependency of this package; what the hot path needs from a WCS is small and
O(nchan): the spectral world value of every channel, the pixel scale matrix, and -- for
``reproject`` -- the celestial parameters that the device-side pixel map
(``sc_wcs_pixel_map``) consumes.  ``CubeWCS`` holds exactly that, in FITS conventions
(1-based CRPIX, FITS axis order lon/lat/spectral), with wcslib's unit normalisation
(celestial -> deg, spectral -> SI).  ``CubeWCS.from_astropy`` adapts a real
``astropy.wcs.WCS`` when astropy is present.
"""
import copy as _copy

import numpy as np

_SI = {'m/s': ('m/s', 1.0), 'km/s': ('m/s', 1.0e3), 'cm/s': ('m/s', 1.0e-2),
       'Hz': ('Hz', 1.0), 'kHz': ('Hz', 1.0e3), 'MHz': ('Hz', 1.0e6), 'GHz': ('Hz', 1.0e9),
       'm': ('m', 1.0), 'cm': ('m', 1.0e-2), 'mm': ('m', 1.0e-3), 'um': ('m', 1.0e-6),
       'nm': ('m', 1.0e-9), 'Angstrom': ('m', 1.0e-10), '': ('', 1.0)}
_DEG = {'deg': 1.0, 'arcmin': 1.0 / 60.0, 'arcsec': 1.0 / 3600.0, 'rad': 180.0 / np.pi, '': 1.0}


_SI_FOLDED = dict((k.lower(), k) for k in _SI)
_SI_FOLDED.update({'angstrom': 'Angstrom', 'm s-1': 'm/s', 'km s-1': 'km/s', 'm.s-1': 'm/s', 'km.s-1': 'km/s',
                   'micron': 'um', 'a': 'Angstrom'})
_DEG_FOLDED = dict((k.lower(), k) for k in _DEG)
_DEG_FOLDED.update({'degree': 'deg', 'degrees': 'deg', 'radian': 'rad', 'arcsecond': 'arcsec', 'arcminute': 'arcmin'})


def _normalise_unit(unit, table, folded, what):
    """wcslib's `wcsutrn`-style tolerance for CUNIT strings: archive files carry 'HZ', 'M/S', 'DEG', 'KM/S' ..."""
    u = str(unit).strip()
    if u in table:
        return u
    k = folded.get(u.lower())
    if k is None:
        raise ValueError("unknown %s unit %r (known: %s)" % (what, unit, ', '.join(sorted(x for x in table if x))))
    return k


def spectral_unit_scale(si_unit, unit):
    """Factor taking a value in the WCS's SI unit to ``unit`` (spectral_axis.py:67-73)."""
    unit = str(unit)
    unit = _SI_FOLDED.get(unit.strip().lower(), unit) if unit not in _SI else unit
    if unit not in _SI or _SI[unit][0] != si_unit:
        raise ValueError("unit %r is not convertible from %r" % (unit, si_unit))
    return 1.0 / _SI[unit][1]


class CubeWCS(object):
    def __init__(self, ctype, crval, crpix, cdelt, cunit=('deg', 'deg', 'm/s'), pc=None, lonpole=None):
        self.ctype = [str(c) for c in ctype]
        crval = np.array(crval, dtype=np.float64)
        cdelt = np.array(cdelt, dtype=np.float64)
        cunit = [str(c) for c in cunit]
        for i in (0, 1):
            f = _DEG[_normalise_unit(cunit[i], _DEG, _DEG_FOLDED, 'celestial')]
            crval[i] *= f
            cdelt[i] *= f
            cunit[i] = 'deg'
        name, f = _SI[_normalise_unit(cunit[2], _SI, _SI_FOLDED, 'spectral')]
        crval[2] *= f
        cdelt[2] *= f
        cunit[2] = name
        self.crval, self.cdelt, self.cunit = crval, cdelt, cunit
        self.crpix = np.array(crpix, dtype=np.float64)
        self.pc = np.eye(3) if pc is None else np.array(pc, dtype=np.float64)
        # moments, reductions, smoothing and convolve_to never evaluate the projection: any cube can be built and
        # read; only `reproject` (celestial_params) needs the closed forms the device pixel map implements
        self.proj = self.ctype[0][-3:]
        if lonpole is None:
            lonpole = 0.0 if self.crval[1] >= 90.0 else 180.0        # FITS paper II, zenithal default
        self.lonpole = float(lonpole)

    # -- construction helpers -----------------------------------------------------------------
    @classmethod
    def from_header(cls, hdr):
        """Build from a FITS-header-like mapping (CTYPEn, CRVALn, CRPIXn, CDELTn, CUNITn, PCi_j / CDi_j)."""
        g = hdr.get
        ctype = [g('CTYPE%d' % i) for i in (1, 2, 3)]
        crval = [g('CRVAL%d' % i, 0.0) for i in (1, 2, 3)]
        crpix = [g('CRPIX%d' % i, 0.0) for i in (1, 2, 3)]
        cunit = [g('CUNIT%d' % i, 'deg' if i < 3 else '') for i in (1, 2, 3)]
        if not str(cunit[2]).strip():
            # a blank CUNIT means the SI unit of the axis type (FITS paper III; what wcslib assumes)
            base = str(ctype[2] or '')[:4].upper()
            cunit[2] = {'VRAD': 'm/s', 'VOPT': 'm/s', 'VELO': 'm/s', 'FELO': 'm/s', 'FREQ': 'Hz', 'WAVE': 'm', 'AWAV': 'm'}.get(base, '')
        if any(('CD%d_%d' % (i, j)) in hdr for i in (1, 2, 3) for j in (1, 2, 3)):
            cd = np.array([[g('CD%d_%d' % (i, j), 0.0) for j in (1, 2, 3)] for i in (1, 2, 3)], dtype=np.float64)
            cdelt = [1.0, 1.0, 1.0]
            pc = cd
        else:
            cdelt = [g('CDELT%d' % i, 1.0) for i in (1, 2, 3)]
            pc = np.array([[g('PC%d_%d' % (i, j), 1.0 if i == j else 0.0) for j in (1, 2, 3)]
                           for i in (1, 2, 3)], dtype=np.float64)
        return cls(ctype, crval, crpix, cdelt, cunit, pc=pc, lonpole=g('LONPOLE', None))

    @classmethod
    def from_astropy(cls, w):
        ww = w.wcs
        pc = ww.get_pc() if ww.has_pc() or not ww.has_cd() else ww.cd
        cdelt = ww.cdelt if not ww.has_cd() else [1.0, 1.0, 1.0]
        return cls(list(ww.ctype), list(ww.crval), list(ww.crpix), list(cdelt),
                   [str(c) for c in ww.cunit], pc=np.array(pc), lonpole=float(ww.lonpole) if np.isfinite(ww.lonpole) else None)

    def copy(self):
        return _copy.deepcopy(self)

    # -- what the hot path needs ----------------------------------------------------------------
    @property
    def pixel_scale_matrix(self):
        return self.cdelt[:, None] * self.pc

    def spectral_pix2world(self, pz):
        """SI world value of 0-based spectral pixel(s) ``pz`` (linear axis)."""
        pz = np.asarray(pz, dtype=np.float64)
        return self.crval[2] + self.cdelt[2] * self.pc[2, 2] * (pz + 1.0 - self.crpix[2])

    def celestial_params(self):
        """The 12 doubles ``sc_wcs_pixel_map`` takes."""
        if self.proj not in ('TAN', 'SIN'):
            raise NotImplementedError("celestial projection %r: the device pixel map implements TAN and SIN" % self.proj)
        m = self.pixel_scale_matrix
        return np.array([self.crpix[0], self.crpix[1], self.crval[0], self.crval[1],
                         m[0, 0], m[0, 1], m[1, 0], m[1, 1], self.lonpole,
                         0.0 if self.proj == 'TAN' else 1.0, 0.0, 0.0], dtype=np.float64)

    def celestial_pix2world_ref(self):
        """(lon, lat) of the reference pixel -- by construction CRVAL."""
        return self.crval[0], self.crval[1]

    def to_header(self):
        hdr = {}
        for i in range(3):
            n = i + 1
            hdr['CTYPE%d' % n] = self.ctype[i]
            hdr['CRVAL%d' % n] = float(self.crval[i])
            hdr['CRPIX%d' % n] = float(self.crpix[i])
            hdr['CDELT%d' % n] = float(self.cdelt[i])
            hdr['CUNIT%d' % n] = self.cunit[i]
            for j in range(3):
                if self.pc[i, j] != (1.0 if i == j else 0.0):
                    hdr['PC%d_%d' % (n, j + 1)] = float(self.pc[i, j])
        hdr['LONPOLE'] = self.lonpole
        return hdr

    def celestial(self):
        """2-D WCS description left after dropping the spectral axis (wcs_utils.py:28-45)."""
        return CelestialWCS(self.ctype[:2], self.crval[:2].copy(), self.crpix[:2].copy(),
                            self.cdelt[:2].copy(), self.pc[:2, :2].copy(), self.lonpole)

    def drop_axis(self, np_axis):
        """WCS of a projection along numpy axis ``np_axis`` (0 spectral, 1 lat, 2 lon)."""
        if np_axis == 0:
            return self.celestial()
        keep = [i for i in range(3) if i != 2 - np_axis]
        return AxesWCS([self.ctype[i] for i in keep], self.crval[keep].copy(), self.crpix[keep].copy(),
                       self.cdelt[keep].copy(), [self.cunit[i] for i in keep])


def _zenithal_pix2world(crpix, crval, psm, lonpole, proj, px, py, origin=0):
    """(lon, lat) in degrees of pixel positions for a zenithal TAN / SIN projection (FITS WCS paper II, sections 2, 5.1):
    pixel -> intermediate world coordinates -> native spherical (phi, theta) -> celestial, via unit vectors.  Host-side
    and meant for a handful of points (image corners); the per-pixel map runs on the device (`sc_wcs_pixel_map`)."""
    if proj not in ('TAN', 'SIN'):
        raise NotImplementedError("celestial projection %r: TAN and SIN are implemented" % proj)
    d2r = np.pi / 180.0
    dx = np.asarray(px, dtype=np.float64) + (1 - origin) - crpix[0]
    dy = np.asarray(py, dtype=np.float64) + (1 - origin) - crpix[1]
    x = (psm[0, 0] * dx + psm[0, 1] * dy) * d2r
    y = (psm[1, 0] * dx + psm[1, 1] * dy) * d2r
    rr = x * x + y * y
    if proj == 'TAN':                               # R_theta = cot(theta)
        sin_t = 1.0 / np.sqrt(1.0 + rr)
        n = np.stack([x * sin_t, -y * sin_t, sin_t])            # (cos t sin p, cos t cos p, sin t) in the native frame
    else:                                           # R_theta = cos(theta)
        with np.errstate(invalid='ignore'):
            n = np.stack([x, -y, np.sqrt(1.0 - rr)])
    # native -> celestial: the native pole (theta = 90 deg) sits at CRVAL, the celestial pole at native longitude LONPOLE
    pp, dp = lonpole * d2r, crval[1] * d2r
    a = n[1] * np.cos(pp) + n[0] * np.sin(pp)       # cos t cos(p - pp)
    b = n[0] * np.cos(pp) - n[1] * np.sin(pp)       # cos t sin(p - pp)
    zc = n[2] * np.sin(dp) + a * np.cos(dp)
    xc = n[2] * np.cos(dp) - a * np.sin(dp)
    lon = crval[0] + np.degrees(np.arctan2(-b, xc))
    lat = np.degrees(np.arctan2(zc, np.hypot(xc, b)))
    return lon, lat


def _zenithal_world2pix(crpix, crval, psm, lonpole, proj, lon, lat, origin=0):
    """Inverse of `_zenithal_pix2world`; positions behind the projection's horizon come back as NaN."""
    if proj not in ('TAN', 'SIN'):
        raise NotImplementedError("celestial projection %r: TAN and SIN are implemented" % proj)
    d2r = np.pi / 180.0
    da = (np.asarray(lon, dtype=np.float64) - crval[0]) * d2r
    d, dp, pp = np.asarray(lat, dtype=np.float64) * d2r, crval[1] * d2r, lonpole * d2r
    sin_t = np.sin(d) * np.sin(dp) + np.cos(d) * np.cos(dp) * np.cos(da)
    a = np.sin(d) * np.cos(dp) - np.cos(d) * np.sin(dp) * np.cos(da)       # cos t cos(p - pp)
    b = -np.cos(d) * np.sin(da)                                               # cos t sin(p - pp)
    nx_ = b * np.cos(pp) + a * np.sin(pp)           # cos t sin p
    ny_ = a * np.cos(pp) - b * np.sin(pp)           # cos t cos p
    with np.errstate(divide='ignore', invalid='ignore'):
        if proj == 'TAN':
            x, y = np.where(sin_t > 0, nx_ / sin_t, np.nan), np.where(sin_t > 0, -ny_ / sin_t, np.nan)
        else:
            x, y = np.where(sin_t >= 0, nx_, np.nan), np.where(sin_t >= 0, -ny_, np.nan)
    inv = np.linalg.inv(np.asarray(psm, dtype=np.float64)[:2, :2])
    xd, yd = np.degrees(x), np.degrees(y)
    return (inv[0, 0] * xd + inv[0, 1] * yd + crpix[0] - (1 - origin),
            inv[1, 0] * xd + inv[1, 1] * yd + crpix[1] - (1 - origin))


class CelestialWCS(object):
    def __init__(self, ctype, crval, crpix, cdelt, pc, lonpole):
        self.ctype, self.crval, self.crpix, self.cdelt, self.pc, self.lonpole = ctype, crval, crpix, cdelt, pc, lonpole
        self.naxis = 2

    @property
    def pixel_scale_matrix(self):
        return np.asarray(self.cdelt, dtype=np.float64)[:, None] * np.asarray(self.pc, dtype=np.float64)

    def pix2world(self, px, py, origin=0):
        return _zenithal_pix2world(self.crpix, self.crval, self.pixel_scale_matrix, self.lonpole, self.ctype[0][-3:], px, py, origin)

    def world2pix(self, lon, lat, origin=0):
        return _zenithal_world2pix(self.crpix, self.crval, self.pixel_scale_matrix, self.lonpole, self.ctype[0][-3:], lon, lat, origin)

    def to_header(self):
        hdr = {}
        for i in range(2):
            n = i + 1
            hdr['CTYPE%d' % n] = self.ctype[i]
            hdr['CRVAL%d' % n] = float(self.crval[i])
            hdr['CRPIX%d' % n] = float(self.crpix[i])
            hdr['CDELT%d' % n] = float(self.cdelt[i])
            hdr['CUNIT%d' % n] = 'deg'
            for j in range(2):
                if float(np.asarray(self.pc)[i, j]) != (1.0 if i == j else 0.0):
                    hdr['PC%d_%d' % (n, j + 1)] = float(np.asarray(self.pc)[i, j])
        if self.lonpole is not None:
            hdr['LONPOLE'] = float(self.lonpole)
        return hdr


class AxesWCS(object):
    def __init__(self, ctype, crval, crpix, cdelt, cunit):
        self.ctype, self.crval, self.crpix, self.cdelt, self.cunit = ctype, crval, crpix, cdelt, cunit
        self.naxis = len(ctype)

    def to_header(self):
        hdr = {}
        for i in range(self.naxis):
            n = i + 1
            hdr['CTYPE%d' % n] = self.ctype[i]
            hdr['CRVAL%d' % n] = float(self.crval[i])
            hdr['CRPIX%d' % n] = float(self.crpix[i])
            hdr['CDELT%d' % n] = float(self.cdelt[i])
            hdr['CUNIT%d' % n] = self.cunit[i]
        return hdr


def as_cube_wcs(w):
    if isinstance(w, CubeWCS):
        return w
    if hasattr(w, 'wcs') and hasattr(w.wcs, 'crval'):
        return CubeWCS.from_astropy(w)
    if hasattr(w, 'get') and hasattr(w, 'keys'):
        return CubeWCS.from_header(w)
    # duck-typed (e.g. the test oracle's WCS object): same attribute names
    if all(hasattr(w, a) for a in ('ctype', 'crval', 'crpix', 'cdelt', 'cunit', 'pc')):
        out = CubeWCS(list(w.ctype), np.array(w.crval), np.array(w.crpix), np.array(w.cdelt),
                      list(w.cunit), pc=np.array(w.pc), lonpole=getattr(w, 'lonpole', None))
        return out
    raise TypeError("cannot interpret %r as a cube WCS" % (type(w),))
