"""
oracle/wcs.py -- TEST INFRASTRUCTURE.  Minimal FITS-WCS evaluation (float64, numpy).

Stands in for ``astropy.wcs`` (wcslib), which the reference calls at
``spectral_cube/base_class.py:227`` (``all_pix2world``) and through
``reproject_interp`` (``spectral_cube/spectral_cube.py:2726-2732``).  astropy is not
installable here, so this follows the published FITS WCS papers
(Greisen & Calabretta 2002, Paper I: linear transform; Calabretta & Greisen 2002,
Paper II: TAN / SIN zenithal projections and the spherical rotation, eqs. 2, 5,
54-55, 59).  Only what the hot path needs is implemented: a 3-axis
(lon, lat, spectral) WCS whose spectral axis is linear and uncorrelated with the
celestial axes (the reference enforces that separability, ``spectral_cube.py:1513-1515``).

wcslib normalises units on ``wcs.set()``: celestial -> deg, spectral -> SI (m/s, Hz,
m); the reference relies on that (``_spectral_scale``, ``spectral_cube.py:1526-1529``).
``OWCS.__init__`` does the same.
"""
import numpy as np

D2R = np.pi / 180.0
R2D = 180.0 / np.pi

# factor that converts a value in <unit> to the SI unit wcslib would hold internally
_TO_SI = {
    'm/s': 1.0, 'km/s': 1.0e3, 'cm/s': 1.0e-2,
    'Hz': 1.0, 'kHz': 1.0e3, 'MHz': 1.0e6, 'GHz': 1.0e9,
    'm': 1.0, 'cm': 1.0e-2, 'mm': 1.0e-3, 'um': 1.0e-6, 'nm': 1.0e-9, 'Angstrom': 1.0e-10,
    '': 1.0,
}
_SI_NAME = {
    'm/s': 'm/s', 'km/s': 'm/s', 'cm/s': 'm/s',
    'Hz': 'Hz', 'kHz': 'Hz', 'MHz': 'Hz', 'GHz': 'Hz',
    'm': 'm', 'cm': 'm', 'mm': 'm', 'um': 'm', 'nm': 'm', 'Angstrom': 'm', '': '',
}
_ANG_TO_DEG = {'deg': 1.0, 'arcmin': 1.0 / 60.0, 'arcsec': 1.0 / 3600.0, 'rad': R2D, '': 1.0}


def unit_scale(si_unit, unit):
    """Multiplicative factor taking a value in ``si_unit`` to ``unit``
    (restates ``spectral_cube/spectral_axis.py:67-73`` ``wcs_unit_scale``)."""
    if _SI_NAME[unit] != si_unit:
        raise ValueError("unit %r is not convertible from %r" % (unit, si_unit))
    return 1.0 / _TO_SI[unit]


class OWCS(object):
    """3-axis FITS WCS in FITS axis order (1=lon, 2=lat, 3=spectral), 1-based CRPIX."""

    def __init__(self, ctype, crval, crpix, cdelt, cunit=('deg', 'deg', 'm/s'), pc=None,
                 lonpole=None):
        self.ctype = list(ctype)
        crval = np.array(crval, dtype=np.float64)
        cdelt = np.array(cdelt, dtype=np.float64)
        self.crpix = np.array(crpix, dtype=np.float64)
        self.pc = np.eye(3) if pc is None else np.array(pc, dtype=np.float64)
        cunit = list(cunit)
        # unit normalisation as wcslib does on set()
        for i in (0, 1):
            s = _ANG_TO_DEG[cunit[i]]
            crval[i] *= s
            cdelt[i] *= s
            cunit[i] = 'deg'
        s = _TO_SI[cunit[2]]
        crval[2] *= s
        cdelt[2] *= s
        cunit[2] = _SI_NAME[cunit[2]]
        self.crval, self.cdelt, self.cunit = crval, cdelt, cunit
        self.proj = self.ctype[0][-3:]
        if self.proj not in ('TAN', 'SIN'):
            raise NotImplementedError("projection %s" % self.proj)
        if lonpole is None:
            # Paper II section 2.8: default LONPOLE for zenithal projections (theta0 = 90)
            lonpole = 0.0 if self.crval[1] >= 90.0 else 180.0
        self.lonpole = float(lonpole)

    # -- linear part (Paper I eq. 1) ---------------------------------------------
    @property
    def pixel_scale_matrix(self):
        """CDELT-scaled PC matrix (what ``astropy.wcs.WCS.pixel_scale_matrix`` returns)."""
        return self.cdelt[:, None] * self.pc

    def spectral_pix2world(self, pz, origin=0):
        """World value (SI unit) of spectral pixel ``pz`` (0-based when origin=0)."""
        pz = np.asarray(pz, dtype=np.float64)
        return self.crval[2] + self.cdelt[2] * self.pc[2, 2] * (pz + (1 - origin) - self.crpix[2])

    def _intermediate(self, px, py, origin=0):
        dx = np.asarray(px, dtype=np.float64) + (1 - origin) - self.crpix[0]
        dy = np.asarray(py, dtype=np.float64) + (1 - origin) - self.crpix[1]
        m = self.pixel_scale_matrix
        x = m[0, 0] * dx + m[0, 1] * dy
        y = m[1, 0] * dx + m[1, 1] * dy
        return x, y

    # -- celestial part (Paper II) ---------------------------------------------------
    def celestial_pix2world(self, px, py, origin=0):
        """(lon, lat) in degrees for pixel coordinates (px, py)."""
        x, y = self._intermediate(px, py, origin)
        xr, yr = x * D2R, y * D2R                  # radians on the projection plane
        r2 = xr * xr + yr * yr
        r = np.sqrt(r2)
        # unit vector in the native frame: (cos th sin ph, -cos th cos ph, sin th)
        if self.proj == 'TAN':                     # R = cot(theta)
            st = 1.0 / np.sqrt(1.0 + r2)           # sin(theta)
            ctsp = xr * st                         # cos(theta) sin(phi)
            ctcp = -yr * st                        # cos(theta) cos(phi)
        else:                                      # SIN: R = cos(theta)
            st = np.sqrt(np.clip(1.0 - r2, 0.0, None))
            st = np.where(r2 > 1.0, np.nan, st)
            ctsp = xr
            ctcp = -yr
        # rotate native -> celestial (Paper II eq. 2), pole at (crval), phi_p = lonpole
        php = self.lonpole * D2R
        dp = self.crval[1] * D2R
        # components relative to phi_p
        cps, sps = np.cos(php), np.sin(php)
        ct_cos_dphi = ctcp * cps + ctsp * sps      # cos th cos(phi - phi_p)
        ct_sin_dphi = ctsp * cps - ctcp * sps      # cos th sin(phi - phi_p)
        sdp, cdp = np.sin(dp), np.cos(dp)
        # celestial unit vector in a frame whose x axis points at alpha_p
        zc = st * sdp + ct_cos_dphi * cdp          # sin(delta)
        xc = st * cdp - ct_cos_dphi * sdp          # cos(delta) cos(alpha - alpha_p)
        yc = -ct_sin_dphi                          # cos(delta) sin(alpha - alpha_p)
        lon = self.crval[0] + np.arctan2(yc, xc) * R2D
        lat = np.arctan2(zc, np.hypot(xc, yc)) * R2D
        return lon, lat

    def celestial_world2pix(self, lon, lat, origin=0):
        """Pixel coordinates (px, py) of celestial positions (degrees)."""
        lon = np.asarray(lon, dtype=np.float64)
        lat = np.asarray(lat, dtype=np.float64)
        da = (lon - self.crval[0]) * D2R
        d = lat * D2R
        dp = self.crval[1] * D2R
        sd, cd = np.sin(d), np.cos(d)
        sdp, cdp = np.sin(dp), np.cos(dp)
        hav = 2.0 * np.sin(0.5 * da) ** 2          # 1 - cos(da), no cancellation
        # Paper II eq. 5 written to avoid cancellation near the pole
        st = np.cos(d - dp) - cd * cdp * hav       # sin(theta)
        ct_cos_dphi = np.sin(d - dp) + cd * sdp * hav
        ct_sin_dphi = -cd * np.sin(da)
        php = self.lonpole * D2R
        cps, sps = np.cos(php), np.sin(php)
        ctsp = ct_sin_dphi * cps + ct_cos_dphi * sps   # cos th sin(phi)
        ctcp = ct_cos_dphi * cps - ct_sin_dphi * sps   # cos th cos(phi)
        if self.proj == 'TAN':
            bad = st <= 0.0
            with np.errstate(divide='ignore', invalid='ignore'):
                xr = ctsp / st
                yr = -ctcp / st
        else:
            bad = st < 0.0
            xr = ctsp
            yr = -ctcp
        x = np.where(bad, np.nan, xr * R2D)
        y = np.where(bad, np.nan, yr * R2D)
        minv = np.linalg.inv(self.pixel_scale_matrix[:2, :2])
        dx = minv[0, 0] * x + minv[0, 1] * y
        dy = minv[1, 0] * x + minv[1, 1] * y
        px = dx + self.crpix[0] - (1 - origin)
        py = dy + self.crpix[1] - (1 - origin)
        return px, py

    def all_pix2world(self, px, py, pz, origin=0):
        px, py, pz = np.broadcast_arrays(px, py, pz)
        lon, lat = self.celestial_pix2world(px, py, origin)
        return lon, lat, self.spectral_pix2world(pz, origin)

    def copy(self):
        w = OWCS.__new__(OWCS)
        w.ctype = list(self.ctype)
        w.crval = self.crval.copy()
        w.crpix = self.crpix.copy()
        w.cdelt = self.cdelt.copy()
        w.pc = self.pc.copy()
        w.cunit = list(self.cunit)
        w.proj = self.proj
        w.lonpole = self.lonpole
        return w


def angular_separation(lon1, lat1, lon2, lat2):
    """Vincenty great-circle separation (radians in, radians out); restates
    ``astropy.coordinates.angular_separation`` used at ``spectral_cube.py:1482-1486``."""
    sdlon = np.sin(lon2 - lon1)
    cdlon = np.cos(lon2 - lon1)
    slat1, slat2 = np.sin(lat1), np.sin(lat2)
    clat1, clat2 = np.cos(lat1), np.cos(lat2)
    num1 = clat2 * sdlon
    num2 = clat1 * slat2 - slat1 * clat2 * cdlon
    denominator = slat1 * slat2 + clat1 * clat2 * cdlon
    return np.arctan2(np.hypot(num1, num2), denominator)
