"""
``mosaic_cubes`` / ``combine_headers`` (spectral_cube/cube_utils.py:744-856), SURVEY.md section 8(f)-4: reproject
several cubes onto one grid that contains them all and average where they overlap.

Device pipeline: per input cube ONE pixel map + ONE `sc_reproject_ex` pass (the kernels of `SpectralCube.reproject`),
then `sc_mosaic_accumulate` folds ``nan_to_num`` of the float64 result into the float64 accumulator and the channel-0
footprint into the coverage plane; `sc_mosaic_normalize` divides at the end.  Nothing returns to the host in between.

The common grid is what ``reproject.mosaicking.find_optimal_celestial_wcs`` builds (reproject >= 0.9.1 is a dependency
of the reference, ``pyproject.toml:41``, not installable here; its published algorithm is restated): a TAN projection
in the first cube's frame, unrotated, centred on the mean of the inputs' reference positions, at the finest input pixel
scale, sized to the extreme image corners.
"""
import warnings

import numpy as np

from . import _lib
from .wcs import CelestialWCS, as_cube_wcs


def _celestial_of(obj):
    """((ny, nx), CelestialWCS) of a cube, a header-like mapping, or an ((ny, nx), wcs) pair."""
    if isinstance(obj, tuple) and len(obj) == 2:
        shape, w = obj
        w = w if isinstance(w, CelestialWCS) else as_cube_wcs(w).celestial()
        return tuple(int(n) for n in shape[-2:]), w
    if hasattr(obj, 'get') and hasattr(obj, 'keys'):
        return (int(obj['NAXIS2']), int(obj['NAXIS1'])), as_cube_wcs(obj).celestial()
    return tuple(obj.shape[-2:]), as_cube_wcs(obj.wcs).celestial()


def find_optimal_celestial_wcs(input_data, auto_rotate=False, projection='TAN', resolution=None, reference=None):
    """(CelestialWCS, (ny, nx)) of the smallest unrotated `projection` image that contains every input.

    ``input_data``: cubes, headers or ((ny, nx), wcs) pairs.  ``resolution`` (degrees) defaults to the finest input
    pixel scale, ``reference`` ((lon, lat) in degrees) to the mean direction of the inputs' reference positions."""
    if auto_rotate:
        raise NotImplementedError("auto_rotate=True (minimum-area rotation) is not available; the reference passes False "
                                  "(cube_utils.py:775)")
    inputs = [_celestial_of(x) for x in input_data]
    corners_lon, corners_lat, refs, scales = [], [], [], []
    for (ny, nx), w in inputs:
        xc = np.array([-0.5, nx - 0.5, nx - 0.5, -0.5])
        yc = np.array([-0.5, -0.5, ny - 0.5, ny - 0.5])
        lon, lat = w.pix2world(xc, yc, origin=0)
        corners_lon.append(lon)
        corners_lat.append(lat)
        rl, rb = w.pix2world(w.crpix[0], w.crpix[1], origin=1)           # the reference pixel's position (= CRVAL)
        refs.append((float(rl), float(rb)))
        scales.append(np.min(np.sqrt((w.pixel_scale_matrix ** 2).sum(axis=0))))        # proj_plane_pixel_scales
    corners_lon, corners_lat = np.concatenate(corners_lon), np.concatenate(corners_lat)
    if reference is None:
        # mean of the unit vectors (what averaging a cartesian representation does), back to a direction
        rl, rb = np.radians([r[0] for r in refs]), np.radians([r[1] for r in refs])
        v = np.array([np.mean(np.cos(rb) * np.cos(rl)), np.mean(np.cos(rb) * np.sin(rl)), np.mean(np.sin(rb))])
        reference = (np.degrees(np.arctan2(v[1], v[0])) % 360.0, np.degrees(np.arctan2(v[2], np.hypot(v[0], v[1]))))
    cdelt = float(np.min(scales) if resolution is None else resolution)
    lonname, latname = inputs[0][1].ctype[0].split('-')[0], inputs[0][1].ctype[1].split('-')[0]
    ctype = ['%s%s' % (lonname.ljust(4, '-'), '-' + projection), '%s%s' % (latname.ljust(4, '-'), '-' + projection)]
    out = CelestialWCS(ctype, np.array(reference, dtype=np.float64), np.array([1.0, 1.0]), np.array([-cdelt, cdelt]),
                       np.eye(2), 180.0 if reference[1] < 90.0 else 0.0)
    xp, yp = out.world2pix(corners_lon, corners_lat, origin=1)
    xmin, xmax, ymin, ymax = xp.min(), xp.max(), yp.min(), yp.max()
    # the lower-left corner of the final image sits at pixel (0.5, 0.5) in FITS counting
    out.crpix = np.array([(1.0 - xmin) + 0.5, (1.0 - ymin) + 0.5])
    return out, (int(round(ymax - ymin)), int(round(xmax - xmin)))


def combine_headers(header1, header2, **kwargs):
    """cube_utils.py:744-789: a header for a field containing both inputs (the first header's cards, the common celestial
    grid, NAXIS3 of the first).  Celestial PCi_j / CDi_j cards of the first header do not survive: the grid is unrotated."""
    wcs_opt, shape_opt = find_optimal_celestial_wcs([header1, header2], auto_rotate=False, **kwargs)
    header = dict(header1)
    for key in list(header):
        k = str(key).upper()
        if (k[:2] in ('PC', 'CD') and '_' in k and k[2:].replace('_', '').isdigit() and
                all(int(n) in (1, 2) for n in k[2:].split('_'))) or k in ('CROTA1', 'CROTA2', 'LATPOLE'):
            del header[key]
    header['NAXIS'] = 3
    header['NAXIS1'] = shape_opt[1]
    header['NAXIS2'] = shape_opt[0]
    header['NAXIS3'] = header1['NAXIS3']
    header.update(wcs_opt.to_header())
    header['LONPOLE'] = wcs_opt.lonpole
    header['WCSAXES'] = 3
    return header


def mosaic_cubes(cubes, spectral_block_size=100, combine_header_kwargs={}, **kwargs):
    """Reproject ``cubes`` onto a common grid and average them where they overlap (cube_utils.py:791-856).

    ``spectral_block_size`` is accepted for drop-in compatibility (the reference uses it to bound reproject's host
    memory; the device pass needs no blocking).  ``kwargs`` go to ``SpectralCube.reproject`` (``order=...``)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    cubes = list(cubes)
    cube1 = cubes[0]
    header = cube1.header
    for cu in cubes[1:]:
        header = combine_headers(header, cu.header, **combine_header_kwargs)
    nchan, ny, nx = int(header['NAXIS3']), int(header['NAXIS2']), int(header['NAXIS1'])
    dev = cube1._data.device
    kwargs.pop('block_size', None)
    kwargs.pop('roundtrip_coords', None)                       # reproject_interp's round-trip check has no device analogue
    with _lib.on_device_of(cube1._data):
        stream = torch.cuda.current_stream().cuda_stream
        final = torch.zeros((nchan, ny, nx), dtype=torch.float64, device=dev)          # np.zeros(shape_opt), :820
        weight = torch.zeros((ny, nx), dtype=torch.float64, device=dev)                # mask_opt, :821
        for cube in cubes:
            huge = cube.allow_huge_operations
            cube.allow_huge_operations = True                   # the reference blocks the spectral axis instead (:826-831)
            try:
                rep = cube.reproject(header, **kwargs)
            finally:
                cube.allow_huge_operations = huge
            hi = rep._data_hi                                   # float64, NaN outside the footprint (= the filled data)
            foot = rep._mask._mask                              # uint8 (nchan, ny, nx): reproject's footprint
            _lib.check(lib.sc_mosaic_accumulate(final.data_ptr(), weight.data_ptr(), hi.data_ptr(), _lib.F64,
                                                foot.data_ptr(), nchan, ny, nx, stream))
            del rep, hi, foot
        _lib.check(lib.sc_mosaic_normalize(final.data_ptr(), weight.data_ptr(), nchan, ny, nx, stream))
        hdr = dict((k, v) for k, v in header.items())
        result = type(cube1)(final.to(torch.float32), as_cube_wcs(header), unit=cube1.unit, header=hdr,
                             allow_huge_operations=cube1.allow_huge_operations)
        result._data_hi = final                                 # the reference's cube holds the float64 array (:855)
    return result
