"""
FITS ingest / egress.  CPU part: the header parser and writer round-trip without a
GPU.  GPU part (``-m gpu``): `SpectralCube.read` decodes the big-endian data block on the device
bit-exactly like numpy's ``'>f4'`` view, applies BSCALE / BZERO / BLANK for integer BITPIX like
astropy.io.fits does, attaches the isfinite LazyMask (io/fits.py:214), and a moment map written with
`Projection.write` reads back unchanged.
"""
import numpy as np
import pytest

from spectral_cube_b200 import io_fits

HDR = {'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN', 'CTYPE3': 'VRAD', 'CRVAL1': 24.0, 'CRVAL2': 30.0,
       'CRVAL3': -321214.698632, 'CRPIX1': 8.5, 'CRPIX2': 8.5, 'CRPIX3': 1.0, 'CDELT1': -5.55555561268e-4,
       'CDELT2': 5.55555561268e-4, 'CDELT3': 1288.21496879, 'CUNIT1': 'deg', 'CUNIT2': 'deg', 'CUNIT3': 'm/s',
       'BUNIT': 'K', 'OBJECT': "it's a cube"}


def _cube(shape=(5, 16, 20), seed=3):
    rng = np.random.default_rng(seed)
    d = rng.normal(0, 1, shape).astype(np.float32)
    d[rng.random(shape) < 0.05] = np.nan
    return d


def test_header_round_trip(tmp_path):
    path = str(tmp_path / 'c.fits')
    d = _cube()
    io_fits.write_fits(path, d, HDR)
    assert (tmp_path / 'c.fits').stat().st_size % 2880 == 0
    with open(path, 'rb') as f:
        hdr, off = io_fits.read_header(f)
        raw = f.read(d.size * 4)
    assert off % 2880 == 0 and hdr['BITPIX'] == -32 and hdr['NAXIS'] == 3
    assert (hdr['NAXIS1'], hdr['NAXIS2'], hdr['NAXIS3']) == (20, 16, 5)
    for k, v in HDR.items():
        assert hdr[k] == v, k
    back = np.frombuffer(raw, dtype='>f4').reshape(d.shape)
    assert np.array_equal(back.astype(np.float32).view(np.uint32), d.view(np.uint32))
    with pytest.raises(OSError):
        io_fits.write_fits(path, d, HDR)


def test_float64_images_are_written_as_bitpix_minus_64(tmp_path):
    path = str(tmp_path / 'm.fits')
    m = np.arange(12, dtype=np.float64).reshape(3, 4) / 7
    io_fits.write_fits(path, m, {'BUNIT': 'K km/s'})
    with open(path, 'rb') as f:
        hdr, off = io_fits.read_header(f)
        back = np.frombuffer(f.read(m.size * 8), dtype='>f8').reshape(m.shape)
    assert hdr['BITPIX'] == -64 and np.array_equal(back, m)


@pytest.mark.gpu
def test_read_decodes_on_the_device_bit_exactly(tmp_path):
    import spectral_cube_b200 as scb
    path = str(tmp_path / 'c.fits')
    d = _cube((7, 33, 50), seed=9)               # 11550 samples: exercises the vector body and the scalar tail
    io_fits.write_fits(path, d, HDR)
    for use_dask in (False, True):
        cube = scb.SpectralCube.read(path, use_dask=use_dask)
        assert type(cube) is (scb.DaskSpectralCube if use_dask else scb.SpectralCube)
        assert cube.shape == d.shape and cube.unit == 'K' and cube.meta['BUNIT'] == 'K'
        got = cube._data.cpu().numpy()
        assert np.array_equal(got.view(np.uint32), d.view(np.uint32))
        assert np.array_equal(cube.mask.include(), np.isfinite(d))            # io/fits.py:214
        assert cube.wcs.crval[2] == HDR['CRVAL3'] and cube.wcs.ctype[0] == 'RA---TAN'


@pytest.mark.gpu
def test_read_streams_large_files_in_blocks(tmp_path):
    path = str(tmp_path / 'big.fits')
    rng = np.random.default_rng(1)
    d = rng.normal(0, 1, (9, 64, 96)).astype(np.float32)
    io_fits.write_fits(path, d, HDR)
    out, hdr = io_fits.read_fits_to_device(path, chunk_bytes=40000)           # six blocks through two buffers
    assert np.array_equal(out.cpu().numpy().view(np.uint32), d.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize('bitpix,dtype', [(16, '>i2'), (32, '>i4'), (-64, '>f8'), (8, 'u1')])
def test_scaled_and_blanked_integer_images(tmp_path, bitpix, dtype):
    """What astropy.io.fits applies on read: physical = BZERO + BSCALE * stored, stored == BLANK -> NaN."""
    path = str(tmp_path / 'i.fits')
    rng = np.random.default_rng(bitpix % 7)
    shape = (3, 10, 14)
    if bitpix > 0:
        lo, hi = (0, 200) if bitpix == 8 else (-3000, 3000)
        stored = rng.integers(lo, hi, shape)
        stored[0, 0, :5] = 123
    else:
        stored = rng.normal(0, 1, shape)
    cards = [io_fits._format_card('SIMPLE', True), io_fits._format_card('BITPIX', bitpix), io_fits._format_card('NAXIS', 3),
             io_fits._format_card('NAXIS1', 14), io_fits._format_card('NAXIS2', 10), io_fits._format_card('NAXIS3', 3),
             io_fits._format_card('BSCALE', 0.25), io_fits._format_card('BZERO', 100.0)]
    if bitpix > 0:
        cards.append(io_fits._format_card('BLANK', 123))
    cards += [io_fits._format_card(k, v) for k, v in HDR.items()] + ['END'.ljust(80)]
    text = ''.join(cards)
    text += ' ' * (-len(text) % 2880)
    raw = np.ascontiguousarray(stored.astype(dtype)).tobytes()
    with open(path, 'wb') as f:
        f.write(text.encode('ascii') + raw + b'\0' * (-len(raw) % 2880))
    out, hdr = io_fits.read_fits_to_device(path)
    want = 100.0 + 0.25 * stored.astype(np.float64)
    if bitpix > 0:
        want[stored == 123] = np.nan
    got = out.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_array_equal(got[~np.isnan(got)], want[~np.isnan(want)].astype(np.float32))


@pytest.mark.gpu
def test_moment_map_written_and_read_back(tmp_path):
    import warnings
    import spectral_cube_b200 as scb
    path, mpath = str(tmp_path / 'c.fits'), str(tmp_path / 'm0.fits')
    d = _cube((12, 16, 24), seed=4) + np.float32(3)
    io_fits.write_fits(path, d, HDR)
    cube = scb.SpectralCube.read(path)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m0 = cube.moment0()
    m0.write(mpath)
    with open(mpath, 'rb') as f:
        hdr, off = io_fits.read_header(f)
        back = np.frombuffer(f.read(m0.size * 8), dtype='>f8').reshape(m0.shape)
    assert hdr['BITPIX'] == -64 and hdr['NAXIS'] == 2 and hdr['CTYPE1'] == 'RA---TAN' and hdr['BUNIT'] == m0.unit
    assert np.array_equal(np.isnan(back), np.isnan(m0.value))
    ok = ~np.isnan(back)
    assert np.array_equal(back[ok], m0.value[ok])


@pytest.mark.gpu
def test_degenerate_stokes_axis_is_dropped(tmp_path):
    """A 4-axis file with NAXIS4 = 1 (the layout of the reference's example_cube.fits): one Stokes plane."""
    import spectral_cube_b200 as scb
    path = str(tmp_path / 's.fits')
    d = _cube((1, 4, 7, 9), seed=2)
    io_fits.write_fits(path, d, dict(HDR, CTYPE4='STOKES', CRVAL4=1.0, CRPIX4=1.0, CDELT4=1.0))
    cube = scb.SpectralCube.read(path)
    assert cube.shape == (4, 7, 9)
    assert np.array_equal(cube._data.cpu().numpy().view(np.uint32), d[0].view(np.uint32))
    d2 = np.zeros((2, 3, 4), dtype=np.float32)
    io_fits.write_fits(str(tmp_path / 'img.fits'), d2[0], HDR)
    with pytest.raises(io_fits.FITSReadError):
        scb.SpectralCube.read(str(tmp_path / 'img.fits'))               # "Data should be 3- or 4-dimensional"
