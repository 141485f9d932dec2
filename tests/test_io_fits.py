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


GOLDEN_FITS = __import__('os').path.join(__import__('os').path.dirname(__file__), 'golden', 'example_cube.fits')


def test_reference_fixture_header_parses():
    """tests/golden/example_cube.fits is the reference's own fixture (spectral_cube/tests/data/example_cube.fits,
    read by tests/test_io.py:16-18 and tests/test_dask.py:240-242): BITPIX -32, NAXIS 3 x 4 x 7 x 1."""
    with open(GOLDEN_FITS, 'rb') as f:
        hdr, off = io_fits.read_header(f)
    assert off == 5760 and hdr['BITPIX'] == -32 and hdr['NAXIS'] == 4
    assert (hdr['NAXIS1'], hdr['NAXIS2'], hdr['NAXIS3'], hdr['NAXIS4']) == (3, 4, 7, 1)
    assert hdr['BUNIT'] == 'Jy/beam' and hdr['CTYPE1'] == 'RA---ARC' and hdr['CTYPE3'] == 'VRAD'


@pytest.mark.gpu
@pytest.mark.parametrize('use_dask', [False, True])
def test_reference_fixture_reads_like_numpy(use_dask):
    """`SpectralCube.read` (io/fits.py:171-260) of the reference-held file against a plain numpy '>f4' view of its
    data block: same voxels bit for bit, the degenerate fourth axis dropped, isfinite mask (:214), unit and beam from
    the header (cube_utils.try_load_beam), both classes alike (tests/test_dask.py:240-252)."""
    import spectral_cube_b200 as scb
    want = np.fromfile(GOLDEN_FITS, dtype='>f4', offset=5760, count=3 * 4 * 7).reshape(7, 4, 3).astype(np.float32)
    cube = scb.SpectralCube.read(GOLDEN_FITS, use_dask=use_dask)
    assert type(cube) is (scb.DaskSpectralCube if use_dask else scb.SpectralCube)
    assert cube.shape == (7, 4, 3) and cube.unit == 'Jy/beam'
    assert np.array_equal(cube._data.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert np.array_equal(cube.mask.include(), np.isfinite(want))
    np.testing.assert_array_equal(cube.filled_data[:], np.where(np.isfinite(want), want, np.nan))
    assert abs(cube.beam.major - 0.0003467814737101) < 1e-15 and abs(cube.beam.pa - 22.12153897146) < 1e-10
    # VRAD with a blank CUNIT3 is in m/s: channel i sits at CRVAL3 + CDELT3 (i + 1 - CRPIX3)
    np.testing.assert_allclose(cube.spectral_axis, 7000.0 - 103.6813929677 * (np.arange(7) + 1 - 77.62811279297), rtol=1e-12)
    assert np.nanmin(want) == np.float32(-0.0140879265964) and np.nanmax(want) == np.float32(0.01936739496887)   # DATAMIN / DATAMAX


@pytest.mark.gpu
@pytest.mark.parametrize('use_dask', [False, True])
def test_masked_cube_is_written_filled(tmp_path, use_dask):
    """`cube.write` saves `PrimaryHDU(self.unitless_filled_data[:])` (spectral_cube.py:2563-2570, dask :1400-1405):
    masked voxels reach the file as the fill value, not as the raw data."""
    import spectral_cube_b200 as scb
    path, out = str(tmp_path / 'c.fits'), str(tmp_path / 'masked.fits')
    d = _cube((6, 9, 12), seed=11)
    io_fits.write_fits(path, d, HDR)
    cube = scb.SpectralCube.read(path, use_dask=use_dask)
    masked = cube.with_mask(cube > 0.25)
    masked.write(out)
    def data_of(fn):
        with open(fn, 'rb') as f:
            hdr, off = io_fits.read_header(f)
        assert hdr['BITPIX'] == -32
        return np.fromfile(fn, dtype='>f4', offset=off, count=d.size).reshape(d.shape).astype(np.float32)
    back = data_of(out)
    want = np.where(np.isfinite(d) & (d > 0.25), d, np.nan).astype(np.float32)
    assert np.array_equal(np.isnan(back), np.isnan(want)) and np.isnan(back).sum() > d.size // 3
    assert np.array_equal(back[~np.isnan(back)], want[~np.isnan(want)])
    masked.with_fill_value(-7.0).write(out, overwrite=True)
    back = data_of(out)
    assert np.array_equal(back, np.where(np.isnan(want), np.float32(-7.0), want))
    # a reprojected cube holds float64 (reproject_interp's output): the file does too
    rp = cube.reproject(dict(cube.header))
    rp.write(out, overwrite=True)
    with open(out, 'rb') as f:
        hdr, off = io_fits.read_header(f)
    assert hdr['BITPIX'] == -64
