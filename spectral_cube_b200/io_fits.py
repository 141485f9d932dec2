"""
FITS cubes in and moment maps out, without astropy.

Mirrors ``load_fits_cube`` (spectral_cube/io/fits.py:171-260): primary-HDU image, 3 axes (or 4 with
degenerate trailing axes, as `cube_utils._split_stokes` leaves for a single Stokes plane), the WCS
from the header, ``BUNIT`` into ``meta``, and ``LazyMask(np.isfinite)`` attached (:214).  The data
block is never decoded on the CPU: the raw big-endian bytes are streamed through two pinned staging
buffers to the device, where ``sc_fits_decode`` byte-swaps (and applies BSCALE / BZERO / BLANK) at
HBM speed, copy and decode overlapping on two CUDA streams.

``write_fits`` is the egress side for `Projection` / cube data (io/fits.py:262-282, 2880-byte blocks,
big-endian float32 / float64).
"""
import os

import numpy as np

from . import _lib

BLOCK = 2880
CARD = 80


READ_THREADS = 8
_POOL = None


class FITSReadError(Exception):
    pass


def _read_pool():
    global _POOL
    if _POOL is None:
        import concurrent.futures
        _POOL = concurrent.futures.ThreadPoolExecutor(max_workers=READ_THREADS)
    return _POOL


def _pread_into(fd, view, offset):
    done = 0
    while done < len(view):
        n = os.preadv(fd, [view[done:]], offset + done)
        if n <= 0:
            break
        done += n
    return done


def _parse_value(raw):
    raw = raw.strip()
    if raw.startswith("'"):
        end = 1
        out = []
        while end < len(raw):
            if raw[end] == "'":
                if end + 1 < len(raw) and raw[end + 1] == "'":
                    out.append("'")
                    end += 2
                    continue
                break
            out.append(raw[end])
            end += 1
        return ''.join(out).rstrip()
    raw = raw.split('/')[0].strip()
    if raw in ('T', 'F'):
        return raw == 'T'
    try:
        return int(raw)
    except ValueError:
        pass
    try:
        return float(raw.replace('D', 'E').replace('d', 'e'))
    except ValueError:
        return raw


def read_header(f):
    """Parse the primary header of an open binary file; returns (ordered dict of cards, data offset)."""
    hdr = {}
    nread = 0
    while True:
        block = f.read(BLOCK)
        if len(block) < BLOCK:
            raise FITSReadError("truncated FITS header")
        nread += BLOCK
        for i in range(0, BLOCK, CARD):
            card = block[i:i + CARD].decode('ascii', errors='replace')
            key = card[:8].strip()
            if key == 'END':
                return hdr, nread
            if not key or key in ('COMMENT', 'HISTORY') or card[8:10] != '= ':
                continue
            hdr[key] = _parse_value(card[10:])


def _format_card(key, value):
    if isinstance(value, bool):
        v = '%20s' % ('T' if value else 'F')
    elif isinstance(value, (int, np.integer)):
        v = '%20d' % int(value)
    elif isinstance(value, (float, np.floating)):
        v = '%20s' % ('%.16G' % float(value))
    else:
        v = "'%-8s'" % str(value).replace("'", "''")
    return ('%-8s= %s' % (key, v)).ljust(CARD)[:CARD]


def write_fits(path, data, header=None, overwrite=False):
    """Write a primary-HDU image: numpy array (any float/int dtype is written as float32 unless it is
    float64) and a flat header mapping (WCS cards, BUNIT ...)."""
    if os.path.exists(path) and not overwrite:
        raise OSError("File %r already exists; use overwrite=True" % path)
    arr = np.asarray(data)
    out_dtype, bitpix = ('>f8', -64) if arr.dtype == np.float64 else ('>f4', -32)
    cards = [_format_card('SIMPLE', True), _format_card('BITPIX', bitpix), _format_card('NAXIS', arr.ndim)]
    for i, n in enumerate(arr.shape[::-1]):
        cards.append(_format_card('NAXIS%d' % (i + 1), n))
    for k, v in (header or {}).items():
        if k in ('SIMPLE', 'BITPIX', 'NAXIS', 'END', 'EXTEND', 'BSCALE', 'BZERO', 'BLANK') or k.startswith('NAXIS') or v is None:
            continue
        cards.append(_format_card(k, v))
    cards.append('END'.ljust(CARD))
    text = ''.join(cards)
    text += ' ' * (-len(text) % BLOCK)
    with open(path, 'wb') as f:
        f.write(text.encode('ascii'))
        raw = np.ascontiguousarray(arr, dtype=out_dtype).tobytes()
        f.write(raw)
        f.write(b'\0' * (-len(raw) % BLOCK))


def read_fits_to_device(path, device=None, chunk_bytes=256 << 20):
    """(float32 device tensor shaped like the FITS image, header dict).  The file is read in blocks of
    `chunk_bytes` into two pinned buffers; block k+1 is read from disk while block k is copied and
    decoded on the device."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else device)
    with open(path, 'rb') as f:
        hdr, offset = read_header(f)
        if not hdr.get('SIMPLE', False):
            raise FITSReadError("not a standard FITS file")
        naxis = int(hdr.get('NAXIS', 0))
        if naxis == 0:
            raise FITSReadError('No data found in HDU 0. You can try using the hdu= '
                                'keyword argument to read data from another HDU.')
        shape = tuple(int(hdr['NAXIS%d' % (i + 1)]) for i in range(naxis))[::-1]
        bitpix = int(hdr['BITPIX'])
        n = int(np.prod(shape))
        bps = abs(bitpix) // 8
        bscale, bzero = float(hdr.get('BSCALE', 1.0)), float(hdr.get('BZERO', 0.0))
        has_blank = 'BLANK' in hdr and bitpix > 0
        blank = int(hdr.get('BLANK', 0)) if has_blank else 0
        out = torch.empty(shape, dtype=torch.float32, device=dev)
        flat = out.view(-1)
        per = max(16, (chunk_bytes // 16) * 16)
        per -= per % (16 * bps)                                   # whole samples, 16-byte multiples
        staging = [torch.empty(per, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        dstage = [torch.empty(per, dtype=torch.uint8, device=dev) for _ in range(2)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        done = [None, None]
        fd = f.fileno()
        pool = _read_pool()
        total = n * bps
        pos = 0
        k = 0
        while pos < total:
            nb = min(per, total - pos)
            b = k & 1
            if done[b] is not None:
                done[b].synchronize()                             # the staging buffer is free again
            view = memoryview(staging[b].numpy()[:nb])
            # the page-cache -> pinned copy is the slowest stage (~7 GB/s per thread): slices in parallel
            # (os.preadv releases the GIL)
            step = -(-nb // READ_THREADS)
            step += -step % 4096
            jobs = [pool.submit(_pread_into, fd, view[o:min(nb, o + step)], offset + pos + o) for o in range(0, nb, step)]
            got = sum(j.result() for j in jobs)
            if got != nb:
                raise FITSReadError("truncated FITS data block (%d of %d bytes)" % (pos + got, total))
            with torch.cuda.stream(streams[b]):
                dstage[b][:nb].copy_(staging[b][:nb], non_blocking=True)
                _lib.check(lib.sc_fits_decode(dstage[b].data_ptr(), flat[pos // bps:].data_ptr(), nb // bps, bitpix,
                                              bscale, bzero, 1 if has_blank else 0, blank, streams[b].cuda_stream))
                ev = torch.cuda.Event()
                ev.record(streams[b])
                done[b] = ev
            pos += nb
            k += 1
        for s in streams:
            s.synchronize()
    return out, hdr


def load_fits_cube(path, cube_cls, use_dask=False, device=None, **kwargs):
    """io/fits.py:171-260 for a single-Stokes cube."""
    from .masks import LazyMask
    data, hdr = read_fits_to_device(path, device=device)
    while data.dim() > 3 and data.shape[0] == 1:                  # degenerate Stokes / extra axes
        data = data[0]
    if data.dim() != 3:
        raise FITSReadError("Data should be 3- or 4-dimensional")
    meta = dict(kwargs.pop('meta', None) or {})
    if 'BUNIT' in hdr:
        meta['BUNIT'] = hdr['BUNIT']
    cube = cube_cls(data, hdr, meta=meta, header=hdr, **kwargs)
    cube._mask = LazyMask(np.isfinite, cube=cube)                 # io/fits.py:214
    return cube
