"""
Dry-run aid, TEST TOOLING ONLY (nothing in the product or in the default test runs imports it).

`tests/test_convolve_to_gpu.py` was written after the round's GPU budget was spent.  To check the TESTS --
their expectations, tolerances, NaN patterns, the host plumbing up to the C ABI -- before their first run on
hardware, this pytest plugin swaps `spectral_cube_b200._lib.load()` for a numpy emulation of the few entry points
`convolve_to` / `statistics` reach, written from their documented semantics in include/sc_b200.h
(sc_spatial_smooth_sep_ex, sc_spatial_smooth_2d, sc_scale, sc_fill_masked, sc_mask_include, sc_reduce_axis0 and the
mask descriptor), keeps tensors on the host and un-skips the gpu-marked tests of the file it is pointed at:

    PYTHONPATH=tools/dryrun python -m pytest -p emu_plugin tests/test_convolve_to_gpu.py -q -p no:cacheprovider

A green run says the tests are self-consistent with the documented kernel semantics; it says NOTHING about the CUDA
kernels.  It found three wrong expectations and one real host-side issue (the separability test for beam kernels).
"""
import ctypes as C
import sys
import numpy as np
import pytest
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import oracle.convolve as oconv


def view(ptr, dtype, shape, strides):
    """numpy view of `shape` elements at address `ptr` with element `strides`."""
    if hasattr(ptr, 'contents') or isinstance(ptr, C._Pointer):
        ptr = C.cast(ptr, C.c_void_p).value
    dtype = np.dtype(dtype)
    n = 1 + sum((s - 1) * st for s, st in zip(shape, strides))
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    flat = np.frombuffer(buf, dtype=dtype)
    return np.lib.stride_tricks.as_strided(flat, shape=shape, strides=tuple(st * dtype.itemsize for st in strides))


def include(desc, cube, shape):
    n = desc.n_nodes
    if n == 0:
        return np.ones(shape, dtype=bool)
    res = []
    for i in range(n):
        nd = desc.nodes[i]
        src = cube if not nd.data else view(nd.data, np.float32, shape, (nd.ds_c, nd.ds_y, 1))
        if nd.kind == 1:
            r = np.isfinite(src)
        elif nd.kind in (2, 3):
            if nd.kind == 2:
                t = nd.value
            else:
                t = view(nd.array, np.float64 if nd.array_dtype == 1 else np.float32, shape, (nd.as_c, nd.as_y, nd.as_x)).astype(np.float64)
            x = src.astype(np.float64)
            with np.errstate(invalid='ignore'):
                r = [x > t, x >= t, x < t, x <= t, x == t, x != t][nd.op]
        elif nd.kind == 4:
            r = view(nd.array, np.uint8, shape, (nd.as_c, nd.as_y, nd.as_x)) != 0
        elif nd.kind == 5:
            r = res[nd.a] & res[nd.b]
        elif nd.kind == 6:
            r = res[nd.a] | res[nd.b]
        elif nd.kind == 7:
            r = res[nd.a] ^ res[nd.b]
        elif nd.kind == 8:
            r = ~res[nd.a]
        res.append(np.broadcast_to(r, shape))
    return res[-1]


class FakeLib(object):
    launches = 0

    def sc_last_error(self):
        return b''

    def sc_workspace_bytes(self, op, nchan, ny, nx, aux):
        return aux * 32 + nchan * 24 + 4096

    def _smooth(self, in_, out, out_dtype, nchan, ny, nx, sc, sy, osc, osy, mask, fill, k2d, halo_rows, passthrough, ws):
        assert halo_rows == 0
        shape = (nchan, ny, nx)
        data = view(in_, np.float32, shape, (sc, sy, 1))
        dst = view(out, np.float64 if out_dtype == 1 else np.float32, shape, (osc, osy, 1))
        inc = include(mask, data, shape)
        filled = np.where(inc, data, np.float32(fill)).astype(np.float64)
        flags = None
        if passthrough and mask.n_nodes > 0:
            flags = view(ws + k2d.size * 8 + 512, np.uint8, (nchan,), (1,))
        for c in range(nchan):
            blank = flags is not None and not inc[c].any()
            if flags is not None:
                flags[c] = 1 if blank else 0
            dst[c] = filled[c] if blank else oconv.convolve(filled[c], k2d, normalize_kernel=True)
        FakeLib.launches += 1
        return 0

    def sc_spatial_smooth_sep_ex(self, in_, out, out_dtype, nchan, ny, nx, sc, sy, osc, osy, mask, fill, ty, nty, tx, ntx,
                                 ht, hb, halo_rows, passthrough, counts, ws, wsb, stream):
        ky = np.ctypeslib.as_array(ty, shape=(nty,)).copy()
        kx = np.ctypeslib.as_array(tx, shape=(ntx,)).copy()
        return self._smooth(in_, out, out_dtype, nchan, ny, nx, sc, sy, osc, osy, mask, fill, np.outer(ky, kx), halo_rows, passthrough, ws)

    def sc_spatial_smooth_2d(self, in_, out, out_dtype, nchan, ny, nx, sc, sy, osc, osy, mask, fill, taps, nty, ntx,
                             ht, hb, halo_rows, passthrough, ws, wsb, stream):
        k = np.ctypeslib.as_array(taps, shape=(nty * ntx,)).copy().reshape(nty, ntx)
        return self._smooth(in_, out, out_dtype, nchan, ny, nx, sc, sy, osc, osy, mask, fill, k, halo_rows, passthrough, ws)

    def sc_scale(self, data, dtype, nchan, ny, nx, sc, sy, factor, nan_to_zero, skip, stream):
        if data is None:
            return -1
        T = np.float64 if dtype == 1 else np.float32
        a = view(data, T, (nchan, ny, nx), (sc, sy, 1))
        sk = view(skip, np.uint8, (nchan,), (1,)) if skip else np.zeros(nchan, np.uint8)
        for c in range(nchan):
            if sk[c]:
                continue
            v = a[c].astype(np.float64) * factor
            if nan_to_zero:
                v[np.isnan(a[c])] = 0.0
            a[c] = v.astype(T)
        return 0

    def sc_fill_masked(self, cube, nchan, ny, nx, sc, sy, mask, fill, out, stream):
        shape = (nchan, ny, nx)
        data = view(cube, np.float32, shape, (sc, sy, 1))
        view(out, np.float32, shape, (ny * nx, nx, 1))[...] = np.where(include(mask, data, shape), data, np.float32(fill))
        return 0

    def sc_mask_include(self, cube, nchan, ny, nx, sc, sy, mask, out, stream):
        shape = (nchan, ny, nx)
        data = view(cube, np.float32, shape, (sc, sy, 1))
        view(out, np.uint8, shape, (ny * nx, nx, 1))[...] = include(mask, data, shape)
        return 0

    def sc_moments_axis0(self, cube, nchan, ny, nx, sc, sy, mask, offsets, pix_size, world0, want_bits, o0, o1, o2, ws, wsb, stream):
        assert want_bits == 1, "only moment 0 is emulated"
        shape = (nchan, ny, nx)
        data = view(cube, np.float32, shape, (sc, sy, 1))
        f = np.where(include(mask, data, shape) & ~np.isnan(data), data, 0.0).astype(np.float64)
        view(o0, np.float64, (ny, nx), (nx, 1))[...] = f.sum(axis=0) * pix_size
        return 0

    def sc_reduce_spatial(self, cube, nchan, ny, nx, sc, sy, axis, mask, osum, ocnt, om2, omin, omax, oamin, oamax, stream):
        return self._reduce(cube, nchan, ny, nx, sc, sy, axis, mask, osum, ocnt, om2, omin, omax, oamin, oamax)

    def sc_reduce_axis0(self, cube, nchan, ny, nx, sc, sy, mask, osum, ocnt, om2, omin, omax, oamin, oamax, stream):
        return self._reduce(cube, nchan, ny, nx, sc, sy, 0, mask, osum, ocnt, om2, omin, omax, oamin, oamax)

    def _reduce(self, cube, nchan, ny, nx, sc, sy, axis, mask, osum, ocnt, om2, omin, omax, oamin, oamax):
        import warnings
        shape = (nchan, ny, nx)
        data = view(cube, np.float32, shape, (sc, sy, 1))
        f = np.where(include(mask, data, shape), data, np.nan)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            d = f.astype(np.float64)
            cnt = (~np.isnan(d)).sum(axis=axis)
            s = np.where(cnt > 0, np.nansum(d, axis=axis), np.nan)
            m2 = np.where(cnt > 0, np.nansum((d - np.expand_dims(s / np.maximum(cnt, 1), axis)) ** 2, axis=axis), np.nan)
            lo, hi = np.nanmin(f, axis=axis), np.nanmax(f, axis=axis)
            ahi = np.argmax(np.where(np.isnan(f), -np.inf, f), axis=axis); ahi[cnt == 0] = 0
            alo = np.argmin(np.where(np.isnan(f), np.inf, f), axis=axis); alo[cnt == 0] = 0
        for ptr, val, T in ((osum, s, np.float64), (ocnt, cnt, np.int32), (om2, m2, np.float64), (omin, lo, np.float32), (omax, hi, np.float32),
                            (oamin, alo, np.int32), (oamax, ahi, np.int32)):
            if ptr:
                oshape = tuple(n for i, n in enumerate(shape) if i != axis)
                view(ptr, T, oshape, (oshape[1], 1))[...] = val
        return 0


@pytest.hookimpl(trylast=True)
def pytest_collection_modifyitems(config, items):
    for item in items:
        item.own_markers = [m for m in item.own_markers if m.name != 'skip']


@pytest.fixture(autouse=True)
def emulated_device(monkeypatch):
    import torch
    from spectral_cube_b200 import cube as Cb, _lib

    class _Stream(object):
        cuda_stream = 0
    fake = FakeLib()
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda *a, **k: _Stream())
    monkeypatch.setattr(_lib, 'require_cuda', lambda: torch)
    monkeypatch.setattr(_lib, 'load', lambda: fake)
    monkeypatch.setattr(Cb, '_stream', lambda: 0)
    yield
