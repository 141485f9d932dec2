"""
Host-side logic of the multi-GPU path on CPU: world_size-2 and -3 `gloo` process groups exercise
the row partition, the halo exchange (neighbour send/recv and the all-gather variant), the
rows<->channels re-shard and the gather of row-sharded maps; the last test also drives the row-sharded
spatial_smooth / convolve_to wrappers into the real library (calls stop at the missing device).  No GPU.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spectral_cube_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        q.put((rank, fn(rank, world)))
    finally:
        dist.destroy_process_group()


def run_group(world, fn):
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get() for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return [out[r] for r in range(world)]


def test_row_partition_is_balanced_and_covers_everything():
    for ny, w in [(4096, 8), (2048, 3), (7, 8), (10, 4)]:
        parts = D.row_partition(ny, w)
        assert parts[0][0] == 0 and parts[-1][1] == ny
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


def _full_cube(nchan=5, ny=11, nx=6):
    return torch.arange(nchan * ny * nx, dtype=torch.float32).reshape(nchan, ny, nx)


def _halo_job(mode):
    def job(rank, world):
        full = _full_cube()
        y0, y1 = D.row_partition(full.shape[1], world)[rank]
        local = full[:, y0:y1]
        h = 2
        top, bot = D.exchange_halo_rows(local[:, :h].contiguous(), local[:, -h:].contiguous(), mode=mode)
        ok = True
        if rank == 0:
            ok &= top is None
        else:
            ok &= torch.equal(top, full[:, y0 - h:y0])
        if rank == world - 1:
            ok &= bot is None
        else:
            ok &= torch.equal(bot, full[:, y1:y1 + h])
        return bool(ok)
    return job


def _halo_p2p(rank, world):
    return _halo_job('p2p')(rank, world)


def _halo_allgather(rank, world):
    return _halo_job('allgather')(rank, world)


@pytest.mark.parametrize('world', [2, 3])
def test_halo_exchange_p2p(world):
    assert all(run_group(world, _halo_p2p))


def test_halo_exchange_allgather():
    assert all(run_group(2, _halo_allgather))


def _reshard_job(rank, world):
    full = _full_cube()
    nchan, ny, nx = full.shape
    y0, y1 = D.row_partition(ny, world)[rank]
    c0, c1 = D.channel_partition(nchan, world)[rank]
    chan = D.reshard_rows_to_channels(full[:, y0:y1].contiguous(), ny)
    ok = torch.equal(chan, full[c0:c1])
    back = D.reshard_channels_to_rows(chan, nchan)
    ok &= torch.equal(back, full[:, y0:y1])
    return bool(ok)


@pytest.mark.parametrize('world', [2, 3])
def test_reshard_rows_channels_roundtrip(world):
    assert all(run_group(world, _reshard_job))


def _gather_job(rank, world):
    full = torch.arange(11 * 6, dtype=torch.float64).reshape(11, 6)
    y0, y1 = D.row_partition(11, world)[rank]
    got = D.gather_rows(full[y0:y1].contiguous(), 11)
    only0 = D.gather_rows(full[y0:y1].contiguous(), 11, dst=0)
    return bool(torch.equal(got, full)) and ((only0 is None) == (rank != 0))


@pytest.mark.parametrize('world', [2, 3])
def test_gather_rows(world):
    assert all(run_group(world, _gather_job))


def _sharded_smoothing_job(rank, world):
    """Row-sharded spatial_smooth / convolve_to host logic under gloo: tensors stay on the host, the real library is
    called and stops at its first CUDA call (tolerated here only, like the `host` fixture of conftest.py); the halo
    exchange, the job-wide blank-plane all-reduce and the strategy-count all-reduce run for real."""
    import spectral_cube_b200 as S
    from spectral_cube_b200 import cube as C, _lib
    from spectral_cube_b200.masks import LazyMask

    class _Stream(object):
        cuda_stream = 0

    calls = []

    def check(rc):
        msg = _lib.load().sc_last_error().decode() if rc else ''
        calls.append(rc)
        if rc and 'CUDA error' not in msg and 'not available from the driver' not in msg:
            raise AssertionError("the library refused the arguments: %d %s" % (rc, msg))

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    _lib.require_cuda = lambda: torch
    _lib.check = check
    C._stream = lambda: 0
    wcs = S.CubeWCS(ctype=['RA---SIN', 'DEC--SIN', 'VOPT'], crval=[23.0, 30.0, -321.0], crpix=[20.0, 30.0, 1.0],
                    cdelt=[-5.55555561268e-4, 5.55555561268e-4, 1.288], cunit=['deg', 'deg', 'km/s'])
    ny = 60
    y0, y1 = D.row_partition(ny, world)[rank]
    rng = np.random.default_rng(rank)
    ok = True
    for cls in (S.SpectralCube, S.DaskSpectralCube):
        for fill in (np.nan, 5.0):
            sh = D.RowShardedCube.from_full_wcs(cls, rng.normal(size=(3, y1 - y0, 64)).astype(np.float32), wcs, ny,
                                                unit='Jy/beam', beam=S.Beam.from_arcsec(3.0), fill_value=fill)
            sh.local._mask = LazyMask(np.isfinite, cube=sh.local)
            sh = sh.with_mask(sh > -10.0)
            out = sh.convolve_to(S.Beam.from_arcsec(5.0))
            ok &= out.local.beam == S.Beam.from_arcsec(5.0) and out.local.shape == sh.local.shape
            ok &= type(out.local) is cls and out.y0 == y0
            sm = sh.spatial_smooth(S.Gaussian2DKernel(1.0), raise_error_jybm=False, halo_mode='allgather')
            ok &= sm.local.beam == sh.local.beam
            # job-wide blank planes: plane 1 includes nothing on any rank, plane 2 nothing on rank 0 only -> flags 0, 1, 0
            # exactly (a uint8 `any()` followed by `~` once made every plane "blank": the Jy/beam rescale was skipped)
            inc = torch.ones((3, y1 - y0, 64), dtype=torch.uint8)
            inc[1] = 0
            if rank == 0:
                inc[2] = 0
            sh.local._mask._include_tensor = lambda data=None, _inc=inc: _inc
            ok &= sh._blank_planes().tolist() == [0, 1, 0]
    return bool(ok) and len(calls) > 8


@pytest.mark.parametrize('world', [2, 3])
def test_row_sharded_smoothing_host_logic(world):
    assert all(run_group(world, _sharded_smoothing_job))


def test_channel_row_pointers_address_the_owners_rows():
    """The table the scattering interpolation kernel stores through: channel j of a rank's rows [y0, y0 + rows) lands
    in its owner's buffer at (j - c0, y0, 0); checked by writing through the table into numpy buffers."""
    from spectral_cube_b200 import distributed as D
    nout, ny_total, nx, world = 11, 12, 8, 3
    cparts = D.channel_partition(nout, world)
    chans_max = max(b - a for a, b in cparts)
    bufs = [np.full((chans_max, ny_total, nx), -1.0, dtype=np.float32) for _ in range(world)]
    base = [b.ctypes.data for b in bufs]
    for rank, (y0, y1) in enumerate(D.row_partition(ny_total, world)):
        ptrs = D.channel_row_pointers(base, cparts, ny_total, nx, y0)
        assert ptrs.dtype == np.int64 and ptrs.shape == (nout,)
        for j in range(nout):
            owner = next(d for d, (a, b) in enumerate(cparts) if a <= j < b)
            assert (int(ptrs[j]) - base[owner]) % 16 == 0                  # 16-byte vector stores need nx % 4 == 0 rows
            off = (int(ptrs[j]) - base[owner]) // 4
            assert 0 <= off and off + (y1 - y0) * nx <= bufs[owner].size
            rows = bufs[owner].reshape(-1)[off:off + (y1 - y0) * nx].reshape(y1 - y0, nx)
            rows[:] = 100 * j + np.arange(y0, y1)[:, None]                 # what a kernel storing through the table would do
    for d, (a, b) in enumerate(cparts):
        want = 100 * np.arange(a, b)[:, None, None] + np.arange(ny_total)[None, :, None] + np.zeros((1, 1, nx))
        np.testing.assert_array_equal(bufs[d][:b - a], want.astype(np.float32))
        assert np.all(bufs[d][b - a:] == -1.0)
