"""
Multi-GPU parity (needs >= 2 GPUs; run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`).
Two NCCL ranks each hold a row block; results must equal the single-GPU result on the whole cube:
moments (no exchange), spatial_smooth (halo rows from the neighbour, both exchange modes) and
reproject (rows -> channels all-to-all, then whole planes locally).
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import spectral_cube_b200 as scb
        from spectral_cube_b200 import distributed as D
        from spectral_cube_b200.synth import synth_cube, benchmark_wcs
        nchan, ny, nx = 24, 96, 256
        wcs = benchmark_wcs(nchan, ny, nx)
        full = synth_cube(nchan, ny, nx, nan_permille=5, border=3)          # every rank can regenerate all of it
        y0, y1 = D.row_partition(ny, world)[rank]
        local = synth_cube(nchan, y1 - y0, nx, y0=y0, ny_total=ny, nx_total=nx, nan_permille=5, border=3)
        assert torch.equal(torch.nan_to_num(local, nan=-1.0), torch.nan_to_num(full[:, y0:y1], nan=-1.0))

        def with_isfinite(c):
            c._mask = scb.LazyMask(np.isfinite, cube=c)
            return c
        whole = with_isfinite(scb.DaskSpectralCube(full, wcs, unit='K'))
        shard = D.RowShardedCube.from_full_wcs(scb.DaskSpectralCube, local, wcs, ny, unit='K')
        with_isfinite(shard.local)
        res = {}
        # moments: local rows equal the same rows of the whole-cube map, and the gathered map is whole
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            for order in (0, 1, 2):
                ref = whole.moment(order=order).value
                got = shard.moment(order=order).value
                res['moment%d' % order] = bool(np.array_equal(got, ref[y0:y1], equal_nan=True))
                res['gather%d' % order] = bool(np.array_equal(shard.moment(order=order, gather=True), ref, equal_nan=True))
        # spatial smooth with halos, both exchange modes
        k = scb.Gaussian2DKernel(8 / 2.3548200450309493)
        ref = whole.spatial_smooth(k)._data
        for mode in ('p2p', 'allgather'):
            got = shard.spatial_smooth(k, halo_mode=mode).local._data
            res['spatial_' + mode] = bool(torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref[:, y0:y1], nan=-7.0)))
        # convolve_to on row shards: both classes, Jy/beam rescale and the convolve_fft rule included
        pix = float(abs(wcs.cdelt[1]))
        for cls in (scb.SpectralCube, scb.DaskSpectralCube):
            w2 = with_isfinite(cls(full, wcs, unit='Jy/beam', beam=scb.Beam(3 * pix)))
            s2 = D.RowShardedCube.from_full_wcs(cls, local, wcs, ny, unit='Jy/beam', beam=scb.Beam(3 * pix))
            with_isfinite(s2.local)
            ref = w2.convolve_to(scb.Beam(5 * pix))._data
            got = s2.convolve_to(scb.Beam(5 * pix)).local._data
            res['convolve_to_' + cls.__name__] = bool(torch.equal(torch.nan_to_num(got, nan=-7.0),
                                                                  torch.nan_to_num(ref[:, y0:y1], nan=-7.0)))
        # reproject through the row->channel re-shard
        a = np.radians(30.0)
        hdr = dict(whole.header)
        hdr.update({'PC1_1': np.cos(a), 'PC1_2': -np.sin(a), 'PC2_1': np.sin(a), 'PC2_2': np.cos(a)})
        ref = whole.reproject(hdr)._data_hi
        sub, (c0, c1) = shard.reproject(hdr)
        res['reproject'] = bool(torch.equal(torch.nan_to_num(sub._data_hi, nan=-7.0), torch.nan_to_num(ref[c0:c1], nan=-7.0)))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
def test_two_rank_row_sharding_matches_single_gpu():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get() for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for r in range(world):
        bad = [k for k, v in out[r].items() if not v]
        assert not bad, "rank %d: %s" % (r, bad)
