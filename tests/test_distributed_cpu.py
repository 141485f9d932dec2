"""
Host-side logic of the multi-GPU path on CPU: world_size-2 and -3 `gloo` process groups exercise
the row partition, the halo exchange (neighbour send/recv and the all-gather variant), the
rows<->channels re-shard and the gather of row-sharded maps.  No GPU, no CUDA library calls.
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
