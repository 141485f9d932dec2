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
        from spectral_cube_b200.selftest import sharded_parity
        res = sharded_parity()
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
        bad = [k for k, v in out[r].items() if not k.startswith('_') and not v]
        assert not bad, "rank %d: %s %s" % (r, bad, out[r].get('_detail'))
