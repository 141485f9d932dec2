"""
Sharded-versus-whole parity on the GPUs of one job (no oracle involved: both sides are the product).

``sharded_parity()`` runs inside an initialised ``torch.distributed`` NCCL job, one rank per GPU.  Every
rank regenerates the same small synthetic cube, processes its own row block through ``RowShardedCube``
and compares with the single-GPU result on the whole cube: moments (no exchange), ``spatial_smooth``
(halo rows from the neighbours, both exchange modes), ``convolve_to``, ``reproject`` (rows -> channels
re-shard) and the interpolate -> reproject job whose interpolation kernel scatters to the channel owners.  Results must be bit-identical.  Used by ``tests/test_multigpu_gpu.py`` and by
``bench.py --gpus N`` (N >= 2), so that the sharded path has evidence wherever a multi-GPU job runs.
"""
import warnings

import numpy as np


def sharded_parity(group=None, rows_per_rank=48):
    import torch
    import torch.distributed as dist
    from . import SpectralCube, DaskSpectralCube, LazyMask, Gaussian2DKernel, Beam
    from . import distributed as D
    from .synth import synth_cube, benchmark_wcs

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nchan, ny, nx = 24, rows_per_rank * world, 256
    wcs = benchmark_wcs(nchan, ny, nx)
    full = synth_cube(nchan, ny, nx, nan_permille=5, border=3)          # every rank can regenerate all of it
    y0, y1 = D.row_partition(ny, world)[rank]
    local = synth_cube(nchan, y1 - y0, nx, y0=y0, ny_total=ny, nx_total=nx, nan_permille=5, border=3)
    res = {'synthetic_rows': bool(torch.equal(torch.nan_to_num(local, nan=-1.0), torch.nan_to_num(full[:, y0:y1], nan=-1.0)))}

    def with_isfinite(c):
        c._mask = LazyMask(np.isfinite, cube=c)
        return c

    detail = {}

    def same(a, b, name=None):
        ok = bool(torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)))
        if not ok and name is not None:
            a64, b64 = a.double(), b.double()
            nan_differs = int((torch.isnan(a64) != torch.isnan(b64)).sum().item())
            d = torch.nan_to_num(a64 - b64, nan=0.0).abs()
            rows = torch.nonzero(d.amax(dim=(0, 2)) > 0).flatten().tolist()
            detail[name] = 'rank %d: %d NaN mismatches, max |diff| %.3g, local rows %s' % (rank, nan_differs, float(d.max().item()), rows[:12])
        return ok

    whole = with_isfinite(DaskSpectralCube(full, wcs, unit='K', allow_huge_operations=True))
    shard = D.RowShardedCube.from_full_wcs(DaskSpectralCube, local, wcs, ny, unit='K', allow_huge_operations=True)
    with_isfinite(shard.local)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for order in (0, 1, 2):
            ref = whole.moment(order=order).value
            got = shard.moment(order=order).value
            res['moment%d' % order] = bool(np.array_equal(got, ref[y0:y1], equal_nan=True))
            res['gather%d' % order] = bool(np.array_equal(shard.moment(order=order, gather=True), ref, equal_nan=True))
        k = Gaussian2DKernel(8 / 2.3548200450309493)
        ref = whole.spatial_smooth(k)._data
        for mode in ('p2p', 'allgather'):
            got = shard.spatial_smooth(k, halo_mode=mode).local._data
            res['spatial_' + mode] = same(got, ref[:, y0:y1], 'spatial_' + mode)
        pix = float(abs(wcs.cdelt[1]))
        for cls in (SpectralCube, DaskSpectralCube):
            w2 = with_isfinite(cls(full, wcs, unit='Jy/beam', beam=Beam(3 * pix), allow_huge_operations=True))
            s2 = D.RowShardedCube.from_full_wcs(cls, local, wcs, ny, unit='Jy/beam', beam=Beam(3 * pix),
                                                allow_huge_operations=True)
            with_isfinite(s2.local)
            ref = w2.convolve_to(Beam(5 * pix))._data
            got = s2.convolve_to(Beam(5 * pix)).local._data
            res['convolve_to_' + cls.__name__] = same(got, ref[:, y0:y1], 'convolve_to_' + cls.__name__)
        # the rows -> channels re-shard: NCCL all-to-all against the peer-memory scatter kernel (when the platform has it)
        via_nccl = D.reshard_rows_to_channels(local, ny, mode='nccl')
        c0, c1 = D.channel_partition(nchan, world)[rank]
        res['reshard_nccl'] = same(via_nccl, full[c0:c1], 'reshard_nccl')
        try:
            via_peer = D.reshard_rows_to_channels(local, ny, mode='peer')
            res['reshard_peer'] = same(via_peer, full[c0:c1], 'reshard_peer')
        except Exception as exc:                      # reported, not a parity failure: the NCCL path serves
            detail['reshard_peer_unavailable'] = repr(exc)[:300]
        a = np.radians(30.0)
        hdr = dict(whole.header)
        hdr.update({'PC1_1': np.cos(a), 'PC1_2': -np.sin(a), 'PC2_1': np.sin(a), 'PC2_2': np.cos(a)})
        ref = whole.reproject(hdr)._data_hi
        sub, (c0, c1) = shard.reproject(hdr)
        res['reproject'] = same(sub._data_hi, ref[c0:c1], 'reproject')
        # config 5 as one sharded job: the interpolation kernel scatters its output to the channel owners (peer memory when
        # the platform has it), then every rank reprojects its channels -- against the same two calls on the whole cube
        grid = whole.spectral_axis[::2]
        ref_i = whole.spectral_interpolate(grid)
        ci0, ci1 = D.channel_partition(len(grid), world)[rank]
        for mode in ('twostep', 'peer'):
            try:
                sub, (a0, a1) = shard.spectral_interpolate_to_channels(grid, mode=mode)
                res['interp_to_channels_' + mode] = (a0, a1) == (ci0, ci1) and same(sub._data, ref_i._data[ci0:ci1], 'interp_to_channels_' + mode)
            except Exception as exc:
                if mode == 'twostep':
                    raise
                detail['interp_to_channels_peer_unavailable'] = repr(exc)[:300]
        hdr2 = dict(ref_i.header)
        hdr2.update({'PC1_1': np.cos(a), 'PC1_2': -np.sin(a), 'PC2_1': np.sin(a), 'PC2_2': np.cos(a)})
        ref = ref_i.reproject(hdr2)._data_hi
        sub, (a0, a1) = shard.spectral_interpolate_reproject(grid, hdr2)
        res['interp_reproject'] = same(sub._data_hi, ref[a0:a1], 'interp_reproject')
    res['_detail'] = detail
    return res


def all_ranks_agree(res, group=None):
    """AND of every entry over the ranks: (ok, names that failed on some rank)."""
    import torch
    import torch.distributed as dist
    names = sorted(n for n in res if not n.startswith('_'))
    t = torch.tensor([1 if res[n] else 0 for n in names], dtype=torch.int32, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    flags = t.cpu().tolist()
    bad = [n for n, f in zip(names, flags) if not f]
    return (not bad), bad
