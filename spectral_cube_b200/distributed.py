"""
Spatial-plane sharding across GPUs: one process per GPU, every rank owns a
contiguous block of image rows ``[:, y0:y1, :]`` with whole spectra and whole rows, so moments,
spectral smoothing and spectral interpolation need no communication at all.  The only exchanges
on the path are

  * the ``ceil(k/2)`` halo rows ``spatial_smooth`` needs from the two neighbouring ranks
    (``exchange_halo_rows``: neighbour send/recv in one batch, or an all-gather of the edge strips
    as the north star words it), and
  * one all-to-all re-shard from rows to channels before ``reproject`` (a rotated output row
    block maps to a slanted strip of the input, so each rank would need remote rows; after the
    re-shard every rank reprojects whole planes locally).

``torch.distributed`` (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests) is the
plumbing; tensors are only memory holders.  The reference has no counterpart: its parallelism is
joblib/dask on one host (spectral_cube.py:2951-3019, dask_spectral_cube.py:501-638).
"""
import numpy as np


def row_partition(ny, world_size):
    """Balanced contiguous row blocks: list of (y0, y1) per rank; the first ``ny % world`` ranks get
    one extra row."""
    base, extra = divmod(int(ny), int(world_size))
    out, y = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        out.append((y, y + n))
        y += n
    return out


def channel_partition(nchan, world_size):
    return row_partition(nchan, world_size)


def _dist():
    import torch.distributed as dist
    return dist


def exchange_halo_rows(top_edge, bot_edge, group=None, mode='p2p'):
    """Swap edge strips with the neighbouring ranks.

    ``top_edge`` / ``bot_edge`` are this rank's first / last ``h`` FILLED rows, shape
    (nchan, h, nx).  Returns ``(halo_top, halo_bot)``: the rows just above this shard (the previous
    rank's bottom strip) and just below it (the next rank's top strip); ``None`` at the image
    boundary.  ``mode='p2p'`` moves 2 strips per rank; ``mode='allgather'`` gathers every rank's two
    strips everywhere (world x the bytes, same result).
    """
    import torch
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return None, None
    halo_top = torch.empty_like(bot_edge) if rank > 0 else None
    halo_bot = torch.empty_like(top_edge) if rank < world - 1 else None
    if mode == 'allgather':
        both = torch.stack([top_edge, bot_edge]).contiguous()
        gathered = [torch.empty_like(both) for _ in range(world)]
        dist.all_gather(gathered, both, group=group)
        if rank > 0:
            halo_top.copy_(gathered[rank - 1][1])
        if rank < world - 1:
            halo_bot.copy_(gathered[rank + 1][0])
        return halo_top, halo_bot
    if mode != 'p2p':
        raise ValueError("mode must be 'p2p' or 'allgather'")
    ops = []
    peer = lambda r: dist.get_global_rank(group, r) if group is not None else r
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, top_edge.contiguous(), peer(rank - 1), group))
        ops.append(dist.P2POp(dist.irecv, halo_top, peer(rank - 1), group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, bot_edge.contiguous(), peer(rank + 1), group))
        ops.append(dist.P2POp(dist.irecv, halo_bot, peer(rank + 1), group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return halo_top, halo_bot


_PEER = {'buffers': {}, 'broken': None}


def _peer_buffer(shape, device, group):
    """A symmetric (same shape on every rank) float32 buffer with every peer's copy mapped into this process:
    (tensor, handle).  Cached per shape: the allocation and the rendezvous (an exchange of memory handles) happen once."""
    import torch
    import torch.distributed._symmetric_memory as symm
    dist = _dist()
    key = (tuple(shape), str(device), id(group))
    if key not in _PEER['buffers']:
        t = symm.empty(*shape, dtype=torch.float32, device=device)
        hdl = symm.rendezvous(t, group=group if group is not None else dist.group.WORLD)
        _PEER['buffers'][key] = (t, hdl)
    return _PEER['buffers'][key]


def release_peer_buffers():
    _PEER['buffers'].clear()


def channel_row_pointers(buffer_ptrs, cparts, ny_total, nx, y0, itemsize=4):
    """For every channel j of a cube that is being handed from row shards to channel shards: the address of
    (channel j - c0(owner), row y0, column 0) in the OWNER's (chans, ny_total, nx) buffer -- ``buffer_ptrs[d]`` is rank d's
    buffer as mapped into this process, ``cparts[d] = (c0, c1)`` the channels rank d owns.  int64 array of c1(last) entries:
    the table `sc_spectral_interp_scatter` takes."""
    nout = cparts[-1][1]
    plane = int(ny_total) * int(nx) * itemsize
    ptrs = np.empty(nout, dtype=np.int64)
    for d, (a, b) in enumerate(cparts):
        ptrs[a:b] = int(buffer_ptrs[d]) + np.arange(b - a, dtype=np.int64) * plane + int(y0) * int(nx) * itemsize
    return ptrs


def _peer_available(device, group):
    """Can this job map peer buffers (torch symmetric memory)?  Tried once; the reason it cannot is kept for the warning."""
    if _PEER['broken'] is not None:
        return False
    try:
        _peer_buffer((1, 1, 4), device, group)
        return True
    except Exception as exc:
        import warnings
        _PEER['broken'] = repr(exc)
        warnings.warn("peer-memory buffers unavailable (%s); using NCCL" % (_PEER['broken'][:200],))
        return False


def _reshard_rows_to_channels_peer(local, ny_total, group, borrow):
    """One kernel per rank stores every channel of the local row block straight into its final place in the
    destination rank's buffer over NVLink peer memory (`sc_reshard_scatter`): no staging, no concatenation pass."""
    import ctypes as C
    import torch
    from . import _lib
    dist = _dist()
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nchan, rows, nx = local.shape
    cparts = channel_partition(nchan, world)
    y0 = row_partition(ny_total, world)[rank][0]
    chans_max = max(c1 - c0 for c0, c1 in cparts)
    buf, hdl = _peer_buffer((chans_max, ny_total, nx), local.device, group)
    ptrs = (C.c_uint64 * world)(*[int(p) for p in hdl.buffer_ptrs])
    bounds = (C.c_int64 * (world + 1))(*([c0 for c0, _ in cparts] + [nchan]))
    with _lib.on_device_of(local):
        stream = torch.cuda.current_stream().cuda_stream
        hdl.barrier()                                     # every rank is done reading the buffer's previous contents
        _lib.check(lib.sc_reshard_scatter(local.data_ptr(), nchan, rows, nx, local.stride(0), local.stride(1),
                                          ptrs, world, rank, bounds, ny_total, y0, stream))
        hdl.barrier()                                     # every rank's stores have landed
    mine = buf[:cparts[rank][1] - cparts[rank][0]]
    return mine if borrow else mine.clone()


def reshard_rows_to_channels(local, ny_total, group=None, mode='auto', borrow=False):
    """(nchan, rows_r, nx) row shard -> (chans_r, ny_total, nx) channel shard.

    ``mode='peer'``: one scatter kernel per rank over NVLink peer memory (needs CUDA tensors, float32, nx % 4 == 0 and
    torch's symmetric memory); ``mode='nccl'``: one NCCL all-to-all into per-source blocks plus a concatenation pass;
    ``'auto'`` takes the peer path when it is available.  ``borrow=True`` returns a view of the (reused) peer buffer,
    valid until the next re-shard of the same shape."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nchan, _, nx = local.shape
    cparts = channel_partition(nchan, world)
    rparts = row_partition(ny_total, world)
    if world == 1:
        return local
    if mode not in ('auto', 'peer', 'nccl'):
        raise ValueError("mode must be 'auto', 'peer' or 'nccl'")
    peer_ok = (local.is_cuda and local.dtype == torch.float32 and nx % 4 == 0 and local.stride(2) == 1 and
               local.stride(0) % 4 == 0 and local.stride(1) % 4 == 0 and world <= 16 and dist.get_backend(group) == 'nccl')
    if mode == 'peer' and not peer_ok:
        raise ValueError("the peer re-shard needs CUDA float32 rows of a multiple of 4 samples under an NCCL group of <= 16 ranks")
    if mode != 'nccl' and peer_ok and _PEER['broken'] is None:
        try:
            return _reshard_rows_to_channels_peer(local, ny_total, group, borrow)
        except Exception as exc:                           # no symmetric memory on this platform: say so once, use NCCL
            if mode == 'peer':
                raise
            import warnings
            _PEER['broken'] = repr(exc)
            warnings.warn("peer-memory re-shard unavailable (%s); using the NCCL all-to-all" % (_PEER['broken'][:200],))
    send = [local[c0:c1].contiguous() for (c0, c1) in cparts]
    my_c0, my_c1 = cparts[rank]
    recv = [torch.empty((my_c1 - my_c0, y1 - y0, nx), dtype=local.dtype, device=local.device) for (y0, y1) in rparts]
    dist.all_to_all(recv, send, group=group) if dist.get_backend(group) == 'nccl' else _all_to_all_fallback(recv, send, group)
    return torch.cat(recv, dim=1)


def reshard_channels_to_rows(local, nchan_total, group=None):
    """(chans_r, ny, nx) channel shard -> (nchan_total, rows_r, nx) row shard."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    _, ny, nx = local.shape
    cparts = channel_partition(nchan_total, world)
    rparts = row_partition(ny, world)
    if world == 1:
        return local
    send = [local[:, y0:y1].contiguous() for (y0, y1) in rparts]
    my_y0, my_y1 = rparts[rank]
    recv = [torch.empty((c1 - c0, my_y1 - my_y0, nx), dtype=local.dtype, device=local.device) for (c0, c1) in cparts]
    dist.all_to_all(recv, send, group=group) if dist.get_backend(group) == 'nccl' else _all_to_all_fallback(recv, send, group)
    return torch.cat(recv, dim=0)


def _all_to_all_fallback(recv, send, group):
    """gloo has no all_to_all: pairwise isend/irecv (CPU tests only)."""
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    reqs = []
    for r in range(world):
        if r == rank:
            recv[r].copy_(send[r])
            continue
        reqs.append(dist.isend(send[r], r, group=group))
        reqs.append(dist.irecv(recv[r], r, group=group))
    for q in reqs:
        q.wait()


def gather_rows(local_map, ny_total, group=None, dst=None):
    """Assemble row-sharded 2-D maps (rows_r, nx) into the full (ny_total, nx) map (all ranks, or only
    ``dst``)."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return local_map
    parts = row_partition(ny_total, world)
    nx = local_map.shape[-1]
    pad = max(y1 - y0 for y0, y1 in parts)
    buf = torch.zeros((pad, nx), dtype=local_map.dtype, device=local_map.device)
    buf[:local_map.shape[0]] = local_map
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    if dst is not None and rank != dst:
        return None
    return torch.cat([g[:(y1 - y0)] for g, (y0, y1) in zip(gathered, parts)], dim=0)


class RowShardedCube(object):
    """This rank's row block of a cube that is sharded over the spatial plane.

    ``local`` is a ``SpectralCube`` / ``DaskSpectralCube`` holding rows [y0, y1) of the full
    (nchan, ny_total, nx) cube; its WCS is the FULL cube's WCS with CRPIX2 shifted by -y0, so world
    coordinates of local pixels are the global ones.
    """

    def __init__(self, local, ny_total, y0, group=None):
        self.local, self.ny_total, self.y0, self.group = local, int(ny_total), int(y0), group

    @classmethod
    def from_full_wcs(cls, cube_cls, local_data, full_wcs, ny_total, group=None, **kw):
        dist = _dist()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        y0, y1 = row_partition(ny_total, world)[rank]
        assert local_data.shape[1] == y1 - y0, "local block has %d rows, expected %d" % (local_data.shape[1], y1 - y0)
        w = full_wcs.copy()
        w.crpix[1] -= y0
        return cls(cube_cls(local_data, w, **kw), ny_total, y0, group)

    @property
    def shape(self):
        return (self.local.shape[0], self.ny_total, self.local.shape[2])

    def _wrap(self, local):
        return RowShardedCube(local, self.ny_total, self.y0, self.group)

    def with_mask(self, mask, **kw):
        return self._wrap(self.local.with_mask(mask, **kw))

    def __gt__(self, v):
        return self.local > v

    def __lt__(self, v):
        return self.local < v

    # ---- embarrassingly parallel over spaxels: no communication ------------------------------------
    def moment(self, order=0, axis=0, gather=False, **kw):
        if axis != 0:
            raise NotImplementedError("row-sharded moments are along the spectral axis")
        m = self.local.moment(order=order, axis=0, **kw)
        return m if not gather else self.gather_map(m)

    def moment0(self, **kw):
        return self.moment(order=0, **kw)

    def moment1(self, **kw):
        return self.moment(order=1, **kw)

    def moment2(self, **kw):
        return self.moment(order=2, **kw)

    def gather_map(self, proj):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(proj.value))
        dev = self.local._data.device
        full = gather_rows(t.to(dev), self.ny_total, self.group)
        return full.cpu().numpy()

    def spectral_smooth(self, kernel, **kw):
        return self._wrap(self.local.spectral_smooth(kernel, **kw))

    def spectral_interpolate(self, grid, **kw):
        return self._wrap(self.local.spectral_interpolate(grid, **kw))

    def spectral_interpolate_to_channels(self, grid, mode='auto', **kw):
        """``spectral_interpolate`` whose result is sharded over CHANNELS: (cube of this rank's output channels over
        the whole image, (c0, c1)) -- what ``reproject`` needs next (``cube.spectral_interpolate(grid).reproject(hdr)``
        on a row-sharded cube is config 5).  ``mode='peer'``: the interpolation kernel stores every output channel
        straight into its owner's buffer over NVLink peer memory (`sc_spectral_interp_scatter`): its stores ARE the
        re-shard, the row-sharded intermediate is never written.  ``mode='twostep'``: interpolate locally, then
        ``reshard_rows_to_channels``.  ``'auto'`` takes the peer path when it is available.  The data arrive NaN-filled
        where the interpolated mask excludes them (no mask travels), so a non-NaN fill value needs ``'twostep'``."""
        import torch
        dist = _dist()
        loc = self.local
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        nout = int(np.asarray(getattr(grid, 'value', grid)).size)
        nchan, rows, nx = loc.shape
        if mode not in ('auto', 'peer', 'twostep'):
            raise ValueError("mode must be 'auto', 'peer' or 'twostep'")
        src = loc._data
        peer_ok = (world > 1 and world <= 16 and src.is_cuda and nx % 4 == 0 and dist.get_backend(self.group) == 'nccl' and
                   loc._interp_nan_filled(kw.get('fill_value')))
        if mode == 'peer' and not peer_ok:
            raise ValueError("the peer path needs an NCCL job of 2..16 CUDA ranks, nx % 4 == 0 and a NaN fill value")
        cparts = channel_partition(nout, world)
        c0, c1 = cparts[rank]
        if mode == 'twostep' or not peer_ok or (mode == 'auto' and not _peer_available(src.device, self.group)):
            reshard_mode = kw.pop('reshard_mode', 'auto')
            rows_cube = loc.spectral_interpolate(grid, **kw)
            chan_local = reshard_rows_to_channels(rows_cube._filled_tensor(rows_cube._fill_value), self.ny_total, self.group,
                                                  mode=reshard_mode, borrow=True)
            newwcs = rows_cube._wcs.copy()
        else:
            chans_max = max(b - a for a, b in cparts)
            buf, hdl = _peer_buffer((chans_max, self.ny_total, nx), src.device, self.group)
            chan_ptrs = torch.from_numpy(channel_row_pointers(hdl.buffer_ptrs, cparts, self.ny_total, nx, self.y0)).to(src.device)
            from . import _lib
            kw.pop('reshard_mode', None)
            with _lib.on_device_of(src):
                hdl.barrier()                                 # every rank is done with the buffer's previous contents
                newwcs = loc._spectral_interpolate_scatter(grid, chan_ptrs, phase=rank, nphases=world, **kw)
                hdl.barrier()                                 # every rank's stores have landed
            chan_local = buf[:c1 - c0]
        w = newwcs
        w.crpix[1] += self.y0                                  # back to the full image's WCS
        w.crpix[2] -= c0
        sub = type(loc)(chan_local, w, unit=loc._unit, fill_value=loc._fill_value, spectral_unit=loc._spectral_unit,
                        allow_huge_operations=loc.allow_huge_operations)
        sub._nan_filled_already = True
        return sub, (c0, c1)

    # ---- spatial_smooth / convolve_to: halo rows from the two neighbours ------------------------------------
    def _is_sharded(self):
        dist = _dist()
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _blank_planes(self):
        """Device uint8 (nchan): 1 where the mask includes nothing of the WHOLE image plane (all shards) -- the test
        of the numpy class's `_apply_spatial_function` (spectral_cube.py:161-172).  None without a mask."""
        import torch
        loc = self.local
        if loc._mask is None:
            return None
        inc = loc._mask._include_tensor(loc._data)
        # (`.any()` of a uint8 tensor is a uint8 tensor, and `~` of that is a bitwise complement: go through bool)
        blank = (~inc.bool().reshape(inc.shape[0], -1).any(dim=1)).to(torch.int32)
        if self._is_sharded():
            _dist().all_reduce(blank, op=_dist().ReduceOp.MIN, group=self.group)
        return blank.to(torch.uint8)

    def _smooth_rows(self, k2d, halo_mode='p2p', counts=None, every_tap=False, need_blank=False):
        """The spatial kernels on this rank's rows with the neighbours' filled edge rows as halos.
        Returns (float32 result, job-wide blank-plane flags or None)."""
        import torch
        from . import _lib
        lib = _lib.load()
        loc = self.local
        h = k2d.shape[0] // 2
        nchan, ny, nx = loc.shape
        if h > ny:
            raise ValueError("kernel half-height %d exceeds the %d rows of this shard" % (h, ny))
        src = loc._data
        desc, keep = loc._mask_desc()
        stream = torch.cuda.current_stream().cuda_stream

        def pack(row0):
            out = torch.empty((nchan, h, nx), dtype=torch.float32, device=src.device)
            _lib.check(lib.sc_pack_filled_rows(src.data_ptr(), nchan, ny, nx, src.stride(0), src.stride(1), desc,
                                               float(loc._fill_value), row0, h, out.data_ptr(), stream))
            return out
        halo_top = halo_bot = None
        dask = loc._mirrors_dask
        sharded = self._is_sharded()
        if h > 0 and sharded:
            halo_top, halo_bot = exchange_halo_rows(pack(0), pack(ny - h), self.group, mode=halo_mode)
        if counts is None:
            # one denominator strategy for the whole job (bit-identical with the unsharded result): the ranks'
            # sample counts are summed (8 bytes; the only other exchange of this op are the halo rows)
            counts = loc._spatial_strategy_counts()
            if sharded:
                _dist().all_reduce(counts, op=_dist().ReduceOp.SUM, group=self.group)
        # numpy class: `_apply_spatial_function` copies a channel image through when the mask includes nothing of
        # the WHOLE image (spectral_cube.py:161-172).  A shard cannot take that decision on its own rows -- a plane
        # blank here may have data next door that reaches across the boundary through the halo rows -- so shards
        # always convolve.  A plane blank everywhere convolves to what the copy would give for the usual fill values
        # (NaN stays NaN, 0 stays 0); for any other finite fill the job-wide blank planes are copied afterwards.
        passthrough = 0 if (sharded and not dask) else None
        out = loc._run_spatial_smooth(k2d, _lib.F32, halo_top=halo_top, halo_bot=halo_bot, halo_rows=h,
                                      strategy_counts=counts, passthrough=passthrough, every_tap=every_tap)
        fill = float(loc._fill_value)
        blank = None
        if not dask:
            if not sharded:
                blank = loc._passthrough_flags                      # what the library found for this (whole) image
            elif need_blank or (np.isfinite(fill) and fill != 0.0):
                blank = self._blank_planes()
                if blank is not None and np.isfinite(fill) and fill != 0.0 and bool(blank.any()):
                    out[blank.bool()] = loc._filled_tensor(fill)[blank.bool()]
        return out, blank

    def spatial_smooth(self, kernel, halo_mode='p2p', **kw):
        loc = self.local
        loc.check_jybeam_smoothing(raise_error_jybm=kw.pop('raise_error_jybm', True))
        out, _ = self._smooth_rows(loc._kernel_array(kernel, 2), halo_mode=halo_mode)
        new = loc._new_cube_with(data=out) if loc._mirrors_dask else loc._new_cube_reporting_f64(out)
        return self._wrap(new)

    def convolve_to(self, beam, convolve=None, halo_mode='p2p', **kw):
        """``SpectralCube.convolve_to`` on a row-sharded cube (spectral_cube.py:3335-3392; dask :1412-1464): the
        deconvolved beam's kernel through the halo-exchanging spatial kernels, then the in-place epilogue (Jy/beam
        rescale, ``convolve_fft``'s zero for empty windows) on this rank's rows -- no other communication."""
        import warnings
        from .beam import Beam
        loc = self.local
        loc._check_convolve_kwargs(kw)
        beam = Beam.coerce(beam)
        if beam == loc.beam:
            warnings.warn("The given beam is identical to the current beam. "
                          "Skipping convolution.")
            return self
        kernel = beam.deconvolve(loc.beam).as_kernel(loc._pixscale_deg())
        factor = beam.sr / loc.beam.sr if loc._is_jybeam() else 1.0
        fft = loc._fft_semantics(convolve, default=not loc._mirrors_dask)
        out, blank = self._smooth_rows(loc._kernel_array(kernel, 2), halo_mode=halo_mode, every_tap=True,
                                       counts=loc._convolved_denominator_counts(),
                                       need_blank=(factor != 1.0 or fft))
        loc._convolve_epilogue(out, factor, fft, skip=blank)
        new = loc._new_cube_with(data=out) if loc._mirrors_dask else loc._new_cube_reporting_f64(out)
        new._attach_beam(beam)
        return self._wrap(new)

    # ---- reproject: one re-shard to channels, then whole planes locally -------------------------------
    def reproject(self, header, order='bilinear', filled=True, **kw):
        """Returns a ``ChannelShardedCube``-like pair: (local cube of this rank's channels over the
        full output plane, (c0, c1)).  Masked voxels are filled BEFORE the re-shard so lazy masks
        (which refer to the row-sharded data) never have to travel."""
        dist = _dist()
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        loc = self.local
        filled_rows = loc._filled_tensor(loc._fill_value) if filled else loc._data
        chan_local = reshard_rows_to_channels(filled_rows, self.ny_total, self.group, mode=kw.pop('reshard_mode', 'auto'),
                                              borrow=True)     # consumed by the reproject call below
        c0, c1 = channel_partition(loc.shape[0], world)[rank]
        w = loc._wcs.copy()
        w.crpix[1] += self.y0                                  # back to the full image's WCS
        w.crpix[2] -= c0
        sub = type(loc)(chan_local, w, unit=loc._unit, fill_value=loc._fill_value, spectral_unit=loc._spectral_unit,
                        allow_huge_operations=loc.allow_huge_operations)
        return self._reproject_channel_shard(sub, c0, c1, header, order, **kw), (c0, c1)

    @staticmethod
    def _reproject_channel_shard(sub, c0, c1, header, order='bilinear', **kw):
        """`reproject` of one rank's channels: the target header / WCS restricted to channels [c0, c1)."""
        if hasattr(header, 'celestial_params'):
            so = tuple(kw.pop('shape_out'))
            kw['shape_out'] = (c1 - c0,) + so[1:]
            return sub.reproject(header, order=order, filled=False, **kw)
        hdr = dict(header)
        hdr['NAXIS3'] = c1 - c0
        hdr['CRPIX3'] = hdr.get('CRPIX3', 1.0) - c0
        return sub.reproject(hdr, order=order, filled=False, **kw)

    def spectral_interpolate_reproject(self, grid, header, order='bilinear', mode='auto', **kw):
        """Config 5 on a row-sharded cube, ``cube.spectral_interpolate(grid).reproject(header)``: the interpolation
        kernel scatters its output to the channel owners (``spectral_interpolate_to_channels``), every rank reprojects
        its channels.  Returns (cube of this rank's channels over the output image, (c0, c1))."""
        interp_kw = {k: kw.pop(k) for k in ('suppress_smooth_warning', 'fill_value') if k in kw}
        sub, (c0, c1) = self.spectral_interpolate_to_channels(grid, mode=mode, **interp_kw)
        return self._reproject_channel_shard(sub, c0, c1, header, order, **kw), (c0, c1)
