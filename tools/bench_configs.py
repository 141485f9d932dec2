"""
Throughput of every BASELINE.json config beyond the headline one (bench.py covers configs[1]).
Run on the GPU box:   python tools/bench_configs.py [c1] [c3] [c4] [c5] [reduce] [ingest] [convolve] [target]
              or   python -m torch.distributed.run --nproc-per-node N ... tools/bench_configs.py c4 c5
Prints one JSON line per measurement (rank 0).  Timing: CUDA events, warm-up 2, mean of 5, max over ranks.
"""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import spectral_cube_b200 as scb
from spectral_cube_b200 import _lib, distributed as D
from spectral_cube_b200.synth import synth_cube, benchmark_wcs

rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) else 6650.0


def timeit(f, n=5, warm=2):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def emit(config, what, voxels, ms, algo_bytes, **extra):
    if rank == 0:
        line = dict(config=config, what=what, n_gpus=world, voxels=voxels, ms=ms, voxels_per_s=voxels / ms * 1e3,
                    algorithmic_GBps_per_gpu=algo_bytes / world / ms / 1e6, frac_of_measured_hbm=algo_bytes / world / ms / 1e6 / PEAK)
        line.update(extra)
        print(json.dumps(line), flush=True)


def isfinite_cube(cls, dev, w):
    c = cls(dev, w, unit='K', allow_huge_operations=True)
    c._mask = scb.LazyMask(np.isfinite, cube=c)
    return c


which = [a for a in sys.argv[1:]] or ['c1', 'c3', 'c4', 'c5']      # 'target' (137 GB on one GPU) only on request

if 'c1' in which and world == 1:
    nchan, ny, nx = 128, 256, 256
    dev = synth_cube(nchan, ny, nx, nan_permille=0, border=0)
    c = scb.SpectralCube(dev, benchmark_wcs(nchan, ny, nx), unit='K')       # no mask (configs[0])
    ms = timeit(lambda: c._moments_axis0_raw(1), n=50, warm=5)
    V = nchan * ny * nx
    emit('c1', 'moment0, no mask, 256x256x128 (33.6 MB: L2 resident)', V, ms, 4 * V + 8 * ny * nx)
    # (the CPU leg of this config -- the oracle, cube strategy, one thread -- is `python bench.py --config0`: only
    #  bench.py, the tests and smoke() may execute oracle/)
    del dev, c

if 'c3' in which and world == 1:
    nchan, ny, nx = 1024, 2048, 2048
    V = nchan * ny * nx
    dev = synth_cube(nchan, ny, nx, border=51)
    c = isfinite_cube(scb.DaskSpectralCube, dev, benchmark_wcs(nchan, ny, nx))
    k = scb.Gaussian1DKernel(5 / 2.3548200450309493)
    sm = c.spectral_smooth(k)
    ms = timeit(lambda: sm._moments_axis0_raw(2))
    emit('c3', 'spectral_smooth(FWHM 5 ch, 17 taps) -> moment1, FUSED (smoothed cube never written)', V, ms, 4 * V + 8 * ny * nx)
    ms_s = timeit(lambda: c._run_spectral_smooth(k.array, _lib.F32))
    emit('c3', 'spectral_smooth alone, float32 out (materialised)', V, ms_s, 8 * V)
    mat = c.spectral_smooth(k, save_to_tmp_dir=True)
    ms_m = timeit(lambda: mat._moments_axis0_raw(2))
    emit('c3', 'moment1 of the materialised smoothed cube (mask refers to the OLD data: +4 B/voxel)', V, ms_m, 8 * V + 8 * ny * nx)
    emit('c3', 'unfused total (smooth + moment1)', V, ms_s + ms_m, 16 * V + 8 * ny * nx)
    del dev, c, sm, mat
    torch.cuda.empty_cache()

if 'c4' in which:
    nchan, ny, nx = 512, 4096, 4096
    V = nchan * ny * nx
    y0, y1 = D.row_partition(ny, world)[rank]
    dev = synth_cube(nchan, y1 - y0, nx, y0=y0, ny_total=ny, nx_total=nx, border=102)
    w = benchmark_wcs(nchan, ny, nx)
    k = scb.Gaussian2DKernel(8 / 2.3548200450309493)
    if world == 1:
        c = isfinite_cube(scb.DaskSpectralCube, dev, w)
        ms = timeit(lambda: c._run_spatial_smooth(k.array, _lib.F32), n=3, warm=1)
        emit('c4', 'spatial_smooth(FWHM 8 px, 29x29 separable), whole 4096x4096x512 cube on one GPU', V, ms, 8 * V)
    else:
        sh = D.RowShardedCube.from_full_wcs(scb.DaskSpectralCube, dev, w, ny, unit='K')
        sh.local._mask = scb.LazyMask(np.isfinite, cube=sh.local)
        for mode in ('p2p', 'allgather'):
            ms = timeit(lambda: sh.spatial_smooth(k, halo_mode=mode), n=3, warm=1)
            emit('c4', 'spatial_smooth(FWHM 8 px) row-sharded, halo exchange = %s (exchange inside the timed region)' % mode, V, ms, 8 * V)
    del dev
    torch.cuda.empty_cache()

if 'c5' in which:
    nchan, ny, nx = 2048, 4096, 4096
    V = nchan * ny * nx
    nout = 1024
    # one GPU cannot hold the 137 GB cube plus its outputs: N = 1 runs the 1/8 row shard an 8-GPU job gives each rank
    shards = world if world >= 4 else 8        # fewer than 4 GPUs cannot hold the whole cube: time one of 8 shards
    y0, y1 = D.row_partition(ny, shards)[rank]
    dev = synth_cube(nchan, y1 - y0, nx, y0=y0, ny_total=ny, nx_total=nx, border=102)
    w = benchmark_wcs(nchan, ny, nx)
    wl = w.copy(); wl.crpix[1] -= y0
    c = isfinite_cube(scb.SpectralCube, dev, wl)
    sa = c.spectral_axis
    grid = np.linspace(sa[0], sa[-1], nout)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ms_i = timeit(lambda: c.spectral_interpolate(grid), n=3, warm=1)
        Vs = nchan * (y1 - y0) * nx
        emit('c5', 'spectral_interpolate 2048 -> 1024 channels, %d-row shard (1/%d of the cube) per GPU' % (y1 - y0, shards),
             Vs * world, ms_i, world * (4 * Vs + 5 * Vs // 2))
        interp = c.spectral_interpolate(grid)
    del dev, c
    torch.cuda.empty_cache()
    a = np.radians(30.0)
    hdr = dict(w.to_header())
    hdr.update({'NAXIS': 3, 'NAXIS1': nx, 'NAXIS2': ny, 'NAXIS3': nout, 'PC1_1': np.cos(a), 'PC1_2': -np.sin(a),
                'PC2_1': np.sin(a), 'PC2_2': np.cos(a)})
    if world < 4:
      if world == 1:
        # the channel shard a rank holds after the rows->channels re-shard: 128 whole planes
        nloc = nout // shards
        planes = synth_cube(nloc, ny, nx, border=102)
        cc = isfinite_cube(scb.SpectralCube, planes, w)
        hdr['NAXIS3'] = nloc
        ms_r = timeit(lambda: cc.reproject(hdr), n=3, warm=1)
        emit('c5', 'reproject (bilinear, WCS rotated 30 deg, incl. pixel map) of a %d-plane channel shard 4096x4096' % nloc,
             nloc * ny * nx, ms_r, 4 * nloc * ny * nx + 9 * nloc * ny * nx + 16 * ny * nx)
    else:
        sh = D.RowShardedCube(interp, ny, y0, None)
        t0 = timeit(lambda: D.reshard_rows_to_channels(interp._data, ny), n=3, warm=1)
        emit('c5', 'rows -> channels re-shard (all-to-all) of the interpolated cube', nout * ny * nx, t0, 0)
        ms_r = timeit(lambda: sh.reproject(hdr), n=2, warm=1)
        emit('c5', 'reproject of the row-sharded interpolated cube (fill + all-to-all + pixel map + bilinear)', nout * ny * nx, ms_r,
             (4 + 9) * nout * ny * nx)
        emit('c5', 'config 5 total: spectral_interpolate + reproject', V, ms_i + ms_r, 0)

if 'reduce' in which and world == 1:
    # the noise / peak maps behind a ">3 sigma" mask, one pass for all seven statistics
    nchan, ny, nx = 1024, 2048, 2048
    V = nchan * ny * nx
    dev = synth_cube(nchan, ny, nx, border=51)
    c = isfinite_cube(scb.SpectralCube, dev, benchmark_wcs(nchan, ny, nx))
    allstats = {'sum', 'count', 'm2', 'min', 'max', 'argmin', 'argmax'}
    ms = timeit(lambda: c._reduce_axis0_raw(allstats))
    emit('reduce', 'sum+count+m2+min+max+argmin+argmax along the spectral axis in one pass, 2048x2048x1024', V, ms, 4 * V + 36 * ny * nx)
    ms = timeit(lambda: c._reduce_axis0_raw({'max'}))
    emit('reduce', 'max along the spectral axis (peak map)', V, ms, 4 * V + 4 * ny * nx)
    if os.environ.get('SC_REDUCE_SPATIAL') == '1':                     # opt-in kernels (sc_reduce_spatial)
        for axis in (1, 2):
            ms = timeit(lambda: c._reduce_spatial_raw(axis, allstats))
            emit('reduce', 'all seven statistics along numpy axis %d (sc_reduce_spatial)' % axis, V, ms, 4 * V)
    del dev, c
    torch.cuda.empty_cache()

if 'convolve' in which and world == 1:
    # convolve_to.  Round beams -> the separable kernels (+ the sc_scale epilogue for Jy/beam),
    # rotated elliptical beams -> direct2d_kernel, per-channel beams -> one launch pair per plane.
    nchan, ny, nx = 64, 4096, 4096
    V = nchan * ny * nx
    dev = synth_cube(nchan, ny, nx, border=102)
    w = benchmark_wcs(nchan, ny, nx)
    pix = float(np.sqrt(abs(np.linalg.det(w.pixel_scale_matrix[:2, :2]))))
    for unit in ('K', 'Jy/beam'):
        c = scb.SpectralCube(dev, w, unit=unit, beam=scb.Beam(3 * pix))
        c._mask = scb.LazyMask(np.isfinite, cube=c)
        ms = timeit(lambda: c.convolve_to(scb.Beam(5 * pix)), n=3, warm=1)
        emit('convolve', 'convolve_to, round 3 px -> 5 px beam (29x29 separable, convolved denominator), unit %s, 4096x4096x64' % unit,
             V, ms, (8 + 8) * V)
        c = scb.SpectralCube(dev, w, unit=unit, beam=scb.Beam(3 * pix, 2 * pix, 60.0))
        c._mask = scb.LazyMask(np.isfinite, cube=c)
        k = scb.Beam(5 * pix, 4 * pix, 25.0).deconvolve(c.beam).as_kernel(pix)
        for tiled in ('0', '1'):
            os.environ['SC_DIRECT2D'] = tiled
            ms = timeit(lambda: c.convolve_to(scb.Beam(5 * pix, 4 * pix, 25.0)), n=2, warm=1)
            emit('convolve', 'convolve_to, rotated elliptical beams (%dx%d taps, %s), unit %s'
                 % (k.shape + ('direct2d_tiled_kernel' if tiled == '1' else 'direct2d_kernel', unit)), V, ms, (8 + 8) * V)
        os.environ.pop('SC_DIRECT2D', None)
    # a round beam too wide for the separable kernels (half-width > 16): the outer product on the direct path
    c = scb.SpectralCube(dev, w, unit='K', beam=scb.Beam(3 * pix))
    c._mask = scb.LazyMask(np.isfinite, cube=c)
    k = scb.Beam(8 * pix).deconvolve(c.beam).as_kernel(pix)
    for tiled in ('0', '1'):
        os.environ['SC_DIRECT2D'] = tiled
        ms = timeit(lambda: c.convolve_to(scb.Beam(8 * pix)), n=1, warm=1)
        emit('convolve', 'convolve_to, round 3 px -> 8 px beam (%dx%d outer product, %s)'
             % (k.shape + ('direct2d_tiled_kernel' if tiled == '1' else 'direct2d_kernel',)), V, ms, (8 + 8) * V)
    os.environ.pop('SC_DIRECT2D', None)
    rng = np.random.default_rng(0)
    beams = scb.Beams(major=rng.uniform(2.5, 3.5, nchan) * pix, minor=rng.uniform(2.0, 2.5, nchan) * pix, pa=rng.uniform(0, 180, nchan))
    vr = scb.VaryingResolutionSpectralCube(dev, w, unit='Jy/beam', beams=beams)
    vr._mask = scb.LazyMask(np.isfinite, cube=vr)
    ms = timeit(lambda: vr.convolve_to(scb.Beam(5 * pix)), n=2, warm=1)
    emit('convolve', 'VaryingResolutionSpectralCube.convolve_to, %d per-channel elliptical beams -> round 5 px, Jy/beam' % nchan, V, ms, (8 + 8) * V)
    lib = _lib.load()
    out = torch.empty((nchan, ny, nx), dtype=torch.float32, device='cuda')
    ms = timeit(lambda: _lib.check(lib.sc_scale(out.data_ptr(), _lib.F32, nchan, ny, nx, out.stride(0), out.stride(1), 1.0000001, 1, None,
                                                torch.cuda.current_stream().cuda_stream)))
    emit('convolve', 'sc_scale alone (in place, float32)', V, ms, 8 * V)
    del dev, c, vr, out
    torch.cuda.empty_cache()

if 'ingest' in which and world == 1:
    # FITS file (page cache) -> pinned staging -> device, decoded on the device
    import tempfile
    from spectral_cube_b200 import io_fits
    nchan, ny, nx = 256, 1024, 2048                       # 2.1 GB of big-endian float32
    V = nchan * ny * nx
    host = np.random.default_rng(0).normal(0, 1, (nchan, ny, nx)).astype(np.float32)
    path = os.path.join(tempfile.gettempdir(), 'sc_b200_ingest.fits')
    io_fits.write_fits(path, host, dict(benchmark_wcs(nchan, ny, nx).to_header()), overwrite=True)
    t0 = time.perf_counter(); cube = scb.SpectralCube.read(path); torch.cuda.synchronize(); t1 = time.perf_counter()
    t0 = time.perf_counter(); cube = scb.SpectralCube.read(path); torch.cuda.synchronize(); t1 = time.perf_counter()
    ok = bool(torch.equal(cube._data.cpu(), torch.from_numpy(host)))
    emit('ingest', 'SpectralCube.read of a 2.1 GB FITS cube (page cache -> pinned -> device, decode on device), wall clock', V, (t1 - t0) * 1e3, 4 * V, bit_exact=ok)
    raw = torch.from_numpy(host.astype('>f4').view(np.uint8).reshape(-1)).cuda()
    out = torch.empty((nchan, ny, nx), dtype=torch.float32, device='cuda')
    lib = _lib.load()
    ms = timeit(lambda: _lib.check(lib.sc_fits_decode(raw.data_ptr(), out.data_ptr(), V, -32, 1.0, 0.0, 0, 0, torch.cuda.current_stream().cuda_stream)))
    emit('ingest', 'sc_fits_decode alone (device-resident raw bytes -> float32)', V, ms, 8 * V)
    t0 = time.perf_counter(); ref = np.fromfile(path, dtype='>f4', offset=2880 * 2, count=V).astype(np.float32); t1 = time.perf_counter()
    if rank == 0:
        print(json.dumps(dict(config='ingest', what='CPU: numpy fromfile + byte swap of the same file, 1 thread', voxels=V, ms=(t1 - t0) * 1e3)), flush=True)
    os.unlink(path)
    del host, cube, raw, out
    torch.cuda.empty_cache()

if 'target' in which:
    # north-star target: moment0/1/2 + spectral_smooth on the 4096x4096x2048 cube.  One GPU holds the whole
    # 137.4 GB cube (smoothing runs in place); N GPUs hold 1/N of the rows each.
    nchan, ny, nx = 2048, 4096, 4096
    y0, y1 = D.row_partition(ny, world)[rank]
    V = nchan * ny * nx
    dev = synth_cube(nchan, y1 - y0, nx, y0=y0, ny_total=ny, nx_total=nx, border=102)
    w = benchmark_wcs(nchan, ny, nx)
    wl = w.copy(); wl.crpix[1] -= y0
    c = isfinite_cube(scb.DaskSpectralCube, dev, wl)
    c = c.with_mask(c > 3.0)
    S = ny * nx
    for bits, name in ((1, 'moment0'), (2, 'moment1'), (4, 'moment2'), (7, 'moment0+1+2 in one pass')):
        ms = timeit(lambda: c._moments_axis0_raw(bits), n=3, warm=1)
        emit('target', '%s under isfinite & >3 sigma, 4096x4096x2048 (137.4 GB)' % name, V, ms, 4 * V + 8 * S * bin(bits).count('1'))
    k = scb.Gaussian1DKernel(5 / 2.3548200450309493)
    c2 = isfinite_cube(scb.DaskSpectralCube, dev, wl)
    ms = timeit(lambda: c2._run_spectral_smooth(k.array, _lib.F32, out=dev), n=3, warm=1)
    emit('target', 'spectral_smooth FWHM 5 ch (17 taps), float32, IN PLACE, 4096x4096x2048', V, ms, 8 * V)
    del dev, c, c2
    torch.cuda.empty_cache()

if dist is not None:
    dist.destroy_process_group()
