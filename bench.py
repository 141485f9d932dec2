#!/usr/bin/env python
"""
bench.py -- throughput of the per-spaxel hot path (BASELINE.json metric: voxels/sec).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config 1|3|4|5]

ONE JSON line.  Headline (`value`, `e2e`, `roofline`, `cpu_baseline`): BASELINE.json configs[1] -- a
2048x2048x1024 float32 synthetic cube PER GPU, moment0 + moment1 + moment2 under a >3 sigma LazyMask on top of
the isfinite mask (weak scaling over the spatial plane, no data-path collective).  A step = the three drop-in
reductions; `value` = voxels of the cube / step time with the cube resident in HBM (timed on the device through
`SpectralCube._moments_axis0_raw`, the body of `moment()` before the map is copied to the host); `e2e` = the same
step through the public API from a pinned HOST cube (upload + moment0/1/2() + maps back on the host).

Extra keys on the same line (each with its own roofline; a failure in one never loses the headline):
  spectral_smooth   configs[1] cube, Gaussian FWHM 5 channels (17 taps), float32 out
  c3                configs[2]: spectral_smooth -> moment1, fused (DaskSpectralCube) and unfused (SpectralCube)
  c4_strong         configs[3]: the FIXED 4096x4096x512 cube row-sharded over the N ranks through
                    RowShardedCube.spatial_smooth (29x29 Gaussian), halo exchange (NCCL p2p and all-gather)
                    INSIDE the timed region -- strong scaling of the only op with a data-plane exchange
  c5                configs[4] pieces: spectral_interpolate 2048 -> 1024 channels and reproject (30 deg), on the
                    shards an 8-GPU job gives a rank (N < 4), or the whole job incl. the rows -> channels
                    all-to-all (N >= 4)
  target_strong     the north-star cube 4096x4096x2048 (137.4 GB) split over N ranks by rows: moment0/1/2 +
                    spectral_smooth in place; N = 1 holds all of it
  selftest          (N >= 2) sharded results bit-identical to the single-GPU ones (spectral_cube_b200/selftest.py)

`--impl reference` times the CPU restatement of the reference algorithm (oracle/; the reference itself cannot be
installed: astropy/dask/reproject absent) with worker processes on all host cores, on a bounded sample of the
config `--config` names (default 1).  The worker pool is forked BEFORE the timed window.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NCHAN, NY, NX = 1024, 2048, 2048
THRESHOLD = 3.0
BORDER = 51                         # 2.5 % NaN frame on every side
SEED = 247825498
FWHM2SIGMA = 1.0 / 2.3548200450309493
C4_SHAPE = (512, 4096, 4096)
TARGET_SHAPE = (2048, 4096, 4096)


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(profile, kernel_regex):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, read from a committed ncu summary under
    profiles/ (tools/ncu_summary.py --rep): (bytes, file) or (None, None).  Never a literal."""
    path = os.path.join(ROOT, 'profiles', profile)
    if not os.path.exists(path):
        return None, None
    unit = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
    cur, vals = False, {}
    for line in open(path):
        if line.startswith('== '):
            if vals:
                break
            cur = re.search(kernel_regex, line) is not None
        elif cur:
            f = line.split()
            if len(f) >= 3 and f[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum') and f[1] in unit:
                vals[f[0]] = float(f[2]) * unit[f[1]]
    if len(vals) == 2:
        return vals['dram__bytes_read.sum'] + vals['dram__bytes_write.sum'], 'profiles/' + profile
    return None, None


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                 '-lms', '50'], stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(',')]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out['sm_mhz'] = statistics.median(sm)
            out['sm_max_mhz'] = max(mx)
            out['samples'] = len(sm)
        out['reasons'] = sorted(reasons)
        return out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and with it the pages of the pinned staging buffers it allocates afterwards: first touch)
    to the CPUs of the NUMA node the rank's GPU hangs off.  Without it the 8 ranks of a job allocate their pinned
    cubes wherever the launcher happened to run and all uploads cross one memory controller / the socket link."""
    info = {'bound': False}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        info['pci'] = bdf
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bdf).read().strip())
        info['numa_node'] = node
        if node < 0:
            return info
        cpus = set()
        for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info['bound'] = True
            info['cpus'] = len(allowed)
    except Exception as exc:
        info['error'] = repr(exc)[:120]
    return info


WCS_KW = dict(ctype=['RA---TAN', 'DEC--TAN', 'VRAD'], crval=[24.0, 30.0, -321.214698632],
              cdelt=[-5.55555561268e-4, 5.55555561268e-4, 1.28821496879], cunit=['deg', 'deg', 'km/s'])


def wcs_kw(ny_total, nx_total, y0=0):
    kw = dict(WCS_KW)
    kw['crpix'] = [nx_total / 2.0 + 0.5, ny_total / 2.0 + 0.5 - y0, 1.0]
    return kw


# ==== CPU side: the oracle port of the reference algorithm on bounded samples ======================================
# One sample = a list of independent work items (row blocks / channel planes); a pool of forked worker processes
# (dask's 'processes' scheduler: numpy's masked-array path holds the GIL) maps over them.  The pool is created
# before the clock starts; the sample is inherited copy-on-write.
_CPU_SAMPLE = None
_SYNTH_CACHE = {}


def _cached_block(*args, **kw):
    """oracle.synth.synth_block, memoised: the sample of a config is the same every step."""
    from oracle.synth import synth_block
    key = (args, tuple(sorted(kw.items())))
    if key not in _SYNTH_CACHE:
        _SYNTH_CACHE[key] = synth_block(*args, **kw)
    return _SYNTH_CACHE[key]


def _cpu_moments_work(i):
    import warnings
    from oracle.cube import OracleCube
    from oracle.wcs import OWCS
    data, wkw, bounds = _CPU_SAMPLE
    blk = data[:, bounds[i]:bounds[i + 1], :]
    cube = OracleCube(blk, OWCS(**wkw), unit='K')
    cube = cube.with_mask(cube > THRESHOLD)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return [cube.moment(order=o, how='slice')[0].shape for o in (0, 1, 2)]


def _cpu_c3_work(i):
    """configs[2] on the dask class: one 3-d convolution with the (17,1,1) kernel per block
    (dask_spectral_cube.py:880-917), then moment1 (:1031-1132)."""
    import warnings
    from oracle.cube import OracleCube
    from oracle.wcs import OWCS
    from oracle import convolve as oconv
    data, wkw, bounds = _CPU_SAMPLE
    blk = data[:, bounds[i]:bounds[i + 1], :]
    cube = OracleCube(blk, OWCS(**wkw), unit='K', use_dask=True)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        sm = cube.spectral_smooth(oconv.Gaussian1DKernel(5 * FWHM2SIGMA))
        return sm.moment(order=1)[0].shape


def _cpu_c4_work(i):
    """configs[3]: astropy's direct 2-D convolution of one channel image (29x29 taps), here a row block of it with
    its 14 halo rows (spectral_cube.py:2808-2842, dask :962-993)."""
    from oracle import convolve as oconv
    data, h = _CPU_SAMPLE
    out = oconv.convolve(data[i], oconv.Gaussian2DKernel(8 * FWHM2SIGMA), normalize_kernel=True)
    return out[h:-h].shape


def _cpu_c5_interp_work(i):
    """configs[4], first half on the numpy class: two np.interp per spaxel (spectral_cube.py:3298-3310)."""
    import numpy as np
    from oracle import interp as ointerp
    data, inaxis, grid, bounds = _CPU_SAMPLE
    blk = data[:, bounds[i]:bounds[i + 1], :]
    out, mask, rev = ointerp.spectral_interpolate_numpy(blk, np.isfinite(blk), inaxis, grid)
    return out.shape


def _cpu_c5_reproject_work(i):
    """configs[4], second half: reproject_interp's bilinear sampling of one channel image (spectral_cube.py:2726-2732)."""
    from oracle import reproject as orep
    data, yin, xin = _CPU_SAMPLE
    return orep.sample_bilinear(data[i % data.shape[0]], yin, xin).shape


def _timed_pool_map(func, n_items, threads):
    import multiprocessing as mp
    n = max(1, min(threads, n_items))
    if n == 1:
        t0 = time.perf_counter()
        for i in range(n_items):
            func(i)
        return time.perf_counter() - t0, 1
    with mp.get_context('fork').Pool(n) as pool:
        pool.map(_noop, range(n))                   # workers are up before the clock starts
        t0 = time.perf_counter()
        pool.map(func, range(n_items), chunksize=1)
        dt = time.perf_counter() - t0
    return dt, n


def _noop(i):
    return i


def cpu_sample(config, threads, scale=1.0):
    """Time the oracle port of BASELINE.json configs[config] on a bounded sample.  Returns (voxels, seconds,
    description, workers)."""
    global _CPU_SAMPLE
    import numpy as np
    from oracle.synth import synth_block
    # everything the workers use is imported HERE, before the fork: an `import scipy` inside sixteen freshly forked
    # workers would otherwise sit in the timed window
    import oracle.cube, oracle.wcs, oracle.convolve, oracle.interp, oracle.reproject, oracle.moments   # noqa: F401,E401
    if config in (1, 3):
        rows = max(threads, int((128 if config == 1 else 32) * scale))
        y0 = NY // 2 - rows // 2
        data = _cached_block(NCHAN, rows, NX, y0=y0, ny_total=NY, nx_total=NX, seed=SEED, nan_permille=1, border=BORDER)
        nblk = max(1, min(threads, rows))
        bounds = np.linspace(0, rows, nblk + 1).astype(int)
        _CPU_SAMPLE = (data, wcs_kw(NY, NX, y0), bounds)
        dt, n = _timed_pool_map(_cpu_moments_work if config == 1 else _cpu_c3_work, nblk, threads)
        what = 'slicewise moment0+1+2 under isfinite & >3 sigma' if config == 1 else \
               'dask-class spectral_smooth (17 taps, one 3-d convolution per block) + moment1'
        return NCHAN * rows * NX, dt, "rows [%d,%d) of the %dx%dx%d cube (%d of %d rows), oracle %s, %d worker process(es)" % (
            y0, y0 + rows, NX, NY, NCHAN, rows, NY, what, n), n
    if config == 4:
        nchan, ny, nx = C4_SHAPE
        h, rows, planes = 14, int(96 * scale), threads
        y0 = ny // 2
        blk = _cached_block(planes, rows + 2 * h, nx, y0=y0 - h, ny_total=ny, nx_total=nx, seed=SEED, nan_permille=1, border=102)
        _CPU_SAMPLE = (blk, h)
        dt, n = _timed_pool_map(_cpu_c4_work, planes, threads)
        return planes * rows * nx, dt, ("%d channel planes x rows [%d,%d) of the %dx%dx%d cube (+14 halo rows each side), oracle "
                                        "convolve with the 29x29 Gaussian (841 taps per voxel, as astropy's direct 2-D "
                                        "convolution), %d worker process(es)" % (planes, y0, y0 + rows, nx, ny, nchan, n)), n
    if config == 5:
        nchan, ny, nx = TARGET_SHAPE
        nout = 1024
        rows = max(threads, int(16 * scale))
        y0 = ny // 2
        data = _cached_block(nchan, rows, nx, y0=y0, ny_total=ny, nx_total=nx, seed=SEED, nan_permille=1, border=102)
        inaxis = -321.214698632 + 1.28821496879 * np.arange(nchan)
        grid = np.linspace(inaxis[0], inaxis[-1], nout)
        bounds = np.linspace(0, rows, min(threads, rows) + 1).astype(int)
        _CPU_SAMPLE = (data, inaxis, grid, bounds)
        dt_i, n = _timed_pool_map(_cpu_c5_interp_work, len(bounds) - 1, threads)
        vox_i = nchan * rows * nx
        # reproject: `planes` whole 4096x4096 channel images through the 30 degree pixel map
        from oracle.wcs import OWCS
        from oracle import reproject as orep
        planes = threads
        img = _cached_block(2, ny, nx, seed=SEED + 1, nan_permille=1, border=102)       # the workers alternate between two images
        w_in = OWCS(**wcs_kw(ny, nx))
        a = np.radians(30.0)
        kw = wcs_kw(ny, nx)
        kw['pc'] = [[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]]
        w_out = OWCS(**kw)
        if 'pixmap' not in _SYNTH_CACHE:
            t0 = time.perf_counter()
            pm = orep.input_pixel_coords(w_in, w_out, (ny, nx))            # once per cube, like reproject_interp
            _SYNTH_CACHE['pixmap'] = pm + (time.perf_counter() - t0,)
        yin, xin, dt_map = _SYNTH_CACHE['pixmap']
        _CPU_SAMPLE = (img, yin, xin)
        dt_r, n2 = _timed_pool_map(_cpu_c5_reproject_work, planes, threads)
        vox_r = planes * ny * nx
        # per INPUT voxel of config 5: interpolation on all of them, reproject on the nout/nchan interpolated ones
        s_per_voxel = dt_i / vox_i + (dt_r / vox_r) * nout / nchan
        return vox_i, s_per_voxel * vox_i, (
            "spectral_interpolate 2048->1024 (numpy class, np.interp per spaxel) on rows [%d,%d) of the %dx%dx%d cube: %.2f s; "
            "reproject (bilinear, 30 deg) of %d interpolated 4096x4096 planes: %.2f s (+ %.2f s pixel map, once per cube, not "
            "counted); combined as seconds per input voxel; %d worker process(es)" % (y0, y0 + rows, nx, ny, nchan, dt_i, planes, dt_r, dt_map, n)), n
    raise SystemExit("--config must be 1, 3, 4 or 5")


CONFIG_TEXT = {
    1: 'configs[1]: %dx%dx%d float32 synthetic cube per GPU, moment0+moment1+moment2 (three drop-in calls) under a >3 sigma '
       'LazyMask & isfinite' % (NX, NY, NCHAN),
    3: 'configs[2]: 2048x2048x1024 cube, spectral_smooth Gaussian FWHM=5 chan then moment1',
    4: 'configs[3]: 4096x4096x512 cube, spatial_smooth Gaussian FWHM=8 px',
    5: 'configs[4]: 4096x4096x2048 cube, spectral_interpolate to 1024 channels + reproject to a WCS rotated by 30 deg',
}


def workload_config(n, config=1):
    if config != 1:
        return {'workload': CONFIG_TEXT[config]}
    return {'workload': CONFIG_TEXT[1], 'shape_per_gpu': [NCHAN, NY, NX], 'mask': 'isfinite & (cube > 3.0)',
            'sharding': 'rows (spatial plane), %d shard(s), no data-path collective' % n,
            'timed_call': 'SpectralCube._moments_axis0_raw(order bit) x 3 = the device part of moment0/1/2(); the public '
                          'calls incl. the copy of the maps to the host are what `e2e` times',
            'l2': 'inputs (17.2 GB per pass) far exceed the 126 MB L2; no explicit flush needed'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = host_threads()
    vals, desc, n = [], '', 1
    for i in range(args.warmup + args.steps):
        vox, dt, desc, n = cpu_sample(args.config, threads, scale=args.cpu_scale)
        if i >= args.warmup:
            vals.append((vox, dt))
    tot_v = sum(v for v, _ in vals)
    tot_t = sum(t for _, t in vals)
    value = tot_v / tot_t
    line = {
        'impl': 'reference', 'metric': 'voxels/sec', 'value': value, 'unit': 'voxels/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * tot_t / len(vals), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.gpus, args.config),
        'cpu_baseline': {'value': value, 'unit': 'voxels/s', 'cores': n, 'kind': 'port', 'sample': desc},
        'e2e': {'value': value, 'unit': 'voxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# ==== GPU side ============================================================================================
class Job(object):
    """torch.distributed plumbing + timing helpers shared by the sections."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device; the product has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.all_cpus = os.sched_getaffinity(0)
        self.numa = bind_to_gpu_numa_node(self.local_rank) if not args.no_numa else {'bound': False}
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group('nccl', device_id=torch.device('cuda', self.local_rank))
            self.dist = dist
        from spectral_cube_b200 import _lib
        self.lib = _lib.load()
        self.peak, self.peak_src = load_peaks()
        self.deferred = []

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.dist is None:
            return float(ms)
        t = self.torch.tensor([ms], dtype=self.torch.float64, device='cuda')
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timeit(self, f, n=5, warm=2):
        """Mean device time of f() over n calls (CUDA events on the launching stream), max over ranks."""
        torch = self.torch
        for _ in range(warm):
            f()
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            f()
        b.record()
        self.barrier()
        return self.max_over_ranks(a.elapsed_time(b) / n)

    def roof(self, algo_bytes_per_gpu, ms, kernel, **extra):
        ach = algo_bytes_per_gpu / (ms * 1e-3) / 1e9
        r = {'bound': 'hbm', 'achieved': ach, 'peak': self.peak, 'unit': 'GB/s', 'frac': ach / self.peak, 'traffic': None,
             'kernel': kernel, 'algorithmic_bytes_per_launch': algo_bytes_per_gpu, 'kernel_ms': ms, 'peak_source': self.peak_src}
        r.update(extra)
        return r

    def cpu_leg(self, config):
        """The CPU arm on ALL host cores (the NUMA binding of the GPU legs is lifted for it)."""
        bound = os.sched_getaffinity(0)
        os.sched_setaffinity(0, self.all_cpus)
        try:
            vox, dt, desc, n = cpu_sample(config, host_threads(), scale=self.args.cpu_scale)
        finally:
            os.sched_setaffinity(0, bound)
        return {'value': vox / dt, 'unit': 'voxels/s', 'cores': n, 'kind': 'port', 'sample': desc}

    def defer_cpu_leg(self, target, config, **extra):
        """CPU legs run AFTER every GPU section: forking sixteen workers out of a process that holds a CUDA context made the
        sections that followed slower (reproject: 26 ms instead of 11 ms per call), so nothing on the GPU is timed after a fork."""
        self.deferred.append((target, config, extra))

    def run_deferred_cpu_legs(self):
        done = {}
        for target, config, extra in self.deferred:
            if config not in done:
                done[config] = self.cpu_leg(config)
            target['cpu_baseline'] = dict(done[config], **extra)

    def free(self):
        import gc
        gc.collect()
        self.torch.cuda.empty_cache()


def isfinite_cube(cls, dev, w, **kw):
    import numpy as np
    import spectral_cube_b200 as scb
    c = cls(dev, w, unit='K', allow_huge_operations=True, **kw)
    c._mask = scb.LazyMask(np.isfinite, cube=c)      # what io/fits.py:214 attaches on read
    return c


def section_headline(job, line):
    """configs[1]: the three moments (value), fused pass, spectral_smooth, e2e, cpu_baseline."""
    import numpy as np
    import spectral_cube_b200 as scb
    from spectral_cube_b200 import _lib
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs
    torch, args, rank, world, lib = job.torch, job.args, job.rank, job.world, job.lib

    ny_total = NY * world
    dev = synth_cube(NCHAN, NY, NX, y0=rank * NY, ny_total=ny_total, nx_total=NX, seed=SEED, nan_permille=1, border=BORDER)
    wcs = benchmark_wcs(NCHAN, ny_total, NX)
    wcs.crpix[1] -= rank * NY

    def make_cube(data):
        c = isfinite_cube(scb.SpectralCube, data, wcs)
        return c.with_mask(c > THRESHOLD)

    cube = make_cube(dev)
    voxels = NCHAN * NY * NX

    def step_device():
        cube._moments_axis0_raw(1)
        cube._moments_axis0_raw(2)
        cube._moments_axis0_raw(4)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    job.barrier()
    # clocks: nvidia-smi needs a few hundred ms to start, far longer than the timed region, so it is started here
    # and the GPU is kept under the SAME load (untimed pre-roll steps) until it reports
    sampler = ClockSampler(job.local_rank)
    if rank == 0:
        sampler.start()
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 1.0:
        step_device()
        torch.cuda.synchronize()
    job.barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    n0 = lib.sc_launch_count()
    job.barrier()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin.record()
    for k in range(args.steps):
        ev[k][0].record()
        cube._moments_axis0_raw(1)
        ev[k][1].record()
        cube._moments_axis0_raw(2)
        ev[k][2].record()
        cube._moments_axis0_raw(4)
        ev[k][3].record()
    t_end.record()
    job.barrier()
    launches = lib.sc_launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = job.max_over_ranks(t_begin.elapsed_time(t_end)) / args.steps
    per_call = [statistics.mean(ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(args.steps)) for i in range(3)]
    fused_ms = job.timeit(lambda: cube._moments_axis0_raw(7), n=max(3, args.steps // 2), warm=2)

    spaxels = NY * NX
    algo = 4 * voxels + 8 * spaxels                          # one moment pass
    traffic, tfile = ncu_traffic('r02_moments_c2_ncu_full.txt', 'moments_tma_kernel')
    if traffic is None:
        traffic, tfile = ncu_traffic('r01_moments_c2_ncu_full.txt', 'moments_tma_kernel')
    line.update({
        'metric': 'voxels/sec', 'value': world * voxels / (ms_per_step * 1e-3), 'unit': 'voxels/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': warm, 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(world),
        'clocks': clocks, 'gpu_launches': int(launches),
        'per_call_ms': {'moment0': per_call[0], 'moment1': per_call[1], 'moment2': per_call[2],
                        'fused_moment012_one_pass': fused_ms},
        'fused_voxels_per_s': world * voxels / (fused_ms * 1e-3),
        'roofline': job.roof(algo, per_call[2], 'moments_tma_kernel<8,4,INTERVAL,M2> (the moment2 call)',
                             traffic=traffic, traffic_source=tfile),
        'numa': job.numa,
    })

    # ---- the other half of the north-star target on the same resident cube: spectral_smooth ----
    if not args.no_smooth:
        try:
            iso = isfinite_cube(scb.DaskSpectralCube, dev, wcs)
            k17 = scb.Gaussian1DKernel(5 * FWHM2SIGMA).array
            sm_out = torch.empty_like(dev)
            ms = job.timeit(lambda: iso._run_spectral_smooth(k17, _lib.F32, out=sm_out), n=5, warm=2)
            tr, tf = ncu_traffic('r02_spectral_smooth_kernel_ncu_full.txt', 'smooth_tma_kernel')
            if tr is None:
                tr, tf = ncu_traffic('r01_spectral_smooth_ncu_full_v11.txt', 'smooth_tma_kernel')
            line['spectral_smooth'] = {'ms': ms, 'voxels_per_s': world * voxels / (ms * 1e-3),
                                       'roofline': job.roof(8 * voxels, ms, 'smooth_tma_kernel<8,INTERVAL,f32> (17 taps)',
                                                            traffic=tr, traffic_source=tf)}
            del sm_out, iso
        except Exception as exc:
            line['spectral_smooth'] = {'error': repr(exc)[:300]}
        job.free()

    # ---- end to end: pinned host cube -> upload -> three moments through the public API -> host maps ----
    if not args.no_e2e:
        import warnings
        host = torch.empty((NCHAN, NY, NX), dtype=torch.float32, pin_memory=True)   # 17.2 GB per rank, NUMA-local
        host.copy_(dev)
        torch.cuda.synchronize()

        def step_e2e():
            c = make_cube(host.cuda(non_blocking=True))
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                return [c.moment0().value, c.moment1().value, c.moment2().value]

        step_e2e()
        job.barrier()
        nstep = max(1, min(args.steps, args.e2e_steps))
        e0 = time.perf_counter()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(nstep):
            maps = step_e2e()
        b1.record()
        torch.cuda.synchronize()
        wall_ms = 1e3 * (time.perf_counter() - e0) / nstep
        job.barrier()
        # results land on the host synchronously: use the slower of the two clocks, max over ranks
        e2e_ms = job.max_over_ranks(max(b0.elapsed_time(b1) / nstep, wall_ms))
        line['e2e'] = {'value': world * voxels / (e2e_ms * 1e-3), 'unit': 'voxels/s',
                       'h2d_bytes_per_step': voxels * 4, 'd2h_bytes_per_step': 3 * NY * NX * 8,
                       'ms_per_step': e2e_ms, 'steps': nstep, 'h2d_GBps_per_gpu': voxels * 4 / e2e_ms / 1e6,
                       'path': 'SpectralCube(pinned host tensor) -> with_mask -> moment0/1/2().value'}
        del host, maps
    del cube, dev
    job.free()

    if world == 1 and rank == 0 and not args.no_cpu:
        job.defer_cpu_leg(line, 1)


def section_c3(job, line):
    """configs[2]: spectral_smooth (17 taps) -> moment1 on the 2048x2048x1024 cube, one rank's cube each (weak)."""
    import spectral_cube_b200 as scb
    from spectral_cube_b200 import _lib
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs
    world, rank = job.world, job.rank
    V = NCHAN * NY * NX
    dev = synth_cube(NCHAN, NY, NX, y0=rank * NY, ny_total=NY * world, nx_total=NX, seed=SEED, nan_permille=1, border=BORDER)
    w = benchmark_wcs(NCHAN, NY * world, NX)
    k = scb.Gaussian1DKernel(5 * FWHM2SIGMA)
    c = isfinite_cube(scb.DaskSpectralCube, dev, w)
    sm = c.spectral_smooth(k)                                       # lazy: the moment runs the fused kernel
    ms_f = job.timeit(lambda: sm._moments_axis0_raw(2), n=5, warm=2)
    out = {'workload': CONFIG_TEXT[3], 'scaling': 'weak',
           'fused': {'ms': ms_f, 'voxels_per_s': world * V / (ms_f * 1e-3),
                     'what': 'DaskSpectralCube.spectral_smooth(k).moment1(): the smoothed cube is never written',
                     'roofline': job.roof(4 * V + 8 * NY * NX, ms_f, 'smooth_tma_kernel<8,INTERVAL,moment epilogue>')}}
    cn = isfinite_cube(scb.SpectralCube, dev, w)
    buf = job.torch.empty_like(dev)
    ms_s = job.timeit(lambda: cn._run_spectral_smooth(k.array, _lib.F32, out=buf), n=5, warm=2)
    mat = cn._new_cube_reporting_f64(buf)                           # what SpectralCube.spectral_smooth returns
    ms_m = job.timeit(lambda: mat._moments_axis0_raw(2), n=5, warm=2)
    out['unfused'] = {'ms': ms_s + ms_m, 'smooth_ms': ms_s, 'moment1_ms': ms_m, 'voxels_per_s': world * V / ((ms_s + ms_m) * 1e-3),
                      'what': 'SpectralCube (numpy class): smooth materialised (8 B/voxel), then moment1 under the mask of the '
                              'SOURCE cube (reads both cubes, 8 B/voxel)',
                      'roofline': job.roof(16 * V + 8 * NY * NX, ms_s + ms_m, 'smooth_tma_kernel + moments_axis0_kernel<INTERVAL_OTHER>')}
    line['c3'] = out
    del dev, c, sm, cn, buf, mat
    job.free()
    if world == 1 and rank == 0 and not job.args.no_cpu:
        job.defer_cpu_leg(out, 3)
        if isinstance(line.get('spectral_smooth'), dict) and 'ms' in line['spectral_smooth']:
            job.defer_cpu_leg(line['spectral_smooth'], 3, note='configs[2] sample: smooth + moment1; the smooth is > 90 % of the CPU time')


def section_c4_strong(job, line):
    """configs[3], strong scaling: the fixed 4096x4096x512 cube row-sharded over the ranks; every timed call packs
    its edge rows, swaps them with the neighbours over NCCL and convolves (RowShardedCube.spatial_smooth)."""
    import numpy as np
    import spectral_cube_b200 as scb
    from spectral_cube_b200 import distributed as D
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs
    world, rank = job.world, job.rank
    nchan, ny, nx = C4_SHAPE
    V = nchan * ny * nx
    y0, y1 = D.row_partition(ny, world)[rank]
    dev = synth_cube(nchan, y1 - y0, nx, y0=y0, ny_total=ny, nx_total=nx, seed=SEED, nan_permille=1, border=102)
    w = benchmark_wcs(nchan, ny, nx)
    k = scb.Gaussian2DKernel(8 * FWHM2SIGMA)
    out = {'workload': CONFIG_TEXT[4] + ', row-sharded over %d GPU(s), halo exchange inside the timed region' % world,
           'scaling': 'strong', 'shape': list(C4_SHAPE), 'rows_per_gpu': y1 - y0,
           'halo_bytes_per_neighbour': 14 * nx * nchan * 4 if world > 1 else 0}
    if world == 1:
        sh = D.RowShardedCube(isfinite_cube(scb.DaskSpectralCube, dev, w), ny, 0, None)
        modes = ('none',)
    else:
        sh = D.RowShardedCube.from_full_wcs(scb.DaskSpectralCube, dev, w, ny, unit='K', allow_huge_operations=True)
        sh.local._mask = scb.LazyMask(np.isfinite, cube=sh.local)
        modes = ('p2p', 'allgather')
    n0 = job.lib.sc_launch_count()
    for mode in modes:
        ms = job.timeit(lambda: sh.spatial_smooth(k, halo_mode='p2p' if mode == 'none' else mode), n=3, warm=1)
        out['halo_' + mode] = {'ms': ms, 'voxels_per_s': V / (ms * 1e-3),
                               'roofline': job.roof(8 * V // world, ms, 'sep_pipe_kernel<14> + sep_fixup_kernel (29x29 separable; sep_march_kernel when the sample says crowded)',
                                                    note='FP64-pipe bound: 58 DFMA per voxel; floor at 15.8 TDFMA/s = %.1f ms per GPU'
                                                         % (58.0 * V / world / 15.8e12 * 1e3))}
    out['gpu_launches'] = int(job.lib.sc_launch_count() - n0)
    best = min(out['halo_' + m]['ms'] for m in modes)
    out['ms'] = best
    out['value'] = V / (best * 1e-3)
    out['unit'] = 'voxels/s'
    line['c4_strong'] = out
    del dev, sh
    job.free()
    if world == 1 and rank == 0 and not job.args.no_cpu:
        job.defer_cpu_leg(out, 4)


def section_c5(job, line):
    """configs[4]: spectral_interpolate 2048 -> 1024 channels, then reproject to a WCS rotated by 30 degrees."""
    import warnings
    import numpy as np
    import spectral_cube_b200 as scb
    from spectral_cube_b200 import distributed as D
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs
    world, rank, torch = job.world, job.rank, job.torch
    nchan, ny, nx = TARGET_SHAPE
    nout = 1024
    V = nchan * ny * nx
    shards = world if world >= 4 else 8        # fewer than 4 GPUs cannot hold cube + outputs: time the pieces of an 8-way job
    y0, y1 = D.row_partition(ny, shards)[rank]
    dev = synth_cube(nchan, y1 - y0, nx, y0=y0, ny_total=ny, nx_total=nx, seed=SEED, nan_permille=1, border=102)
    w = benchmark_wcs(nchan, ny, nx)
    wl = w.copy()
    wl.crpix[1] -= y0
    c = isfinite_cube(scb.SpectralCube, dev, wl)
    sa = c.spectral_axis
    grid = np.linspace(sa[0], sa[-1], nout)
    out = {'workload': CONFIG_TEXT[5], 'scaling': 'strong' if world >= 4 else 'one of 8 shards per GPU'}
    a = np.radians(30.0)
    hdr = dict(w.to_header())
    hdr.update({'NAXIS': 3, 'NAXIS1': nx, 'NAXIS2': ny, 'NAXIS3': nout, 'PC1_1': np.cos(a), 'PC1_2': -np.sin(a),
                'PC2_1': np.sin(a), 'PC2_2': np.cos(a)})
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ms_i = job.timeit(lambda: c.spectral_interpolate(grid), n=3, warm=1)
        Vs = nchan * (y1 - y0) * nx
        out['spectral_interpolate'] = {'ms': ms_i, 'rows_per_gpu': y1 - y0, 'voxels_per_s': Vs * min(world, shards) / (ms_i * 1e-3),
                                       'roofline': job.roof(4 * Vs + 5 * Vs // 2, ms_i, 'spectral_interp_tma_kernel (+1 B/out-voxel mask)')}
        ms_job = None
        if world >= 4:
            # the whole job with the interpolation kernel scattering its output rows to the channel owners over NVLink
            # peer memory (its stores are the re-shard), then the local reproject of each rank's channels
            shin = D.RowShardedCube(c, ny, y0, None)
            try:
                ms_f = job.timeit(lambda: shin.spectral_interpolate_to_channels(grid, mode='peer'), n=3, warm=1)
                sent_f = int(nout * (y1 - y0) * nx * 4 * (world - 1) // world)
                out['interpolate_scatter'] = {'ms': ms_f, 'bytes_sent_per_gpu': sent_f, 'GBps_per_gpu_per_direction': sent_f / ms_f / 1e6,
                                              'nvlink_peak_GBps_per_direction': 770.0,
                                              'what': 'sc_spectral_interp_scatter: spectral_interp_tma_kernel storing every output channel '
                                                      'into its owner\'s buffer (peer memory), barriers included'}
                ms_job = job.timeit(lambda: shin.spectral_interpolate_reproject(grid, hdr, mode='peer'), n=3, warm=2)
                out['fused_job'] = {'ms': ms_job, 'what': 'RowShardedCube.spectral_interpolate_reproject: interpolate+scatter, pixel map, bilinear'}
            except Exception as exc:
                out['interpolate_scatter'] = {'unavailable': repr(exc)[:300]}
            del shin
            interp = c.spectral_interpolate(grid)
            del dev, c
            job.free()
            sh = D.RowShardedCube(interp, ny, y0, None)
            sent = int(interp._data.numel() * 4 * (world - 1) // world)
            ms_x = job.timeit(lambda: D.reshard_rows_to_channels(interp._data, ny, mode='nccl'), n=3, warm=1)
            out['reshard_all_to_all'] = {'ms': ms_x, 'bytes_sent_per_gpu': sent, 'GBps_per_gpu_per_direction': sent / ms_x / 1e6,
                                         'what': 'NCCL all_to_all into per-source blocks + concatenation'}
            mode = 'nccl'
            try:
                ms_p = job.timeit(lambda: D.reshard_rows_to_channels(interp._data, ny, mode='peer', borrow=True), n=3, warm=1)
                out['reshard_peer'] = {'ms': ms_p, 'bytes_sent_per_gpu': sent, 'GBps_per_gpu_per_direction': sent / ms_p / 1e6,
                                       'what': 'sc_reshard_scatter: one kernel per rank storing into the destination ranks\' final '
                                               'layout over NVLink peer memory (symmetric buffers), barriers included',
                                       'nvlink_peak_GBps_per_direction': 770.0}
                mode = 'peer'
            except Exception as exc:
                out['reshard_peer'] = {'unavailable': repr(exc)[:300]}
            ms_r = job.timeit(lambda: sh.reproject(hdr, reshard_mode=mode), n=3, warm=2)
            out['reproject_sharded'] = {'ms': ms_r, 'reshard': mode, 'what': 'rows->channels re-shard + pixel map + bilinear'}
            out['two_step_ms'] = ms_i + ms_r
            out['ms'] = min(ms_i + ms_r, ms_job) if ms_job else ms_i + ms_r
            out['value'] = V / (out['ms'] * 1e-3)
            out['unit'] = 'voxels/s'
            del interp, sh
        else:
            del dev, c
            job.free()
            nloc = nout // shards                      # the channel shard a rank holds after the re-shard: whole planes
            planes = synth_cube(nloc, ny, nx, seed=SEED + 1, nan_permille=1, border=102)
            cc = isfinite_cube(scb.SpectralCube, planes, w)
            hdr['NAXIS3'] = nloc
            ms_r = job.timeit(lambda: cc.reproject(hdr), n=5, warm=3)       # (the call allocates its 17 GB result: three warm-up calls fill the allocator's cache)
            Vr = nloc * ny * nx
            # the two device parts on their own (diagnostic: the public call adds header parsing and allocation on the host)
            from spectral_cube_b200.wcs import as_cube_wcs
            neww = as_cube_wcs(hdr)
            ms_map = job.timeit(lambda: cc._pixel_map(neww, ny, nx), n=3, warm=1)
            yin, xin = cc._pixel_map(neww, ny, nx)
            ms_kern = job.timeit(lambda: cc._run_reproject(yin, xin, 1), n=3, warm=2)
            t0 = time.perf_counter()
            for _ in range(3):
                cc.reproject(hdr)
            job.torch.cuda.synchronize()
            out['reproject_parts'] = {'pixel_map_ms': ms_map, 'bilinear_kernel_ms': ms_kern,
                                      'public_call_wall_ms': (time.perf_counter() - t0) / 3 * 1e3}
            del yin, xin
            out['reproject'] = {'ms': ms_r, 'planes_per_gpu': nloc, 'voxels_per_s': Vr * world / (ms_r * 1e-3),
                                'roofline': job.roof(4 * Vr + 9 * Vr + 16 * ny * nx, ms_r,
                                                     'wcs_pixel_map_kernel + reproject_tiled_kernel (f64 + footprint out; the float32 '
                                                     'working copy is made lazily)',
                                                     parity='oracle unpinned (oracle/reproject.py)')}
            del planes, cc
    line['c5'] = out
    job.free()
    if world == 1 and rank == 0 and not job.args.no_cpu:
        job.defer_cpu_leg(out, 5, unit='voxels/s (input voxels of config 5)')


def section_target_strong(job, line):
    """North-star target: moment0/1/2 + spectral_smooth on the 4096x4096x2048 cube (137.4 GB); N ranks hold 1/N of
    the rows each (no collective on this path); one GPU holds all of it and smooths in place."""
    import spectral_cube_b200 as scb
    from spectral_cube_b200 import _lib, distributed as D
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs
    world, rank = job.world, job.rank
    nchan, ny, nx = TARGET_SHAPE
    y0, y1 = D.row_partition(ny, world)[rank]
    V, S = nchan * ny * nx, ny * nx
    Vg, Sg = nchan * (y1 - y0) * nx, (y1 - y0) * nx
    dev = synth_cube(nchan, y1 - y0, nx, y0=y0, ny_total=ny, nx_total=nx, seed=SEED, nan_permille=1, border=102)
    w = benchmark_wcs(nchan, ny, nx)
    w.crpix[1] -= y0
    c = isfinite_cube(scb.DaskSpectralCube, dev, w)
    c = c.with_mask(c > THRESHOLD)
    out = {'workload': 'north-star target: 4096x4096x2048 float32 cube (137.4 GB), moment0+1+2 under isfinite & >3 sigma, and '
                       'spectral_smooth FWHM 5 ch (17 taps) float32 in place; rows split over %d GPU(s)' % world,
           'scaling': 'strong', 'shape': list(TARGET_SHAPE), 'rows_per_gpu': y1 - y0, 'bytes_per_gpu': 4 * Vg}

    def three():
        c._moments_axis0_raw(1)
        c._moments_axis0_raw(2)
        c._moments_axis0_raw(4)
    ms3 = job.timeit(three, n=3, warm=1)
    ms_f = job.timeit(lambda: c._moments_axis0_raw(7), n=3, warm=1)
    out['moments'] = {'ms': ms3, 'voxels_per_s': V / (ms3 * 1e-3), 'fused_one_pass_ms': ms_f,
                      'roofline': job.roof(3 * (4 * Vg + 8 * Sg), ms3, 'moments_tma_kernel x 3 (moment0, moment1, moment2)')}
    k = scb.Gaussian1DKernel(5 * FWHM2SIGMA)
    c2 = isfinite_cube(scb.DaskSpectralCube, dev, w)
    ms_s = job.timeit(lambda: c2._run_spectral_smooth(k.array, _lib.F32, out=dev), n=3, warm=1)
    out['spectral_smooth'] = {'ms': ms_s, 'voxels_per_s': V / (ms_s * 1e-3),
                              'roofline': job.roof(8 * Vg, ms_s, 'smooth_tma_kernel<8,INTERVAL,f32> in place')}
    tot = ms3 + ms_s
    out['ms'] = tot
    out['value'] = V / (tot * 1e-3)
    out['unit'] = 'voxels/s'
    out['roofline'] = job.roof(3 * (4 * Vg + 8 * Sg) + 8 * Vg, tot, 'moment0/1/2 + spectral_smooth')
    line['target_strong'] = out
    del dev, c, c2
    job.free()


def section_selftest(job, line):
    from spectral_cube_b200.selftest import sharded_parity, all_ranks_agree
    res = sharded_parity()
    ok, bad = all_ranks_agree(res)
    line['selftest'] = {'sharded_equals_single_gpu_bit_for_bit': ok, 'checks': sorted(n for n in res if not n.startswith('_')),
                        'failed': bad, 'ranks': job.world, 'detail_rank0': res.get('_detail', {})}


def run_ours(args):
    job = Job(args)
    line = {}
    section_headline(job, line)
    line['hbm_held_after_GB'] = {'headline': round(job.torch.cuda.memory_allocated() / 1e9, 2)}   # tensors still referenced
    sections = [('c3', section_c3, args.no_c3 or job.world > 1), ('c4_strong', section_c4_strong, args.no_c4),
                ('c5', section_c5, args.no_c5 or job.world in (2,)), ('target_strong', section_target_strong, args.no_target),
                ('selftest', section_selftest, args.no_selftest or job.world < 2)]
    for name, fn, skip in sections:
        if skip:
            continue
        try:
            t0 = time.perf_counter()
            fn(job, line)
            if isinstance(line.get(name), dict):
                line[name]['section_wall_s'] = round(time.perf_counter() - t0, 2)
            line['hbm_held_after_GB'][name] = round(job.torch.cuda.memory_allocated() / 1e9, 2)
        except Exception as exc:                      # never lose the headline to an extra
            line[name] = {'error': repr(exc)[:400]}
            try:
                job.free()
            except Exception:
                pass
    if job.world == 1 and not args.no_cpu:
        job.run_deferred_cpu_legs()
    if job.rank == 0:
        print(json.dumps(line))
    if job.dist is not None:
        job.dist.destroy_process_group()


def run_config0(args):
    """BASELINE.json configs[0], the reference's own CPU-runnable case: 256x256x128, moment0, no mask.  The CPU
    leg is the oracle port of `moment_cubewise` (what `moment_auto` picks below 1e8 voxels) on one thread; the
    GPU leg the same call through the product.  One JSON line."""
    import warnings
    import numpy as np
    import torch
    import spectral_cube_b200 as scb
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs
    from oracle.cube import OracleCube
    from oracle.wcs import OWCS
    nchan, ny, nx = 128, 256, 256
    V = nchan * ny * nx
    torch.cuda.set_device(0)
    dev = synth_cube(nchan, ny, nx, nan_permille=0, border=0)
    c = scb.SpectralCube(dev, benchmark_wcs(nchan, ny, nx), unit='K')
    for _ in range(5):
        c._moments_axis0_raw(1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        c._moments_axis0_raw(1)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    oc = OracleCube(dev.cpu().numpy(), OWCS(**wcs_kw(ny, nx)), unit='K', mask=None)
    t0 = time.perf_counter()
    for _ in range(5):
        ref = oc.moment(order=0, how='auto')[0]
    cpu_s = (time.perf_counter() - t0) / 5
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ok = bool(np.allclose(c.moment0().value, ref, rtol=1e-5, equal_nan=True))
    print(json.dumps({'metric': 'voxels/sec', 'value': V / (ms * 1e-3), 'unit': 'voxels/s', 'n_gpus': 1, 'steps': args.steps,
                      'ms_per_step': ms, 'config': {'workload': 'configs[0]: 256x256x128 float32 synthetic cube, moment0, no mask (33.6 MB: L2 resident)'},
                      'parity_rtol_1e5': ok,
                      'cpu_baseline': {'value': V / cpu_s, 'unit': 'voxels/s', 'cores': 1, 'kind': 'port',
                                       'sample': 'the whole cube, oracle moment_cubewise, 1 thread'}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=1, choices=[1, 3, 4, 5],
                    help='--impl reference: which BASELINE.json config the CPU arm samples (1 = configs[1], the headline)')
    ap.add_argument('--cpu-scale', type=float, default=1.0, help='scale of the bounded CPU samples')
    ap.add_argument('--e2e-steps', type=int, default=3)
    for flag in ('e2e', 'cpu', 'smooth', 'c3', 'c4', 'c5', 'target', 'selftest', 'numa'):
        ap.add_argument('--no-' + flag, action='store_true')
    ap.add_argument('--only-headline', action='store_true', help='skip every extra section')
    ap.add_argument('--config0', action='store_true', help="BASELINE configs[0] (the reference's CPU-runnable case) instead of configs[1]")
    args = ap.parse_args()
    if args.only_headline:
        args.no_c3 = args.no_c4 = args.no_c5 = args.no_target = args.no_selftest = True
    if args.config0:
        run_config0(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
