#!/usr/bin/env python
"""
bench.py -- throughput of the moment-map hot path (BASELINE.json metric: voxels/sec).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (N=1): BASELINE.json configs[1] -- a 2048x2048x1024 float32 synthetic cube, moment0,
moment1 and moment2 under a >3 sigma LazyMask on top of the isfinite mask.  A step = the three
drop-in calls ``moment0()``, ``moment1()``, ``moment2()`` (three passes over the cube in this
implementation, four in the reference).  ``value`` = input voxels of the cube / time of a step
with the cube resident in HBM; ``e2e`` = the same step starting from a pinned HOST cube
(upload + three moments + maps back on the host).  N>1: every rank owns its own
2048x2048x1024 row block of a taller cube (weak scaling, no data-path collective).

``--impl reference`` times the CPU restatement of the reference algorithm (oracle/, slicewise
strategy as ``moment_auto`` picks for >= 1e8 voxels) with a thread pool over row blocks on all
host cores, on a bounded row-block sample of the same cube.
"""
import argparse
import json
import os
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


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


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


# ---- CPU side (oracle port of the reference algorithm) -----------------------------------------
def cpu_moments_sample(rows, threads):
    """Time moment0/1/2 (slicewise, as moment_auto picks for >= 1e8 voxels) of a `rows`-row block of
    the benchmark cube on `threads` host threads.  Returns (voxels, seconds, description)."""
    import warnings
    import numpy as np
    from oracle.synth import synth_block
    from oracle.cube import OracleCube
    from oracle.wcs import OWCS

    y0 = NY // 2 - rows // 2
    data = synth_block(NCHAN, rows, NX, y0=y0, ny_total=NY, nx_total=NX, seed=SEED, nan_permille=1, border=BORDER)
    wkw = dict(ctype=['RA---TAN', 'DEC--TAN', 'VRAD'], crval=[24.0, 30.0, -321.214698632],
               crpix=[NX / 2.0 + 0.5, NY / 2.0 + 0.5, 1.0],
               cdelt=[-5.55555561268e-4, 5.55555561268e-4, 1.28821496879], cunit=['deg', 'deg', 'km/s'])
    nblk = max(1, min(threads, rows))
    bounds = np.linspace(0, rows, nblk + 1).astype(int)
    global _CPU_SAMPLE
    _CPU_SAMPLE = (data, wkw, bounds)

    t0 = time.perf_counter()
    if nblk == 1:
        _cpu_work(0)
    else:
        # worker PROCESSES (fork; the sample is inherited copy-on-write): numpy's masked-array
        # path holds the GIL, so threads do not scale -- this is dask's 'processes' scheduler
        import multiprocessing as mp
        with mp.get_context('fork').Pool(nblk) as pool:
            pool.map(_cpu_work, range(nblk))
    dt = time.perf_counter() - t0
    desc = "rows [%d,%d) of the %dx%dx%d cube (%d of %d rows), oracle slicewise moment0+1+2, %d worker process(es)" % (
        y0, y0 + rows, NX, NY, NCHAN, rows, NY, nblk)
    return NCHAN * rows * NX, dt, desc


_CPU_SAMPLE = None


def _cpu_work(i):
    import warnings
    from oracle.cube import OracleCube
    from oracle.wcs import OWCS
    data, wkw, bounds = _CPU_SAMPLE
    blk = data[:, bounds[i]:bounds[i + 1], :]
    cube = OracleCube(blk, OWCS(**wkw), unit='K')
    cube = cube.with_mask(cube > THRESHOLD)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return [cube.moment(order=o, how='slice')[0] for o in (0, 1, 2)]


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = host_threads()
    rows = args.cpu_rows
    vals = []
    desc = ''
    for i in range(args.warmup + args.steps):
        vox, dt, desc = cpu_moments_sample(rows, threads)
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
        'config': workload_config(args.gpus),
        'cpu_baseline': {'value': value, 'unit': 'voxels/s', 'cores': threads, 'kind': 'port', 'sample': desc},
        'e2e': {'value': value, 'unit': 'voxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def workload_config(n):
    return {'workload': 'configs[1]: %dx%dx%d float32 synthetic cube per GPU, moment0+moment1+moment2 '
                        '(three drop-in calls) under a >3 sigma LazyMask & isfinite' % (NX, NY, NCHAN),
            'shape_per_gpu': [NCHAN, NY, NX], 'mask': 'isfinite & (cube > 3.0)',
            'sharding': 'rows (spatial plane), %d shard(s), no data-path collective' % n,
            'l2': 'inputs (17.2 GB per pass) far exceed the 126 MB L2; no explicit flush needed'}


# ---- GPU side -------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import spectral_cube_b200 as scb
    from spectral_cube_b200 import _lib
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    lib = _lib.load()

    ny_total = NY * world
    dev = synth_cube(NCHAN, NY, NX, y0=rank * NY, ny_total=ny_total, nx_total=NX, seed=SEED,
                     nan_permille=1, border=BORDER)
    wcs = benchmark_wcs(NCHAN, ny_total, NX)

    def make_cube(data):
        c = scb.SpectralCube(data, wcs, unit='K')
        c._mask = scb.LazyMask(np.isfinite, cube=c)          # what io/fits.py:214 attaches on read
        return c.with_mask(c > THRESHOLD)

    cube = make_cube(dev)
    voxels = NCHAN * NY * NX

    def step_device():
        # the three drop-in reductions, results left on the device
        cube._moments_axis0_raw(1)
        cube._moments_axis0_raw(2)
        cube._moments_axis0_raw(4)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    # clocks: nvidia-smi needs a few hundred ms to start, far longer than the timed region, so it is
    # started here and the GPU is kept under the SAME load (untimed pre-roll steps) until it reports;
    # the samples cover the pre-roll and the timed steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 1.0:
        step_device()
        torch.cuda.synchronize()
    barrier()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    n0 = lib.sc_launch_count()
    barrier()
    t_begin = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
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
    barrier()
    launches = lib.sc_launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_begin.elapsed_time(t_end)
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    per_call = [statistics.mean(ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(args.steps)) for i in range(3)]

    # one fused pass for all three maps (extension; not the headline)
    for _ in range(2):
        cube._moments_axis0_raw(7)
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        cube._moments_axis0_raw(7)
    f1.record()
    torch.cuda.synchronize()
    fused_ms = f0.elapsed_time(f1) / args.steps

    # ---- the other half of the north-star target on the same resident cube: spectral_smooth
    #      (Gaussian FWHM 5 channels = 17 taps, float32 out, 8 B/voxel) ----
    smooth_ms = None
    if not args.no_smooth:
        iso = scb.DaskSpectralCube(dev, wcs, unit='K')
        iso._mask = scb.LazyMask(np.isfinite, cube=iso)
        k17 = scb.Gaussian1DKernel(5 / 2.3548200450309493).array
        sm_out = torch.empty_like(dev)
        for _ in range(2):
            iso._run_spectral_smooth(k17, _lib.F32, out=sm_out)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(5):
            iso._run_spectral_smooth(k17, _lib.F32, out=sm_out)
        s1.record()
        torch.cuda.synchronize()
        smooth_ms = s0.elapsed_time(s1) / 5
        del sm_out, iso
        torch.cuda.empty_cache()

    # ---- end to end: pinned host cube -> upload -> three moments -> host maps ----
    e2e = None
    if not args.no_e2e:
        host = torch.empty((NCHAN, NY, NX), dtype=torch.float32, pin_memory=True)   # 17.2 GB per rank
        host.copy_(dev)
        torch.cuda.synchronize()

        def step_e2e():
            c = make_cube(host.cuda(non_blocking=True))
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                maps = [c.moment0().value, c.moment1().value, c.moment2().value]
            return maps

        for _ in range(1):
            step_e2e()
        barrier()
        e0 = time.perf_counter()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        nstep = max(1, min(args.steps, args.e2e_steps))
        for _ in range(nstep):
            maps = step_e2e()
        b1.record()
        barrier()
        e2e_ms = b0.elapsed_time(b1) / nstep
        wall_ms = 1e3 * (time.perf_counter() - e0) / nstep
        e2e_ms = max(e2e_ms, wall_ms)          # results land on the host synchronously: use the slower clock
        if dist is not None:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e = {'value': world * voxels / (e2e_ms * 1e-3), 'unit': 'voxels/s',
               'h2d_bytes_per_step': voxels * 4, 'd2h_bytes_per_step': 3 * NY * NX * 8,
               'ms_per_step': e2e_ms, 'steps': nstep,
               'path': 'SpectralCube(pinned host tensor) -> with_mask -> moment0/1/2().value'}
        del host

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    spaxels = NY * NX
    algo_bytes = 4 * voxels + 8 * spaxels                    # one moment pass
    k_ms = per_call[2]                                       # moment2 call = the dominant launch
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    line = {
        'metric': 'voxels/sec', 'value': world * voxels / (ms_per_step * 1e-3), 'unit': 'voxels/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(world),
        'clocks': clocks, 'gpu_launches': int(launches),
        'per_call_ms': {'moment0': per_call[0], 'moment1': per_call[1], 'moment2': per_call[2],
                        'fused_moment012_one_pass': fused_ms},
        'fused_voxels_per_s': world * voxels / (fused_ms * 1e-3),
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one launch, from the
                     # ncu --set full capture summarised in profiles/r01_moments_c2_ncu_full.txt
                     'traffic': 17179959000 + 15093504, 'kernel': 'moments_tma_kernel<8,4,INTERVAL,M0|M1|M2> (moment2 call)',
                     'algorithmic_bytes_per_launch': algo_bytes, 'kernel_ms': k_ms, 'peak_source': peak_src},
        'e2e': e2e,
    }
    if smooth_ms is not None:
        sa = 8 * voxels / (smooth_ms * 1e-3) / 1e9
        line['spectral_smooth'] = {'ms': smooth_ms, 'voxels_per_s': world * voxels / (smooth_ms * 1e-3),
                                   'roofline': {'bound': 'hbm', 'achieved': sa, 'peak': peak, 'unit': 'GB/s', 'frac': sa / peak,
                                                'traffic': 17322100000 + 17216667000,
                                                'kernel': 'smooth_tma_kernel<8,INTERVAL,f32> (17 taps)',
                                                'algorithmic_bytes_per_launch': 8 * voxels,
                                                'ncu': 'profiles/r01_spectral_smooth_ncu_full_v11.txt'}}
    if world == 1 and not args.no_cpu:
        vox, dt, desc = cpu_moments_sample(args.cpu_rows, host_threads())
        line['cpu_baseline'] = {'value': vox / dt, 'unit': 'voxels/s', 'cores': host_threads(), 'kind': 'port',
                                'sample': desc}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


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
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    wkw = dict(ctype=['RA---TAN', 'DEC--TAN', 'VRAD'], crval=[24.0, 30.0, -321.214698632], crpix=[nx / 2 + 0.5, ny / 2 + 0.5, 1.0],
               cdelt=[-5.55555561268e-4, 5.55555561268e-4, 1.28821496879], cunit=['deg', 'deg', 'km/s'])
    oc = OracleCube(dev.cpu().numpy(), OWCS(**wkw), unit='K', mask=None)
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
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cpu-rows', type=int, default=128, help='rows of the cube in the bounded CPU sample')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-smooth', action='store_true')
    ap.add_argument('--config0', action='store_true', help="BASELINE configs[0] (the reference's CPU-runnable case) instead of configs[1]")
    args = ap.parse_args()
    if args.config0:
        run_config0(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
