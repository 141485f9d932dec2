"""Time the smoothing kernels at benchmark shapes (GPU box only; scratch tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectral_cube_b200 as scb
from spectral_cube_b200 import _lib
from spectral_cube_b200.synth import synth_cube, benchmark_wcs

def timeit(f, n=5, warm=2):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

which = (sys.argv[1] if len(sys.argv) > 1 else "all") if __name__ == "__main__" else "none"
if which in ('all', 'spectral'):
    nchan, ny, nx = 1024, 2048, 2048
    clean = os.environ.get('SC_BENCH_CLEAN') == '1'
    dev = synth_cube(nchan, ny, nx, border=0, nan_permille=0) if clean else synth_cube(nchan, ny, nx, border=51)
    c = scb.DaskSpectralCube(dev, benchmark_wcs(nchan, ny, nx), unit='K')
    c._mask = scb.LazyMask(np.isfinite, cube=c)
    vox = nchan * ny * nx
    k = scb.Gaussian1DKernel(5 / 2.3548200450309493)
    ms = timeit(lambda: c._run_spectral_smooth(k.array, _lib.F32))
    print("spectral_smooth f32 17 taps   %.3f ms  %.1f GB/s (8 B/vox)  %.3e vox/s" % (ms, vox * 8 / ms / 1e6, vox / ms * 1e3), flush=True)
    sm = c.spectral_smooth(k)
    ms = timeit(lambda: sm._moments_axis0_raw(2))
    print("fused smooth->moment1         %.3f ms  %.1f GB/s (4 B/vox)  %.3e vox/s" % (ms, vox * 4 / ms / 1e6, vox / ms * 1e3), flush=True)
    for sig in (1.0, 4.0):
        kk = scb.Gaussian1DKernel(sig)
        ms = timeit(lambda: c._run_spectral_smooth(kk.array, _lib.F32))
        print("spectral_smooth f32 %2d taps   %.3f ms  %.1f GB/s" % (kk.array.size, ms, vox * 8 / ms / 1e6), flush=True)
    del dev, c, sm
    torch.cuda.empty_cache()
if which in ('all', 'spatial'):
    nchan, ny, nx = 512, 512, 4096          # one of 8 row shards of config 4
    dev = synth_cube(nchan, ny, nx, border=0)
    c = scb.DaskSpectralCube(dev, benchmark_wcs(nchan, ny, nx), unit='K')
    c._mask = scb.LazyMask(np.isfinite, cube=c)
    vox = nchan * ny * nx
    for fw in (8.0, 4.0, 2.0):
        k = scb.Gaussian2DKernel(fw / 2.3548200450309493)
        ms = timeit(lambda: c._run_spatial_smooth(k.array, _lib.F32))
        print("spatial_smooth f32 %dx%d shard 512x512x4096  %.3f ms  %.1f GB/s (8 B/vox)  %.3e vox/s" % (k.shape[0], k.shape[1], ms, vox * 8 / ms / 1e6, vox / ms * 1e3), flush=True)
    dev2 = synth_cube(nchan, ny, nx, border=0, nan_permille=0)
    c2 = scb.DaskSpectralCube(dev2, benchmark_wcs(nchan, ny, nx), unit='K')
    k = scb.Gaussian2DKernel(8.0 / 2.3548200450309493)
    ms = timeit(lambda: c2._run_spatial_smooth(k.array, _lib.F32))
    print("spatial_smooth f32 29x29 no NaNs              %.3f ms  %.1f GB/s" % (ms, vox * 8 / ms / 1e6), flush=True)
