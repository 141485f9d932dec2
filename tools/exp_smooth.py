import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectral_cube_b200 as scb
from spectral_cube_b200 import _lib
from spectral_cube_b200.synth import synth_cube, benchmark_wcs
from tools.time_smooth import timeit
nchan, ny, nx = 1024, 2048, 2048
vox = nchan*ny*nx
k = scb.Gaussian1DKernel(5 / 2.3548200450309493)
for name, kw in [('bench data (0.1% NaN + border)', dict(border=51, nan_permille=1)), ('no NaN at all', dict(border=0, nan_permille=0))]:
    dev = synth_cube(nchan, ny, nx, **kw)
    c = scb.DaskSpectralCube(dev, benchmark_wcs(nchan, ny, nx), unit='K')
    c._mask = scb.LazyMask(np.isfinite, cube=c)
    for dbg in (0, 1, 2, 3):
        os.environ['SC_SMOOTH_DEBUG'] = str(dbg)
        ms = timeit(lambda: c._run_spectral_smooth(k.array, _lib.F32))
        print("%-32s debug=%d  %.3f ms  %.1f GB/s" % (name, dbg, ms, vox*8/ms/1e6), flush=True)
    os.environ['SC_SMOOTH_DEBUG'] = '0'
    sm = c.spectral_smooth(k)
    ms = timeit(lambda: sm._moments_axis0_raw(2))
    print("%-32s fused->moment1 %.3f ms" % (name, ms), flush=True)
    del dev, c, sm; torch.cuda.empty_cache()
