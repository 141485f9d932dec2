"""spectral_interpolate 2048 -> 1024 channels on a config-5 row shard, ring variants (scratch tool)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectral_cube_b200 as scb
from spectral_cube_b200.synth import synth_cube, benchmark_wcs
nchan, ny, nx, nout = 2048, 512, 4096, 1024
dev = synth_cube(nchan, ny, nx, y0=1024, ny_total=4096, nx_total=nx, border=102)
w = benchmark_wcs(nchan, 4096, nx)
c = scb.SpectralCube(dev, w, unit='K'); c._mask = scb.LazyMask(np.isfinite, cube=c)
sa = c.spectral_axis; grid = np.linspace(sa[0], sa[-1], nout)
def timeit(f, n=5, warm=2):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
variants = sys.argv[1:] or ['0', '1', '2', '3']
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    for v in variants:
        os.environ['SC_INTERP_RING'] = v
        print('ring variant %s: %.2f ms' % (v, timeit(lambda: c.spectral_interpolate(grid))), flush=True)
