"""spatial_smooth on the kinds of shards config 4 produces, sparse-denominator kernel against the
convolved-denominator one (GPU box only; scratch tool)."""
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


nchan, ny, nx = 512, 512, 4096
k = scb.Gaussian2DKernel(8.0 / 2.3548200450309493)
cases = (("clean shard, 0.1 % NaNs", dict(border=0), False),
         ("top shard of config 4 (102 blank rows + blank side columns)", dict(y0=0, ny_total=4096, nx_total=4096, border=102), False),
         ("interior shard of config 4 (blank side columns)", dict(y0=1024, ny_total=4096, nx_total=4096, border=102), False),
         ("masked > 3 sigma (crowded everywhere)", dict(border=0), True))
for name, kw, masked in cases:
    dev = synth_cube(nchan, ny, nx, **kw)
    c = scb.DaskSpectralCube(dev, benchmark_wcs(nchan, ny, nx), unit="K")
    c._mask = scb.LazyMask(np.isfinite, cube=c)
    if masked:
        c = c.with_mask(c > 3.0)
    for choice in ("0", "2", "3"):
        os.environ["SC_SPATIAL_KERNEL"] = choice
        ms = timeit(lambda: c._run_spatial_smooth(k.array, _lib.F32))
        print("%-62s kernel=%s  %.2f ms" % (name, {"0": "auto  ", "2": "sparse", "3": "march "}[choice], ms), flush=True)
    del dev, c
