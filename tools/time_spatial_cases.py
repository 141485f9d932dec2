"""spatial_smooth on the kinds of shards config 4 produces: the pipelined kernel (round 2), the round-1 sparse kernel and
the convolved-denominator kernel (GPU box only; scratch tool).  `python tools/time_spatial_cases.py one` runs the clean
shard once per kernel (for ncu)."""
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
if len(sys.argv) > 1 and sys.argv[1] in ('one', 'one_interior', 'one_edges'):
    if sys.argv[1] == 'one_edges':
        nx = 256                                   # two strips, both with the blank frame's 102 side columns
        dev = synth_cube(nchan * 4, ny, nx, y0=1024, ny_total=4096, nx_total=256, border=102)
    else:
        dev = synth_cube(nchan, ny, nx, border=0) if sys.argv[1] == 'one' else synth_cube(nchan, ny, nx, y0=1024, ny_total=4096, nx_total=4096, border=102)
    c = scb.DaskSpectralCube(dev, benchmark_wcs(dev.shape[0], ny, nx), unit="K")
    c._mask = scb.LazyMask(np.isfinite, cube=c)
    os.environ["SC_SPATIAL_KERNEL"] = "5"
    if len(sys.argv) > 2:
        os.environ["SC_SPATIAL_J"] = sys.argv[2]
    for _ in range(3):
        c._run_spatial_smooth(k.array, _lib.F32)
    torch.cuda.synchronize()
    print("%s: %.2f ms for %d voxels" % (sys.argv[1], timeit(lambda: c._run_spatial_smooth(k.array, _lib.F32)), dev.numel()))
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == 'sweep':            # scattered NaNs at rising density: where does the march win?
    for permille in (10, 30, 60, 100, 200, 400):
        dev = synth_cube(nchan, ny, nx, border=0, nan_permille=permille)
        c = scb.DaskSpectralCube(dev, benchmark_wcs(nchan, ny, nx), unit="K")
        c._mask = scb.LazyMask(np.isfinite, cube=c)
        line = "scattered NaNs %4.1f %%:" % (permille / 10)
        for choice in ("0", "5", "3"):
            os.environ["SC_SPATIAL_KERNEL"] = choice
            os.environ.pop("SC_SPATIAL_J", None)
            line += "  %s %.2f ms" % ({"0": "auto", "5": "pipe", "3": "march"}[choice], timeit(lambda: c._run_spatial_smooth(k.array, _lib.F32)))
        print(line, "  counts", c._spatial_strategy_counts().tolist(), flush=True)
        del dev, c
    sys.exit(0)
for name, kw, masked in cases:
    dev = synth_cube(nchan, ny, nx, **kw)
    c = scb.DaskSpectralCube(dev, benchmark_wcs(nchan, ny, nx), unit="K")
    c._mask = scb.LazyMask(np.isfinite, cube=c)
    if masked:
        c = c.with_mask(c > 3.0)
    for choice in ("0", "5j16", "5j8", "4", "3"):
        os.environ["SC_SPATIAL_KERNEL"] = choice[0]
        os.environ.pop("SC_SPATIAL_J", None)                      # "auto" is the library's own choice of kernel and of J
        if choice.startswith("5j"):
            os.environ["SC_SPATIAL_J"] = choice[2:]
        ms = timeit(lambda: c._run_spatial_smooth(k.array, _lib.F32))
        print("%-62s kernel=%s  %.2f ms" % (name, {"0": "auto  ", "5j16": "pipe J=16", "5j8": "pipe J=8", "4": "sparse(r1)", "3": "march "}[choice], ms), flush=True)
    del dev, c
