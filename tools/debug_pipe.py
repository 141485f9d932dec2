import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectral_cube_b200 as scb
from spectral_cube_b200 import _lib
from spectral_cube_b200.synth import synth_cube, benchmark_wcs
ny, nx, y0 = 72, 4096, 60
dev = synth_cube(1, ny, nx, y0=y0, ny_total=4096, nx_total=nx, nan_permille=1, border=102)
c = scb.DaskSpectralCube(dev, benchmark_wcs(1, ny, nx), unit='K'); c._mask = scb.LazyMask(np.isfinite, cube=c)
k = scb.Gaussian2DKernel(8 / 2.3548200450309493)
res = {}
for name, kern, j in (('march', '3', '16'), ('pipe8', '5', '8'), ('pipe16', '5', '16'), ('sparse', '4', '16')):
    os.environ['SC_SPATIAL_KERNEL'] = kern; os.environ['SC_SPATIAL_J'] = j
    res[name] = c._run_spatial_smooth(k.array, _lib.F32).cpu().numpy()[0].astype(np.float64)
ref = res['march']
host = dev.cpu().numpy()[0]
for name in ('pipe8', 'pipe16', 'sparse'):
    g = res[name]
    bad = np.abs(g - ref) > 3e-6 * np.abs(ref)
    bad |= np.isnan(g) != np.isnan(ref)
    ys, xs = np.nonzero(bad)
    print(name, 'differs at', len(ys))
    for y, x in list(zip(ys, xs))[:6]:
        win = host[max(0, y - 14):y + 15, max(0, x - 14):x + 15]
        print('   y %d x %d got %.8g ref %.8g rel %.3g  missing in window %d, strip %d col %d, block row %d' % (
            y, x, g[y, x], ref[y, x], abs(g[y, x] - ref[y, x]) / abs(ref[y, x]), int(np.isnan(win).sum()), x // 128, x % 128, y % 16))
