"""Where the time of SpectralCube.reproject goes (config-5 channel shard: 128 planes of 4096 x 4096, 30 degrees)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectral_cube_b200 as scb
from spectral_cube_b200 import _lib
from spectral_cube_b200.synth import synth_cube, benchmark_wcs
from spectral_cube_b200.wcs import as_cube_wcs
nloc, ny, nx = 128, 4096, 4096
w = benchmark_wcs(nloc, ny, nx)
planes = synth_cube(nloc, ny, nx, border=102)
cc = scb.SpectralCube(planes, w, unit='K', allow_huge_operations=True)
cc._mask = scb.LazyMask(np.isfinite, cube=cc)
a = np.radians(30.0)
hdr = dict(w.to_header())
hdr.update({'NAXIS': 3, 'NAXIS1': nx, 'NAXIS2': ny, 'NAXIS3': nloc, 'PC1_1': np.cos(a), 'PC1_2': -np.sin(a), 'PC2_1': np.sin(a), 'PC2_2': np.cos(a)})
def ev(f, n=3):
    f(); torch.cuda.synchronize()
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a_.record()
    for _ in range(n): f()          # (do not keep the result: a second 17 GB block would be cudaMalloc'ed inside the timed region)
    b_.record(); torch.cuda.synchronize()
    return a_.elapsed_time(b_) / n, (time.perf_counter() - t0) / n * 1e3
neww = as_cube_wcs(hdr)
print('pixel map          %.2f ms (wall %.2f)' % ev(lambda: cc._pixel_map(neww, ny, nx)))
yin, xin = cc._pixel_map(neww, ny, nx)
for want in (False, True):
    print('run_reproject f32=%s %.2f ms (wall %.2f)' % ((want,) + ev(lambda: cc._run_reproject(yin, xin, 1, want_f32=want))))
print('reproject()        %.2f ms (wall %.2f)' % ev(lambda: cc.reproject(hdr)))
