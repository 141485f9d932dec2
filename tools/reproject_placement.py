"""Does the time of the reproject kernel depend on WHERE its buffers were allocated?  (run-to-run it is 9.9, 16.5 or 26.5 ms)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectral_cube_b200 as scb
from spectral_cube_b200.synth import synth_cube, benchmark_wcs
from spectral_cube_b200.wcs import as_cube_wcs
nloc, ny, nx = 128, 4096, 4096
w = benchmark_wcs(nloc, ny, nx)
a = np.radians(30.0)
hdr = dict(w.to_header())
hdr.update({'NAXIS': 3, 'NAXIS1': nx, 'NAXIS2': ny, 'NAXIS3': nloc, 'PC1_1': np.cos(a), 'PC1_2': -np.sin(a), 'PC2_1': np.sin(a), 'PC2_2': np.cos(a)})
neww = as_cube_wcs(hdr)
def ev(f, n=3):
    f(); torch.cuda.synchronize()
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_.record()
    for _ in range(n): r = f()
    b_.record(); torch.cuda.synchronize()
    return a_.elapsed_time(b_) / n
keep = []
for trial in range(6):
    planes = synth_cube(nloc, ny, nx, border=102)
    cc = scb.SpectralCube(planes, w, unit='K', allow_huge_operations=True)
    cc._mask = scb.LazyMask(np.isfinite, cube=cc)
    yin, xin = cc._pixel_map(neww, ny, nx)
    ms = ev(lambda: cc._run_reproject(yin, xin, 1, want_f32=False))
    out = cc._run_reproject(yin, xin, 1, want_f32=False)
    t = out[0] if isinstance(out, (tuple, list)) else out
    print('trial %d: %.2f ms   in @ %#x  out @ %#x  yin @ %#x' % (trial, ms, planes.data_ptr(), t.data_ptr(), yin.data_ptr()), flush=True)
    keep.append((planes, yin, xin))        # keep the inputs alive: the next trial gets other addresses
    if trial == 2:
        keep.clear(); torch.cuda.empty_cache()
