"""reproject kernel alone (preallocated outputs) on a config-5 channel shard, channels-per-CTA variants (scratch tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectral_cube_b200 as scb
from spectral_cube_b200 import _lib
from spectral_cube_b200.masks import lower_mask
from spectral_cube_b200.synth import synth_cube, benchmark_wcs
from spectral_cube_b200.wcs import as_cube_wcs
lib = _lib.load()
nloc, ny, nx = 128, 4096, 4096
w = benchmark_wcs(nloc, ny, nx)
planes = synth_cube(nloc, ny, nx, border=102)
cc = scb.SpectralCube(planes, w, unit='K', allow_huge_operations=True)
cc._mask = scb.LazyMask(np.isfinite, cube=cc)
a = np.radians(float(os.environ.get('ANGLE', '30')))
hdr = dict(w.to_header())
hdr.update({'NAXIS': 3, 'NAXIS1': nx, 'NAXIS2': ny, 'NAXIS3': nloc, 'PC1_1': np.cos(a), 'PC1_2': -np.sin(a), 'PC2_1': np.sin(a), 'PC2_2': np.cos(a)})
yin, xin = cc._pixel_map(as_cube_wcs(hdr), ny, nx)
out = torch.empty((nloc, ny, nx), dtype=torch.float64, device='cuda')
foot = torch.empty((nloc, ny, nx), dtype=torch.uint8, device='cuda')
flag = torch.empty((1,), dtype=torch.int32, device='cuda')
desc, keep = cc._mask_desc()
st = torch.cuda.current_stream().cuda_stream
def run():
    _lib.check(lib.sc_reproject_ex(planes.data_ptr(), out.data_ptr(), _lib.F64, None, foot.data_ptr(), flag.data_ptr(), nloc, ny, nx,
                                   planes.stride(0), planes.stride(1), ny, nx, desc, float('nan'), yin.data_ptr(), xin.data_ptr(), 1, st))
def timeit(f, n=5, warm=2):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
V = nloc * ny * nx
for cap in (sys.argv[1:] or ['0', '64', '32', '16', '8']):
    os.environ['SC_REPROJECT_CHAN'] = cap
    ms = timeit(run)
    print('channels per CTA cap %3s: %.2f ms  (%.0f GB/s of 13 B/voxel)' % (cap, ms, 13 * V / ms / 1e6), flush=True)
