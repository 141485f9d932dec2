"""Time the moment kernel variants on the config-2 cube (GPU box only; scratch tool)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectral_cube_b200 as scb
from spectral_cube_b200.synth import synth_cube, benchmark_wcs

nchan, ny, nx = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (1024, 2048, 2048))]
dev = synth_cube(nchan, ny, nx, border=51)
c = scb.SpectralCube(dev, benchmark_wcs(nchan, ny, nx), unit='K')
c._mask = scb.LazyMask(np.isfinite, cube=c)
c = c.with_mask(c > 3.0)
vox = nchan * ny * nx
res = {}
variants = [('tma cb8 s4', dict(SC_MOM_KERNEL='2', SC_MOM_TMA_CFG='0')),
            ('tma cb4 s8', dict(SC_MOM_KERNEL='2', SC_MOM_TMA_CFG='1')),
            ('tma cb8 s6', dict(SC_MOM_KERNEL='2', SC_MOM_TMA_CFG='2')),
            ('tma cb16 s3', dict(SC_MOM_KERNEL='2', SC_MOM_TMA_CFG='3')),
            ('direct u4', dict(SC_MOM_KERNEL='1', SC_MOM_UNROLL='4')),
            ('direct u8', dict(SC_MOM_KERNEL='1', SC_MOM_UNROLL='8'))]
for name, env in variants:
    for k in ('SC_MOM_KERNEL', 'SC_MOM_TMA_CFG', 'SC_MOM_UNROLL'):
        os.environ.pop(k, None)
    os.environ.update(env)
    for want in (1, 3, 7):
        for _ in range(3):
            c._moments_axis0_raw(want)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n = 10
        for _ in range(n):
            c._moments_axis0_raw(want)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        print("%-12s want=%d  %.3f ms  %.1f GB/s  %.3e vox/s" % (name, want, ms, vox * 4 / ms / 1e6, vox / ms * 1e3), flush=True)
