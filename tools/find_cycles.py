"""Which public calls leave reference cycles behind (a cube in a cycle keeps its HBM until the cyclic GC happens to run)."""
import gc, os, sys, weakref, warnings
from collections import Counter
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectral_cube_b200 as scb
from spectral_cube_b200.synth import synth_cube, benchmark_wcs
warnings.simplefilter('ignore')
nchan, ny, nx = 16, 64, 64
w = benchmark_wcs(nchan, ny, nx)

def base(cls=scb.SpectralCube):
    c = cls(synth_cube(nchan, ny, nx), w, unit='K', allow_huge_operations=True)
    c._mask = scb.LazyMask(np.isfinite, cube=c)
    return c

a = np.radians(30.0)
hdr = dict(w.to_header())
hdr.update({'NAXIS': 3, 'NAXIS1': nx, 'NAXIS2': ny, 'NAXIS3': nchan, 'PC1_1': np.cos(a), 'PC1_2': -np.sin(a), 'PC2_1': np.sin(a), 'PC2_2': np.cos(a)})
ops = {
    'construct': lambda c: None,
    'with_mask': lambda c: c.with_mask(c > 3.0),
    'moment0': lambda c: c.moment0(),
    'moment1.value': lambda c: c.with_mask(c > 3.0).moment1().value,
    'spectral_smooth': lambda c: c.spectral_smooth(scb.Gaussian1DKernel(2.0)),
    'spatial_smooth': lambda c: c.spatial_smooth(scb.Gaussian2DKernel(2.0)),
    'spectral_interpolate': lambda c: c.spectral_interpolate(c.spectral_axis[::2]),
    'reproject': lambda c: c.reproject(hdr),
    'sum axis 0': lambda c: c.sum(axis=0),
}
for cls in (scb.SpectralCube, scb.DaskSpectralCube):
    for name, op in ops.items():
        gc.collect(); gc.disable(); gc.set_debug(0)
        c = base(cls)
        r = op(c)
        refs = [weakref.ref(c)] + ([weakref.ref(r)] if r is not None and hasattr(r, '__weakref__') else [])
        del c, r
        alive = [x() is not None for x in refs]
        gc.set_debug(gc.DEBUG_SAVEALL)
        n = gc.collect()
        kinds = Counter(type(o).__name__ for o in gc.garbage).most_common(6)
        gc.garbage.clear(); gc.set_debug(0); gc.enable()
        print('%-18s %-22s alive after del: %s  cyclic garbage: %d %s' % (cls.__name__, name, alive, n, kinds if n else ''))
