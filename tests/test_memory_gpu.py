"""
Device memory is released by reference counting (``-m gpu``).  A result tensor caught in a reference cycle stays in
HBM until Python's cyclic collector happens to run; with 17 GB results that filled the 180 GB of a B200 inside a
timing loop and `reproject` ran 60 % slower (the allocator had to flush and retry on every call).  No public call
may leave a tensor, a cube or a pending-result object in cyclic garbage.
"""
import gc
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ops():
    import spectral_cube_b200 as scb
    from spectral_cube_b200.synth import benchmark_wcs
    nchan, ny, nx = 16, 64, 64
    w = benchmark_wcs(nchan, ny, nx)
    a = np.radians(30.0)
    hdr = dict(w.to_header())
    hdr.update({'NAXIS': 3, 'NAXIS1': nx, 'NAXIS2': ny, 'NAXIS3': nchan,
                'PC1_1': np.cos(a), 'PC1_2': -np.sin(a), 'PC2_1': np.sin(a), 'PC2_2': np.cos(a)})
    return {
        'with_mask': lambda c: c.with_mask(c > 3.0),
        'moment1': lambda c: c.with_mask(c > 3.0).moment1(),
        'spectral_smooth': lambda c: c.spectral_smooth(scb.Gaussian1DKernel(2.0)),
        'spatial_smooth': lambda c: c.spatial_smooth(scb.Gaussian2DKernel(2.0)),
        'spectral_interpolate': lambda c: c.spectral_interpolate(c.spectral_axis[::2]),
        'reproject': lambda c: c.reproject(hdr),
        'reproject then moment0': lambda c: c.reproject(hdr).moment0(),
        'sum': lambda c: c.sum(axis=0),
    }


@pytest.mark.parametrize('cls_name', ['SpectralCube', 'DaskSpectralCube'])
@pytest.mark.parametrize('op', ['with_mask', 'moment1', 'spectral_smooth', 'spatial_smooth', 'spectral_interpolate',
                                'reproject', 'reproject then moment0', 'sum'])
def test_results_are_freed_without_the_cyclic_collector(cls_name, op):
    import torch
    import spectral_cube_b200 as scb
    from spectral_cube_b200.synth import synth_cube, benchmark_wcs
    cls = getattr(scb, cls_name)
    gc.collect()
    gc.disable()
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            torch.cuda.synchronize()
            before = torch.cuda.memory_allocated()
            c = cls(synth_cube(16, 64, 64), benchmark_wcs(16, 64, 64), unit='K', allow_huge_operations=True)
            c._mask = scb.LazyMask(np.isfinite, cube=c)
            r = _ops()[op](c)
            del c, r
            torch.cuda.synchronize()
            assert torch.cuda.memory_allocated() == before, 'device memory is still held after the last reference went away'
            gc.set_debug(gc.DEBUG_SAVEALL)
            gc.collect()
            held = [type(o).__name__ for o in gc.garbage
                    if isinstance(o, torch.Tensor) or type(o).__module__.startswith('spectral_cube_b200')]
            assert not held, 'objects freed only by the cyclic collector: %s' % held
    finally:
        gc.set_debug(0)
        gc.garbage.clear()
        gc.enable()
