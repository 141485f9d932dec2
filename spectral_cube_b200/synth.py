"""
Synthetic Gaussian-line cubes generated on the device (``sc_synth_cube``), used by bench.py
and the parity tests; see csrc/synth.cu for the construction.
"""
import numpy as np

from . import _lib

DEFAULT_SEED = 247825498          # spectral_cube/tests/utilities.py:59


def line_profile(nchan, sigma=8.0):
    k = np.arange(16 * nchan + 1, dtype=np.float64) / 16.0
    return np.exp(-0.5 * (k / sigma) ** 2).astype(np.float32)


def synth_cube(nchan, ny, nx, y0=0, x0=0, ny_total=None, nx_total=None, seed=DEFAULT_SEED,
               nan_permille=1, border=0, out=None, device=None):
    """float32 device tensor (nchan, ny, nx): rows [y0, y0+ny) x cols [x0, x0+nx) of the full
    (nchan, ny_total, nx_total) synthetic cube."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    ny_total = ny if ny_total is None else ny_total
    nx_total = nx if nx_total is None else nx_total
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else device)
    if out is None:
        out = torch.empty((nchan, ny, nx), dtype=torch.float32, device=dev)
    prof = torch.from_numpy(line_profile(nchan)).to(dev)
    _lib.check(lib.sc_synth_cube(out.data_ptr(), nchan, ny, nx, y0, x0, ny_total, nx_total,
                                 int(seed), prof.data_ptr(), int(nan_permille), int(border),
                                 torch.cuda.current_stream().cuda_stream))
    torch.cuda.current_stream().synchronize()     # prof must outlive the kernel
    return out


def benchmark_wcs(nchan, ny, nx):
    """The WCS the benchmark cubes carry: linear velocity axis and TAN celestial axes with the
    reference's test header values (spectral_cube/tests/data/header_jybeam.hdr:12-26)."""
    from .wcs import CubeWCS
    return CubeWCS(ctype=['RA---TAN', 'DEC--TAN', 'VRAD'],
                   crval=[24.0, 30.0, -321.214698632],
                   crpix=[nx / 2.0 + 0.5, ny / 2.0 + 0.5, 1.0],
                   cdelt=[-5.55555561268e-4, 5.55555561268e-4, 1.28821496879],
                   cunit=['deg', 'deg', 'km/s'])
