"""
oracle/synth.py -- TEST INFRASTRUCTURE.  numpy twin of ``sc_synth_cube``
(spectral_cube_b200/csrc/synth.cu): regenerates any sub-block of the synthetic benchmark
cube bit-identically on the CPU, so the GPU path can be checked against the oracle on
exactly the voxels the benchmark uses.

The cube mirrors the reference's own test generator
(``spectral_cube/tests/utilities.py:53-112``: one Gaussian line of sigma = 8 channels per
spaxel plus unit-variance noise) but every number comes from a counter-based integer hash
of (seed, global voxel index), integer arithmetic, one table lookup and two
correctly-rounded float32 operations -- nothing that differs between libm and CUDA.
"""
import numpy as np

_U = np.uint64
GOLD, C1, C2 = _U(0x9E3779B97F4A7C15), _U(0xBF58476D1CE4E5B9), _U(0x94D049BB133111EB)
STREAM_AMP, STREAM_CEN, STREAM_NAN, STREAM_N1, STREAM_N2 = (
    _U(0x1000000000000000), _U(0x2000000000000000), _U(0x3000000000000000),
    _U(0x4000000000000000), _U(0x5000000000000000))
DEFAULT_SEED = 247825498          # spectral_cube/tests/utilities.py:59


def mix64(k):
    with np.errstate(over='ignore'):
        z = k * GOLD
        z = (z ^ (z >> _U(30))) * C1
        z = (z ^ (z >> _U(27))) * C2
        return z ^ (z >> _U(31))


def _sum16x4(h):
    m = _U(0xFFFF)
    return ((h & m) + ((h >> _U(16)) & m) + ((h >> _U(32)) & m) + (h >> _U(48))).astype(np.int64)


def line_profile(nchan, sigma=8.0):
    """float32 table of exp(-(k/16)^2 / (2 sigma^2)), k = 0 .. 16*nchan (the host builds it once
    and both sides read it)."""
    k = np.arange(16 * nchan + 1, dtype=np.float64) / 16.0
    return np.exp(-0.5 * (k / sigma) ** 2).astype(np.float32)


def synth_block(nchan, ny, nx, y0=0, x0=0, ny_total=None, nx_total=None, seed=DEFAULT_SEED,
                nan_permille=1, border=0, profile=None):
    ny_total = ny if ny_total is None else ny_total
    nx_total = nx if nx_total is None else nx_total
    if profile is None:
        profile = line_profile(nchan)
    seed = _U(seed)
    Y = (np.arange(ny, dtype=np.int64) + y0)[:, None]
    X = (np.arange(nx, dtype=np.int64) + x0)[None, :]
    sp = (Y * nx_total + X).astype(np.uint64)                       # (ny, nx)
    with np.errstate(over='ignore'):
        ha = mix64(seed + STREAM_AMP + sp)
        hc = mix64(seed + STREAM_CEN + sp)
    amp = (ha >> _U(40)).astype(np.float32) * np.float32(10.0 / 16777216.0)
    c0_16 = 4 * nchan + (hc % _U(8 * nchan)).astype(np.int64)
    in_border = (Y < border) | (X < border) | (Y >= ny_total - border) | (X >= nx_total - border)
    nscale = np.float32(1.0 / 53510.0)
    out = np.empty((nchan, ny, nx), dtype=np.float32)
    plane = _U(ny_total * nx_total)
    for c in range(nchan):
        with np.errstate(over='ignore'):
            vi = _U(c) * plane + sp
            h1 = mix64(seed + STREAM_N1 + vi)
            h2 = mix64(seed + STREAM_N2 + vi)
            hn = mix64(seed + STREAM_NAN + vi)
        s = (_sum16x4(h1) + _sum16x4(h2) - 262140).astype(np.float32)
        noise = s * nscale
        k = np.abs(16 * c - c0_16)
        val = amp * profile[k] + noise          # two separately rounded float32 ops
        isnan = in_border.copy()
        if nan_permille > 0:
            isnan |= (((hn >> _U(32)) * _U(1000)) >> _U(32)) < _U(nan_permille)
        val[isnan] = np.nan
        out[c] = val
    return out
