"""
Host logic of the whole-cube reductions (no GPU): the per-spaxel partials `sc_reduce_axis0` returns are
combined into `sum/mean/std/max/min/argmax/argmin(axis=None)` by device-agnostic torch code; here the
partials are made with numpy and the combination is checked against numpy's nan-functions on the whole
cube (what `apply_numpy_function(..., axis=None)` computes, spectral_cube.py:446-454).
"""
import warnings

import numpy as np
import pytest
import torch

from spectral_cube_b200 import cube as C
from oracle.cube import OracleCube
from oracle.wcs import OWCS

WCS = dict(ctype=['RA---TAN', 'DEC--TAN', 'VRAD'], crval=[24.0, 30.0, -321.214698632], crpix=[8.5, 8.5, 1.0],
           cdelt=[-5.55555561268e-4, 5.55555561268e-4, 1.28821496879], cunit=['deg', 'deg', 'km/s'])


def partials(data):
    """What the device pass returns per spaxel (reduce.cu), with numpy."""
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        d = data.astype(np.float64)
        cnt = (~np.isnan(d)).sum(axis=0).astype(np.int32)
        s = np.where(cnt > 0, np.nansum(d, axis=0), np.nan)
        mean = s / np.maximum(cnt, 1)
        m2 = np.where(cnt > 0, np.nansum((d - mean) ** 2, axis=0), np.nan)
        hi = np.nanmax(data, axis=0); lo = np.nanmin(data, axis=0)
        ahi = np.argmax(np.where(np.isnan(data), -np.inf, data), axis=0); ahi[cnt == 0] = 0
        alo = np.argmin(np.where(np.isnan(data), np.inf, data), axis=0); alo[cnt == 0] = 0
    t = torch.from_numpy
    return dict(sum=t(s), count=t(cnt), m2=t(m2), max=t(hi.astype(np.float32)), min=t(lo.astype(np.float32)),
                argmax=t(ahi.astype(np.int32)), argmin=t(alo.astype(np.int32)))


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_whole_cube_statistics_from_partials(seed):
    rng = np.random.default_rng(seed)
    data = (5.0 + rng.normal(0, 2, (17, 6, 9))).astype(np.float32)
    data[rng.random(data.shape) < 0.1] = np.nan
    data[:, 0, :] = np.nan                                              # blank spaxels
    r = partials(data)
    d64 = data.astype(np.float64)
    assert np.isclose(C.whole_sum(torch, r['sum'], r['count']), np.nansum(d64), rtol=1e-12)
    assert np.isclose(C.whole_mean(torch, r['sum'], r['count']), np.nanmean(d64), rtol=1e-12)
    for ddof in (0, 1):
        assert np.isclose(C.whole_std(torch, r['sum'], r['count'], r['m2'], ddof), np.nanstd(d64, ddof=ddof), rtol=1e-10)
    assert C.whole_extremum(torch, r['max'], 'max') == np.nanmax(data)
    assert C.whole_extremum(torch, r['min'], 'min') == np.nanmin(data)
    assert C.whole_arg_extremum(torch, r['max'], r['argmax'], 'max') == int(np.nanargmax(data))
    assert C.whole_arg_extremum(torch, r['min'], r['argmin'], 'min') == int(np.nanargmin(data))


def test_ties_take_the_first_in_c_order_and_blank_cubes_give_nan():
    tie = np.zeros((5, 3, 4), dtype=np.float32)
    tie[2, 1, 1] = tie[1, 2, 3] = tie[1, 0, 2] = 9.0
    r = partials(tie)
    assert C.whole_arg_extremum(torch, r['max'], r['argmax'], 'max') == int(np.argmax(tie))
    blank = partials(np.full((4, 3, 4), np.nan, dtype=np.float32))
    assert np.isnan(C.whole_sum(torch, blank['sum'], blank['count']))
    assert np.isnan(C.whole_mean(torch, blank['sum'], blank['count']))
    assert np.isnan(C.whole_std(torch, blank['sum'], blank['count'], blank['m2']))
    assert np.isnan(C.whole_extremum(torch, blank['max'], 'max'))
    assert C.whole_arg_extremum(torch, blank['max'], blank['argmax'], 'max') == 0


def test_oracle_whole_cube_reductions_are_the_numpy_nan_functions():
    rng = np.random.default_rng(4)
    data = rng.normal(0, 1, (9, 4, 5)).astype(np.float32)
    data[rng.random(data.shape) < 0.2] = np.nan
    oc = OracleCube(data, OWCS(**WCS), unit='K')
    oc = oc.with_mask(oc > -0.5)
    filled = np.where((data > -0.5), data, np.nan)
    assert np.isclose(oc.sum(), np.nansum(filled)) and np.isclose(oc.mean(), np.nanmean(filled))
    assert np.isclose(oc.std(), np.nanstd(filled)) and oc.max() == np.nanmax(filled) and oc.min() == np.nanmin(filled)
    assert np.isnan(OracleCube(np.full((2, 2, 2), np.nan, dtype=np.float32), OWCS(**WCS), unit='K').sum())


@pytest.mark.parametrize('seed', [0, 3])
def test_statistics_from_partials(seed):
    """`DaskSpectralCube.statistics` (dask_spectral_cube.py:769-814) from the partial maps of one pass."""
    rng = np.random.default_rng(seed)
    data = (1.5 + rng.normal(0, 2, (17, 6, 9))).astype(np.float32)
    data[rng.random(data.shape) < 0.1] = np.nan
    data[:, 0, :] = np.nan
    r = partials(data)
    got = C.whole_statistics(torch, r['sum'], r['count'], r['m2'], r['min'], r['max'])
    d = data.astype(np.float64)
    n = int((~np.isnan(d)).sum())
    want = {'npts': n, 'min': np.nanmin(d), 'max': np.nanmax(d), 'sum': np.nansum(d), 'sumsq': np.nansum(d * d)}
    want['mean'] = want['sum'] / n
    want['sigma'] = np.sqrt((want['sumsq'] - want['sum'] ** 2 / n) / (n - 1))
    want['rms'] = np.sqrt(want['sumsq'] / n)
    assert set(got) == set(want) and got['npts'] == n
    for key in want:
        assert np.isclose(got[key], want[key], rtol=1e-12), key
    # the oracle restates the reference in the data's own float32 (one chunk): same numbers to float32 accuracy
    oc = OracleCube(data, OWCS(**WCS), unit='K', use_dask=True).statistics()
    for key in want:
        assert np.isclose(oc[key], want[key], rtol=2e-6), key
    blank = partials(np.full((4, 3, 4), np.nan, dtype=np.float32))
    got = C.whole_statistics(torch, blank['sum'], blank['count'], blank['m2'], blank['min'], blank['max'])
    assert got['npts'] == 0 and got['sum'] == 0.0 and got['sumsq'] == 0.0
    assert all(np.isnan(got[k]) for k in ('min', 'max', 'mean', 'sigma', 'rms'))


@pytest.mark.parametrize('nx', [1, 31, 32, 33, 100, 257])
def test_model_of_the_spatial_reduction_merge_matches_numpy(nx):
    """reduce_axis2_kernel (opt-in, csrc/reduce_spatial.cu) restated in Python: lanes striding over a row, shifted sums,
    xor-butterfly merge with first-occurrence ties -- against numpy's nan-functions."""
    from tools.dryrun.model_reduce_spatial import reduce_row
    rng = np.random.default_rng(nx)
    row = (3.0 + rng.normal(0, 2, nx)).astype(np.float32)
    row[rng.random(nx) < 0.2] = np.nan
    if nx > 40:
        row[7] = row[39] = np.nanmax(row) + 1.0          # a tie between two lanes: the first position wins
        row[5] = row[70] = np.nanmin(row) - 1.0
    n, total, m2, lo, hi, ilo, ihi = reduce_row(row)
    d = row.astype(np.float64)
    assert n == int((~np.isnan(d)).sum())
    if n == 0:
        assert np.isnan(total) and np.isnan(m2) and (ilo, ihi) == (0, 0)
        return
    assert np.isclose(total, np.nansum(d), rtol=1e-13)
    assert np.isclose(m2, np.nansum((d - np.nanmean(d)) ** 2), rtol=1e-10, atol=1e-12)
    assert lo == np.nanmin(row) and hi == np.nanmax(row)
    assert ilo == int(np.nanargmin(row)) and ihi == int(np.nanargmax(row))
    assert reduce_row(np.full(nx, np.nan, dtype=np.float32))[0] == 0
