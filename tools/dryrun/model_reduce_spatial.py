"""Python model of reduce_axis2_kernel's warp merge (csrc/reduce_spatial.cu): 32 lanes stride over a row, keep
shifted sums, and merge (count, sum, M2, extrema with index) through the xor butterfly.  TEST TOOLING: checked against
numpy's nan-functions in tests/test_reduce_host.py before the kernel could run on hardware."""
import numpy as np

INT_MAX = 2 ** 31 - 1


def lane_pass(row, lane):
    n, k, s1, s2, lo, hi, ilo, ihi = 0, 0.0, 0.0, 0.0, np.float32(0), np.float32(0), 0, 0
    for x in range(lane, len(row), 32):
        v = row[x]
        if v == v:
            if n == 0:
                k = float(v)
            d = float(v) - k
            s1 += d
            s2 += d * d
            if v < lo or n == 0:
                lo, ilo = v, x
            if v > hi or n == 0:
                hi, ihi = v, x
            n += 1
    total = n * k + s1
    m2 = s2 - s1 * s1 / n if n > 0 else 0.0
    return [n, total, m2, lo, hi, ilo if n > 0 else INT_MAX, ihi if n > 0 else INT_MAX]


def merge(a, b):
    n, total, m2, lo, hi, ilo, ihi = a
    n_b, total_b, m2_b, lo_b, hi_b, ilo_b, ihi_b = b
    if n_b > 0:
        if n > 0:
            delta = total_b / n_b - total / n
            m2 = (m2 + m2_b) + delta * delta * (n * n_b / (n + n_b))
            total = total + total_b
            if lo_b < lo or (lo_b == lo and ilo_b < ilo):
                lo, ilo = lo_b, ilo_b
            if hi_b > hi or (hi_b == hi and ihi_b < ihi):
                hi, ihi = hi_b, ihi_b
            n += n_b
        else:
            n, total, m2, lo, hi, ilo, ihi = b
    return [n, total, m2, lo, hi, ilo, ihi]


def reduce_row(row):
    """What lane 0 stores for one (channel, y) row: n, sum, M2, min, max, argmin, argmax."""
    lanes = [lane_pass(row, lane) for lane in range(32)]
    off = 16
    while off > 0:
        lanes = [merge(lanes[i], lanes[i ^ off]) for i in range(32)]
        off >>= 1
    assert all(l[:3] == lanes[0][:3] and l[5:] == lanes[0][5:] for l in lanes)       # every lane ends with the same bits
    n, total, m2, lo, hi, ilo, ihi = lanes[0]
    if n == 0:
        return 0, np.nan, np.nan, np.nan, np.nan, 0, 0
    return n, total, max(m2, 0.0), lo, hi, ilo, ihi
