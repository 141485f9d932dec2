// reduce_spatial.cu -- masked one-pass reductions along a SPATIAL axis (numpy axis 1 = y, axis 2 = x): sum,
// count, min, max, argmin, argmax and the sum of squared deviations from ONE read of the cube.
//
// Same reference functions as reduce.cu (`apply_numpy_function(np.nansum / nanmean / nanstd / nanmax / nanmin /
// nanargmax / nanargmin, axis=1 or 2)`, spectral_cube.py:361-470, 578-826; dask_spectral_cube.py:641-767): a voxel
// takes part iff the mask includes it and it is not NaN.  SURVEY.md 8(f) item 2, the axes reduce.cu does not cover.
//  * axis 1: a thread owns one (channel, x) column and walks y -- a warp reads 128 contiguous bytes per row;
//    output (nchan, nx).
//  * axis 2: a warp owns one (channel, y) row, lanes stride over x (coalesced); every lane keeps its own shifted
//    sums and the 32 partial results are merged with warp shuffles in a fixed butterfly order (pairwise update of
//    count / sum / M2, Chan et al.); extrema travel with their index so that ties keep the FIRST position like
//    numpy's.  Output (nchan, ny).
// float64 sums about the first included value of the thread (the variance does not cancel), float32 extrema.
// HBM-bound: 4 B/voxel in.  Measured on a B200 (2048x2048x1024, all seven statistics): axis 1 4.49 ms = 0.59,
// axis 2 5.64 ms = 0.47 of the measured copy peak (profiles/r02_reduce_n1.jsonl).
#include "common.cuh"
#include <limits.h>

namespace scb {

int check_cube_args(const float *cube, int64_t nchan, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y);

struct SpRedParams {
    const float *in;
    int64_t nchan, ny, nx, stride_c, stride_y;
    double *sum, *m2;
    int32_t *count, *argmin, *argmax;
    float *vmin, *vmax;
    DevMask mask;
};

// what one thread has seen: n values, their sum as n k + s1, M2 = s2 - s1^2 / n about the first value k
struct SpAcc {
    double k, s1, s2;
    float lo, hi;
    int n, ilo, ihi;
};

__device__ __forceinline__ void sp_init(SpAcc &a) {
    a.k = 0.0; a.s1 = 0.0; a.s2 = 0.0; a.lo = 0.0f; a.hi = 0.0f; a.n = 0; a.ilo = 0; a.ihi = 0;
}

__device__ __forceinline__ void sp_take(SpAcc &a, float v, int idx) {
    if (a.n == 0) a.k = fabsf(v) <= FLT_MAX ? (double)v : 0.0;   // an infinite first value is no shift: inf - inf would poison the sum (np.nansum gives +-inf)
    const double d = (double)v - a.k;
    a.s1 += d;
    a.s2 = fma(d, d, a.s2);
    if (v < a.lo || a.n == 0) { a.lo = v; a.ilo = idx; }      // strict: the first occurrence wins
    if (v > a.hi || a.n == 0) { a.hi = v; a.ihi = idx; }
    a.n += 1;
}

__device__ __forceinline__ void sp_store(const SpRedParams &p, int64_t o, int n, double total, double m2,
                                         float lo, float hi, int ilo, int ihi) {
    const bool any = n > 0;
    if (p.sum)    p.sum[o] = any ? total : nan64();
    if (p.count)  p.count[o] = n;
    if (p.m2)     p.m2[o] = any ? fmax(m2, 0.0) : nan64();
    if (p.vmin)   p.vmin[o] = any ? lo : nan32();
    if (p.vmax)   p.vmax[o] = any ? hi : nan32();
    if (p.argmin) p.argmin[o] = any ? ilo : 0;
    if (p.argmax) p.argmax[o] = any ? ihi : 0;
}

template <int MODE>
__global__ void __launch_bounds__(128)
reduce_axis1_kernel(const __grid_constant__ SpRedParams p) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.nchan * p.nx) return;
    const int64_t c = g / p.nx, x = g - c * p.nx;
    const float *src = p.in + c * p.stride_c + x;
    SpAcc a;
    sp_init(a);
    for (int64_t y = 0; y < p.ny; ++y) {
        const float v = ldg_stream1(src + y * p.stride_y);
        if (mask_include<MODE>(p.mask, v, c, y, x) && v == v) sp_take(a, v, (int)y);
    }
    const double n = (double)a.n;
    sp_store(p, g, a.n, fma(n, a.k, a.s1), a.n > 0 ? a.s2 - a.s1 * a.s1 / n : 0.0, a.lo, a.hi, a.ilo, a.ihi);
}

template <int MODE>
__global__ void __launch_bounds__(128)
reduce_axis2_kernel(const __grid_constant__ SpRedParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= p.nchan * p.ny) return;                       // whole warps leave together
    const int64_t c = row / p.ny, y = row - c * p.ny;
    const float *src = p.in + c * p.stride_c + y * p.stride_y;
    SpAcc a;
    sp_init(a);
    for (int64_t x = lane; x < p.nx; x += 32) {
        const float v = ldg_stream1(src + x);
        if (mask_include<MODE>(p.mask, v, c, y, x) && v == v) sp_take(a, v, (int)x);
    }
    // this lane's (n, sum, M2, extrema); lanes that saw nothing carry n = 0 and are neutral in the merge
    int n = a.n;
    double total = fma((double)a.n, a.k, a.s1);
    double m2 = a.n > 0 ? a.s2 - a.s1 * a.s1 / (double)a.n : 0.0;
    float lo = a.lo, hi = a.hi;
    int ilo = a.n > 0 ? a.ilo : INT_MAX, ihi = a.n > 0 ? a.ihi : INT_MAX;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const int n_b = __shfl_xor_sync(0xffffffffu, n, off);
        const double total_b = __shfl_xor_sync(0xffffffffu, total, off);
        const double m2_b = __shfl_xor_sync(0xffffffffu, m2, off);
        const float lo_b = __shfl_xor_sync(0xffffffffu, lo, off), hi_b = __shfl_xor_sync(0xffffffffu, hi, off);
        const int ilo_b = __shfl_xor_sync(0xffffffffu, ilo, off), ihi_b = __shfl_xor_sync(0xffffffffu, ihi, off);
        if (n_b > 0) {
            if (n > 0) {
                // the same expression on both partners (symmetric in a <-> b), so every lane ends with the same bits
                const double na = (double)n, nb = (double)n_b;
                const double delta = total_b / nb - total / na;
                m2 = (m2 + m2_b) + delta * delta * (na * nb / (na + nb));
                total = total + total_b;
                if (lo_b < lo || (lo_b == lo && ilo_b < ilo)) { lo = lo_b; ilo = ilo_b; }
                if (hi_b > hi || (hi_b == hi && ihi_b < ihi)) { hi = hi_b; ihi = ihi_b; }
                n += n_b;
            } else {
                n = n_b; total = total_b; m2 = m2_b; lo = lo_b; hi = hi_b; ilo = ilo_b; ihi = ihi_b;
            }
        }
    }
    if (lane == 0) sp_store(p, row, n, total, m2, lo, hi, ilo, ihi);
}

}  // namespace scb

using namespace scb;

extern "C" int sc_reduce_spatial(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                                 int64_t stride_c, int64_t stride_y, int axis, const sc_mask_desc *mask,
                                 double *out_sum, int32_t *out_count, double *out_m2,
                                 float *out_min, float *out_max, int32_t *out_argmin, int32_t *out_argmax,
                                 void *stream) {
    int rc = check_cube_args(cube, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(axis == 1 || axis == 2, "axis must be 1 or 2");
    SC_CHECK_ARG(out_sum || out_count || out_m2 || out_min || out_max || out_argmin || out_argmax, "no output requested");
    SC_CHECK_ARG(ny < ((int64_t)1 << 31) && nx < ((int64_t)1 << 31), "axis too long for int32 indices");
    SpRedParams p{};
    p.in = cube; p.nchan = nchan; p.ny = ny; p.nx = nx; p.stride_c = stride_c; p.stride_y = stride_y;
    p.sum = out_sum; p.count = out_count; p.m2 = out_m2; p.vmin = out_min; p.vmax = out_max;
    p.argmin = out_argmin; p.argmax = out_argmax;
    rc = build_dev_mask(mask, cube, stride_c, stride_y, &p.mask);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int m = p.mask.mode;
    LaunchScope ls(SC_OP_REDUCE, s);
    if (axis == 1) {
        const int64_t grid = cdiv(nchan * nx, 128);
        SC_CHECK_ARG(grid < ((int64_t)1 << 31), "grid too large");
        if (m == MODE_NONE) reduce_axis1_kernel<MODE_NONE><<<(unsigned)grid, 128, 0, s>>>(p);
        else if (m == MODE_INTERVAL) reduce_axis1_kernel<MODE_INTERVAL><<<(unsigned)grid, 128, 0, s>>>(p);
        else reduce_axis1_kernel<MODE_GENERIC><<<(unsigned)grid, 128, 0, s>>>(p);
    } else {
        const int64_t grid = cdiv(nchan * ny, 4);            // 4 warps = 4 rows per CTA
        SC_CHECK_ARG(grid < ((int64_t)1 << 31), "grid too large");
        if (m == MODE_NONE) reduce_axis2_kernel<MODE_NONE><<<(unsigned)grid, 128, 0, s>>>(p);
        else if (m == MODE_INTERVAL) reduce_axis2_kernel<MODE_INTERVAL><<<(unsigned)grid, 128, 0, s>>>(p);
        else reduce_axis2_kernel<MODE_GENERIC><<<(unsigned)grid, 128, 0, s>>>(p);
    }
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}
