// reduce.cu -- masked one-pass reductions along the spectral axis: sum, count, min, max, argmin,
// argmax and the sum of squared deviations, all from ONE read of the cube.
//
// Replaces `apply_numpy_function(np.nansum / nanmean / nanstd / nanmax / nanmin / nanargmax /
// nanargmin, fill=..., axis=0)` (spectral_cube.py:361-470 driver; `sum` :578-588, `mean` :592-652,
// `std` :669-724, `max` :770-781, `min` :785-796, `argmax` :800-811, `argmin` :815-826; dask class
// dask_spectral_cube.py:641-767): the reference fills excluded voxels with NaN (-inf / +inf for the
// arg reductions) and calls the numpy nan-function, i.e. a voxel takes part iff it is included by the
// mask AND is not NaN.  These are the reductions that build the noise map and the peak map behind a
// "> 3 sigma" mask (docs/examples.rst:61-93), SURVEY.md 8(f) item 2.
//
// Kernel: a thread owns 4 adjacent spaxels and walks the channels with 16-byte streaming loads
// (a warp reads 512 contiguous bytes per channel), RD_UNROLL channels in flight; sums in float64
// (sum of deviations from the spaxel's first included value and of their squares, so the variance
// does not cancel), extrema in float32 with first-occurrence indices like numpy's.  HBM-bound:
// 4 B/voxel in, at most 40 B/spaxel out.
#include "common.cuh"

namespace scb {

int check_cube_args(const float *cube, int64_t nchan, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y);

constexpr int RD_UNROLL = 8;

struct ReduceParams {
    const float *in;
    int64_t nchan, ny, nx, stride_c, stride_y;
    double *sum, *m2;
    int32_t *count, *argmin, *argmax;
    float *vmin, *vmax;
    DevMask mask;
};

struct ReduceAcc {
    double k, s1, s2;          // first included value; sums of (v - k) and (v - k)^2
    float lo, hi;
    int n, ilo, ihi;
};

template <int MODE>
__device__ __forceinline__ void reduce_take(const ReduceParams &p, ReduceAcc &a, float v, int64_t c, int64_t y, int64_t x) {
    const bool use = mask_include<MODE>(p.mask, v, c, y, x) && v == v;
    if (use) {
        if (a.n == 0) a.k = (double)v;
        const double d = (double)v - a.k;
        a.s1 += d;
        a.s2 = fma(d, d, a.s2);
        if (v < a.lo || a.n == 0) { a.lo = v; a.ilo = (int)c; }      // strict: first occurrence wins
        if (v > a.hi || a.n == 0) { a.hi = v; a.ihi = (int)c; }
        a.n += 1;
    }
}

template <int MODE, int VEC>
__global__ void __launch_bounds__(128)
reduce_axis0_kernel(const __grid_constant__ ReduceParams p) {
    const int64_t groups_per_row = (p.nx + VEC - 1) / VEC;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups_per_row * p.ny) return;
    const int64_t y = g / groups_per_row, x0 = (g - y * groups_per_row) * VEC;
    const float *src = p.in + y * p.stride_y + x0;
    ReduceAcc acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = ReduceAcc{0.0, 0.0, 0.0, 0.0f, 0.0f, 0, 0, 0};
    int64_t c = 0;
    for (; c + RD_UNROLL <= p.nchan; c += RD_UNROLL) {
        float v[RD_UNROLL][VEC];
#pragma unroll
        for (int u = 0; u < RD_UNROLL; ++u) VecLoad<VEC>::load(src + (c + u) * p.stride_c, v[u]);
#pragma unroll
        for (int u = 0; u < RD_UNROLL; ++u)
#pragma unroll
            for (int k = 0; k < VEC; ++k) reduce_take<MODE>(p, acc[k], v[u][k], c + u, y, x0 + k);
    }
    for (; c < p.nchan; ++c) {
        float v[VEC];
        VecLoad<VEC>::load(src + c * p.stride_c, v);
#pragma unroll
        for (int k = 0; k < VEC; ++k) reduce_take<MODE>(p, acc[k], v[k], c, y, x0 + k);
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        if (x0 + k >= p.nx) break;
        const int64_t o = y * p.nx + x0 + k;
        const ReduceAcc &a = acc[k];
        const bool any = a.n > 0;
        if (p.sum)    p.sum[o] = any ? fma((double)a.n, a.k, a.s1) : nan64();             // all excluded -> NaN (np_compat.py:20-24)
        if (p.count)  p.count[o] = a.n;
        // sum of squared deviations from the mean: s2 - s1^2 / n about the shift k
        if (p.m2)     p.m2[o] = any ? fmax(a.s2 - a.s1 * a.s1 / (double)a.n, 0.0) : nan64();
        if (p.vmin)   p.vmin[o] = any ? a.lo : nan32();
        if (p.vmax)   p.vmax[o] = any ? a.hi : nan32();
        if (p.argmin) p.argmin[o] = a.ilo;                                                 // 0 when nothing is included
        if (p.argmax) p.argmax[o] = a.ihi;
    }
}

}  // namespace scb

using namespace scb;

extern "C" int sc_reduce_axis0(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                               int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask,
                               double *out_sum, int32_t *out_count, double *out_m2,
                               float *out_min, float *out_max, int32_t *out_argmin, int32_t *out_argmax,
                               void *stream) {
    int rc = check_cube_args(cube, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(out_sum || out_count || out_m2 || out_min || out_max || out_argmin || out_argmax, "no output requested");
    SC_CHECK_ARG(nchan < ((int64_t)1 << 31), "too many channels for int32 indices");
    ReduceParams p{};
    p.in = cube; p.nchan = nchan; p.ny = ny; p.nx = nx; p.stride_c = stride_c; p.stride_y = stride_y;
    p.sum = out_sum; p.count = out_count; p.m2 = out_m2; p.vmin = out_min; p.vmax = out_max;
    p.argmin = out_argmin; p.argmax = out_argmax;
    rc = build_dev_mask(mask, cube, stride_c, stride_y, &p.mask);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec4 = ((uintptr_t)cube % 16 == 0) && stride_c % 4 == 0 && stride_y % 4 == 0 && nx % 4 == 0;
    const int64_t groups = vec4 ? (nx / 4) * ny : nx * ny;
    const unsigned grid = (unsigned)cdiv(groups, 128);
    LaunchScope ls(SC_OP_REDUCE, s);
    const int m = p.mask.mode;
    if (vec4) {
        if (m == MODE_NONE) reduce_axis0_kernel<MODE_NONE, 4><<<grid, 128, 0, s>>>(p);
        else if (m == MODE_INTERVAL) reduce_axis0_kernel<MODE_INTERVAL, 4><<<grid, 128, 0, s>>>(p);
        else reduce_axis0_kernel<MODE_GENERIC, 4><<<grid, 128, 0, s>>>(p);
    } else {
        if (m == MODE_NONE) reduce_axis0_kernel<MODE_NONE, 1><<<grid, 128, 0, s>>>(p);
        else if (m == MODE_INTERVAL) reduce_axis0_kernel<MODE_INTERVAL, 1><<<grid, 128, 0, s>>>(p);
        else reduce_axis0_kernel<MODE_GENERIC, 1><<<grid, 128, 0, s>>>(p);
    }
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}
