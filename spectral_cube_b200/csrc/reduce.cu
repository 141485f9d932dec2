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
#include "tma.cuh"

namespace scb {

int check_cube_args(const float *cube, int64_t nchan, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y);
int env_int(const char *name, int dflt);

constexpr int RD_UNROLL = 8;

struct ReduceParams {
    const float *in;
    int64_t nchan, ny, nx, stride_c, stride_y;
    double *sum, *m2;
    int32_t *count, *argmin, *argmax;
    float *vmin, *vmax;
    int tiles_per_row;         // TMA kernel only
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
        if (a.n == 0) a.k = fabsf(v) <= FLT_MAX ? (double)v : 0.0;   // an infinite first value is no shift: inf - inf would poison the sum (np.nansum gives +-inf)
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

// ---- TMA-pipelined variant (aligned cubes with enough tiles to fill the chip): the ring of moments.cu ----
// CTA = 512 adjacent spaxels of one image row; a producer warp streams slabs of RT_CB channels x 2 KB rows
// (`cp.async.bulk`, >= 2 KB requests run at the full copy rate) through an RT_STAGES-deep mbarrier ring;
// 4 consumer warps read float4 from shared memory.  Same accumulators as the direct kernel.
constexpr int RT_TILE_X = 512, RT_CONSUMERS = 128, RT_THREADS = RT_CONSUMERS + 32, RT_CB = 8, RT_STAGES = 4;

struct ReduceSmem {
    float data[RT_STAGES][RT_CB][RT_TILE_X];
    uint64_t full[RT_STAGES];
    uint64_t empty[RT_STAGES];
};

__device__ __forceinline__ void reduce_store(const ReduceParams &p, int64_t o, const ReduceAcc &a) {
    const bool any = a.n > 0;
    if (p.sum)    p.sum[o] = any ? fma((double)a.n, a.k, a.s1) : nan64();
    if (p.count)  p.count[o] = a.n;
    if (p.m2)     p.m2[o] = any ? fmax(a.s2 - a.s1 * a.s1 / (double)a.n, 0.0) : nan64();
    if (p.vmin)   p.vmin[o] = any ? a.lo : nan32();
    if (p.vmax)   p.vmax[o] = any ? a.hi : nan32();
    if (p.argmin) p.argmin[o] = a.ilo;
    if (p.argmax) p.argmax[o] = a.ihi;
}

template <int MODE>
__global__ void __launch_bounds__(RT_THREADS)
reduce_tma_kernel(const __grid_constant__ ReduceParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ReduceSmem &sm = *reinterpret_cast<ReduceSmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t tile = blockIdx.x;
    const int64_t y = tile / p.tiles_per_row;
    const int64_t x0 = (tile - y * p.tiles_per_row) * RT_TILE_X;
    const int width = (int)min((int64_t)RT_TILE_X, p.nx - x0);
    const int n_iter = (int)((p.nchan + RT_CB - 1) / RT_CB);
    if (tid == 0) {
        for (int s = 0; s < RT_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], RT_CONSUMERS / 32); }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == RT_CONSUMERS / 32) {
        const float *src = p.in + y * p.stride_y + x0;
        const uint64_t pol = l2_evict_first_policy();
        const uint32_t row_bytes = (uint32_t)width * 4u;
        for (int it = 0; it < n_iter; ++it) {
            const int s = it % RT_STAGES;
            const int64_t c0 = (int64_t)it * RT_CB;
            const int nch = (int)min((int64_t)RT_CB, p.nchan - c0);
            if (it >= RT_STAGES) mbar_wait(&sm.empty[s], ((it / RT_STAGES) - 1) & 1);
            if (lane == 0) mbar_expect_tx(&sm.full[s], (uint32_t)nch * row_bytes);
            __syncwarp();
            if (lane < nch) tma_load_1d(&sm.data[s][lane][0], src + (c0 + lane) * p.stride_c, row_bytes, &sm.full[s], pol);
        }
        return;
    }
    const int xo = tid * 4;
    const bool active = xo < width;
    ReduceAcc acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = ReduceAcc{0.0, 0.0, 0.0, 0.0f, 0.0f, 0, 0, 0};
    for (int it = 0; it < n_iter; ++it) {
        const int s = it % RT_STAGES;
        const int64_t c0 = (int64_t)it * RT_CB;
        const int nch = (int)min((int64_t)RT_CB, p.nchan - c0);
        mbar_wait(&sm.full[s], (it / RT_STAGES) & 1);
        if (active) {
            if (nch == RT_CB) {
#pragma unroll
                for (int cb = 0; cb < RT_CB; ++cb) {
                    const float4 v = *reinterpret_cast<const float4 *>(&sm.data[s][cb][xo]);
                    reduce_take<MODE>(p, acc[0], v.x, c0 + cb, y, x0 + xo + 0);
                    reduce_take<MODE>(p, acc[1], v.y, c0 + cb, y, x0 + xo + 1);
                    reduce_take<MODE>(p, acc[2], v.z, c0 + cb, y, x0 + xo + 2);
                    reduce_take<MODE>(p, acc[3], v.w, c0 + cb, y, x0 + xo + 3);
                }
            } else {
                for (int cb = 0; cb < nch; ++cb) {
                    const float4 v = *reinterpret_cast<const float4 *>(&sm.data[s][cb][xo]);
                    reduce_take<MODE>(p, acc[0], v.x, c0 + cb, y, x0 + xo + 0);
                    reduce_take<MODE>(p, acc[1], v.y, c0 + cb, y, x0 + xo + 1);
                    reduce_take<MODE>(p, acc[2], v.z, c0 + cb, y, x0 + xo + 2);
                    reduce_take<MODE>(p, acc[3], v.w, c0 + cb, y, x0 + xo + 3);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
    }
    if (!active) return;
    const int64_t o = y * p.nx + x0 + xo;
#pragma unroll
    for (int k = 0; k < 4; ++k) reduce_store(p, o + k, acc[k]);
}

template <int MODE>
static cudaError_t launch_reduce_tma(const ReduceParams &p, unsigned grid, cudaStream_t s) {
    auto kern = reduce_tma_kernel<MODE>;
    static unsigned long long configured = 0;        // per instantiation, one bit per device
    if (cudaError_t e = ensure_dyn_smem(kern, sizeof(ReduceSmem), &configured)) return e;
    kern<<<grid, RT_THREADS, sizeof(ReduceSmem), s>>>(p);
    return cudaGetLastError();
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
    const int m = p.mask.mode;
    const int64_t tiles_per_row = cdiv(nx, RT_TILE_X);
    const int64_t n_tiles = tiles_per_row * ny;
    const int choice = env_int("SC_REDUCE_KERNEL", 0);               // 0 auto, 1 direct, 2 tma
    if (vec4 && choice != 1 && n_tiles < ((int64_t)1 << 31) && choice == 2) {   // measured: 4.5 ms against 3.6 ms for the direct kernel -- the seven accumulators make the 12 consumer warps/SM of the ring issue-bound
        p.tiles_per_row = (int)tiles_per_row;
        LaunchScope ls(SC_OP_REDUCE, s);
        cudaError_t e;
        if (m == MODE_NONE) e = launch_reduce_tma<MODE_NONE>(p, (unsigned)n_tiles, s);
        else if (m == MODE_INTERVAL) e = launch_reduce_tma<MODE_INTERVAL>(p, (unsigned)n_tiles, s);
        else e = launch_reduce_tma<MODE_GENERIC>(p, (unsigned)n_tiles, s);
        if (e != cudaSuccess) return cuda_fail(e, "reduce_tma_kernel launch");
        return SC_OK;
    }
    const int64_t groups = vec4 ? (nx / 4) * ny : nx * ny;
    const unsigned grid = (unsigned)cdiv(groups, 128);
    LaunchScope ls(SC_OP_REDUCE, s);
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
