// spectral_interp.cu -- per-spaxel 1-D linear resampling of the spectral axis.
//
// Replaces the per-spaxel `np.interp(grid, inaxis, spectrum, left, right)` loop of the numpy
// class (spectral_cube.py:3298-3315) and the per-block `scipy.interpolate.interp1d` call of
// the dask class (dask_spectral_cube.py:1342-1349).
//
// The host turns the output grid into a sorted look-up table (one entry per output channel:
// bracketing input channel, knot coordinates, classification).  A thread owns one spaxel and
// streams its spectrum ONCE in ascending-axis order, keeping the previous sample in a register;
// every LUT entry is emitted when the highest input channel it needs arrives, so the inputs are
// read exactly once (4 B/voxel) whatever the grid, warps read/write 128 contiguous bytes per
// channel, and the per-spaxel "anything included?" flag the numpy class needs
// (spectral_cube.py:3299-3313) falls out of the same pass.  Arithmetic is float64 and follows
// numpy's / scipy's formulas term by term (see interp_numpy / interp_scipy below).
#include "common.cuh"
#include "tma.cuh"
#include <vector>
#include <algorithm>

namespace scb {

int check_cube_args(const float *cube, int64_t nchan, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y);
int env_int(const char *name, int dflt);

enum { IK_INTERIOR = 0, IK_KNOT = 1, IK_LEFT = 2, IK_RIGHT = 3 };

struct InterpEntry {          // 32 bytes: the two weights first (one aligned 16-byte load), then need / kind (one 8-byte load)
    double t_lo;              // (x - xlo) / (xhi - xlo): weight of the step from the lower knot
    double t_hi;              // (x - xhi) / (xhi - xlo): the same measured from the upper knot (<= 0)
    int32_t need;             // highest ascending-order input channel this entry needs
    int32_t kind;
    double pad;
};

struct InterpParams {
    const float *in;
    void *out;
    uint8_t *out_mask;
    int64_t nchan, ny, nx, stride_c, stride_y, nchan_out;
    const InterpEntry *lut;   // device, nchan_out entries sorted by x
    float fill;
    int has_fill_value;
    double fill_value;
    int in_reversed, out_reversed, mode;
    // scatter form (row-sharded cube -> channel shards): output channel j of THIS rank's rows is stored at
    // out_chan_ptrs[j] + y * nx + x -- a pointer into the buffer of the rank that owns channel j, local or a peer's
    // over NVLink -- instead of out + (j * ny + y) * nx + x
    const uint64_t *out_chan_ptrs;
    // scatter form, TMA kernel: a CTA's march over the spectrum starts at slab `slab_start[.][ph]` (slabs of 8 / of 4
    // channels: first index 0 / 1) with LUT entry `jj_start[.][ph]`, runs to the end of the spectrum and then from channel 0
    // up to the start again; ph = (phase + blockIdx.x) % nphases.  A job of N ranks passes phase = rank, nphases = N: at
    // any moment the CTAs of every rank store to ALL channel owners in equal shares.  (All ranks marching 0 -> nchan
    // together stored to the same owner at the same time: 425 GB/s per GPU; one phase per RANK: 559; per CTA: see DESIGN.)
    int phase, nphases;
    int slab_start[2][16], jj_start[2][16];
    DevMask mask;
};

template <typename T>
__device__ __forceinline__ T *interp_out_row(const InterpParams &p, int64_t j, int64_t plane_out) {
    if (p.out_chan_ptrs) return reinterpret_cast<T *>(__ldg(p.out_chan_ptrs + j));
    return reinterpret_cast<T *>(p.out) + j * plane_out;
}
template <typename T, bool SCATTER>
__device__ __forceinline__ T *interp_out_row_t(const InterpParams &p, int64_t j, int64_t plane_out) {
    if (SCATTER) return reinterpret_cast<T *>(__ldg(p.out_chan_ptrs + j));
    return reinterpret_cast<T *>(p.out) + j * plane_out;
}

// numpy/_core/src/multiarray/compiled_base.c (arr_interp): interior sample between two knots
// (numpy: slope = (f1-f0)/(xhi-xlo); r = slope*(x-xlo) + f0; if NaN: slope*(x-xhi) + f1; if still NaN and
// f0 == f1: f0).  The knot-spacing division is folded into the host-computed weights; the result
// differs from numpy's rounding sequence by at most an ulp of float64.)
__device__ __forceinline__ double interp_numpy(double t_lo, double t_hi, double f0, double f1) {
    const double d = f1 - f0;
    double r = fma(d, t_lo, f0);
    if (r != r) {
        r = fma(d, t_hi, f1);
        if (r != r && f0 == f1) r = f0;
    }
    return r;
}

// scipy/interpolate/_interpolate.py (interp1d._call_linear)
__device__ __forceinline__ double interp_scipy(double t_lo, double f0, double f1) {
    return fma(f1 - f0, t_lo, f0);
}

template <int MODE, int OUT64>
__global__ void __launch_bounds__(128)
spectral_interp_kernel(const __grid_constant__ InterpParams p) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.ny * p.nx) return;
    const int64_t y = g / p.nx, x = g - y * p.nx;
    const float *src = p.in + y * p.stride_y + x;
    const int64_t plane_out = p.ny * p.nx;
    const int64_t obase = y * p.nx + x;

    float prev = 0.0f, cur = 0.0f;          // filled values of ascending channels i-1, i
    bool mprev = false, mcur = false;       // their include flags
    bool any_included = false;
    int64_t jj = 0;
    for (int64_t i = 0; i < p.nchan; ++i) {
        const int64_t ch = p.in_reversed ? p.nchan - 1 - i : i;
        const float raw = ldg_stream1(src + ch * p.stride_c);
        prev = cur; mprev = mcur;
        mcur = mask_include<MODE>(p.mask, raw, ch, y, x);
        cur = mcur ? raw : p.fill;
        any_included |= mcur;
        while (jj < p.nchan_out) {
            const InterpEntry e = p.lut[jj];
            if (e.need != (int32_t)i) break;
            double r;
            bool m;
            if (p.mode == 0) {
                if (e.kind == IK_INTERIOR) { r = interp_numpy(e.t_lo, e.t_hi, (double)prev, (double)cur); m = mprev | mcur; }
                else if (e.kind == IK_KNOT) { r = (double)cur; m = mcur; }
                else { r = p.has_fill_value ? p.fill_value : (double)cur; m = mcur; }       // left / right of the axis
            } else {
                if (e.kind == IK_LEFT || e.kind == IK_RIGHT) r = p.has_fill_value ? p.fill_value : nan64();
                else r = interp_scipy(e.t_lo, (double)prev, (double)cur);
                m = r == r;
            }
            const int64_t jo = p.out_reversed ? p.nchan_out - 1 - jj : jj;
            // dask: the mask is taken BEFORE the output is flipped back (dask_spectral_cube.py:1364-1367)
            const int64_t jm = (p.mode == 1) ? jj : jo;
            if (OUT64) interp_out_row<double>(p, jo, plane_out)[obase] = r;
            else       interp_out_row<float>(p, jo, plane_out)[obase] = (float)r;
            if (p.out_mask) p.out_mask[jm * plane_out + obase] = m ? 1 : 0;
            ++jj;
        }
    }
    if (p.mode == 0 && !any_included && (p.fill == p.fill || p.has_fill_value)) {
        // nothing included in this spaxel: data NaN, mask False (spectral_cube.py:3311-3313); only
        // differs from what was emitted when a finite fill / fill_value was in play
        for (int64_t j = 0; j < p.nchan_out; ++j) {
            if (OUT64) interp_out_row<double>(p, j, plane_out)[obase] = nan64();
            else       interp_out_row<float>(p, j, plane_out)[obase] = nan32();
            if (p.out_mask) p.out_mask[j * plane_out + obase] = 0;
        }
    }
}

// LUT entry j through the read-only path (every thread reads the same sequence: L1-resident): the weights as one
// 16-byte load, {need, kind} as one 8-byte load; past the end {-1, 0} -- no channel ever matches
__device__ __forceinline__ double2 lut_weights(const InterpEntry *lut, int64_t j, int64_t n) {
    return __ldg(reinterpret_cast<const double2 *>(lut + (j < n ? j : n - 1)));
}
__device__ __forceinline__ int2 lut_need_kind(const InterpEntry *lut, int64_t j, int64_t n) {
    if (j >= n) return make_int2(-1, 0);
    return __ldg(reinterpret_cast<const int2 *>(reinterpret_cast<const char *>(lut + j) + 16));
}

// v * 2^-896 as a double, exactly, for finite v; NaN stays NaN and +-inf stays +-inf (see spectral_smooth.cu)
__device__ __forceinline__ double place_scaled_keep(float v) {
    const int b = __float_as_int(v);
    const long long t = (long long)b << 29;
    int hi = (int)(t >> 32) & 0x8FFFFFFF;
    if ((b & 0x7F800000) == 0x7F800000) hi |= 0x7FF00000;
    return __hiloint2double(hi, (int)t);
}

// ---- TMA-pipelined variant: same arithmetic, the spectrum arrives through a shared-memory ring ----------
// CTA = 512 adjacent spaxels of one image row (float4 per thread, 128 consumer threads + a producer warp);
// slabs of IT_CB channels x 2 KB rows stream through an IT_STAGES-deep ring like in moments.cu, in
// ascending-axis order (the producer reverses the channel index for descending axes).  Outputs are
// written as 16-byte vectors (data) and 4-byte vectors (mask).
constexpr int IT_TILE = 512, IT_CONS = 128, IT_THREADS = IT_CONS + 32;
// ring geometry (channels per slab, slabs in flight): the default and the variants SC_INTERP_RING selects for experiments

template <int IT_CB, int IT_STAGES>
struct InterpSmem {
    float data[IT_STAGES][IT_CB][IT_TILE];
    uint64_t full[IT_STAGES];
    uint64_t empty[IT_STAGES];
};

template <int MODE, int OUT64, int IT_CB, int IT_STAGES, bool SCATTER>
__global__ void __launch_bounds__(IT_THREADS, (IT_CB * IT_STAGES <= 20 ? 5 : 4))
spectral_interp_tma_kernel(const __grid_constant__ InterpParams p, int tiles_per_row) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    InterpSmem<IT_CB, IT_STAGES> &sm = *reinterpret_cast<InterpSmem<IT_CB, IT_STAGES> *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t tile = blockIdx.x;
    const int64_t y = tile / tiles_per_row;
    const int64_t x0 = (tile - y * tiles_per_row) * IT_TILE;
    const int width = (int)min((int64_t)IT_TILE, p.nx - x0);
    const int n_iter = (int)((p.nchan + IT_CB - 1) / IT_CB);
    // the march: slabs start .. n_iter - 1, then (scatter form with a phase) 0 .. start, the last one for its first channel only
    // (the plain form is compiled without any of this: start = 0 folds the second leg and the phase table away)
    const int ph = (SCATTER && p.nphases > 1) ? (int)((p.phase + blockIdx.x) % (unsigned)p.nphases) : 0;
    const int start = SCATTER ? p.slab_start[IT_CB == 8 ? 0 : 1][ph] : 0;
    const int n_first = n_iter - start, n_total = start > 0 ? n_iter + 1 : n_iter;
    if (tid == 0) {
        for (int s = 0; s < IT_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], IT_CONS / 32); }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == IT_CONS / 32) {
        const float *src = p.in + y * p.stride_y + x0;
        const uint64_t pol = l2_evict_first_policy();
        const uint32_t row_bytes = (uint32_t)width * 4u;
        for (int q = 0; q < n_total; ++q) {
            const int s = q % IT_STAGES;
            const int it = q < n_first ? start + q : q - n_first;
            const int64_t i0 = (int64_t)it * IT_CB;
            const int nch = (start > 0 && q == n_total - 1) ? 1 : (int)min((int64_t)IT_CB, p.nchan - i0);
            if (q >= IT_STAGES) mbar_wait(&sm.empty[s], ((q / IT_STAGES) - 1) & 1);
            if (lane == 0) mbar_expect_tx(&sm.full[s], (uint32_t)nch * row_bytes);
            __syncwarp();
            if (lane < nch) {
                const int64_t ch = p.in_reversed ? p.nchan - 1 - (i0 + lane) : i0 + lane;
                tma_load_1d(&sm.data[s][lane][0], src + ch * p.stride_c, row_bytes, &sm.full[s], pol);
            }
        }
        return;
    }
    const int xo = tid * 4;
    const bool active = xo < width;
    const int64_t x = x0 + xo;
    const int64_t plane_out = p.ny * p.nx;
    const int64_t obase = y * p.nx + x;
    // The two most recent filled samples of each spaxel live in two float32 register sets that swap roles with the
    // channel's parity (no register moves, the slab height is even); they are widened to float64 only when an output
    // needs them.  Instruction issue bounds this kernel, not the FP64 pipe (10 % busy): widening every input by bit
    // placement, as the smoothing kernels do, cost 7 instructions per voxel here for nothing (42 -> see DESIGN.md).
    float va[4] = {0, 0, 0, 0}, vb[4] = {0, 0, 0, 0};
    uint32_t inc_a = 0u, inc_b = 0u;                     // include bits (bit k = spaxel k) of the samples in va / vb
    uint32_t any_inc = 0u;
    int64_t jj = SCATTER ? p.jj_start[IT_CB == 8 ? 0 : 1][ph] : 0;
    // LUT entries ride in registers, requested TWO (weights) and THREE ({need, kind}) outputs before they are used: one
    // ahead left 28 % of the warp samples waiting on that load when an output is due every second channel.  The weights
    // of even / odd outputs live in two register pairs that are reloaded in place (a copy would wait for the load).
    double2 w_even, w_odd;
    int2 nk0, nk1, nk2;
    auto seek = [&](int64_t j) {                         // position the look-ahead registers on LUT entry j
        jj = j;
        w_even = lut_weights(p.lut, j + (j & 1), p.nchan_out);
        w_odd = lut_weights(p.lut, j + 1 - (j & 1), p.nchan_out);
        nk0 = lut_need_kind(p.lut, j, p.nchan_out); nk1 = lut_need_kind(p.lut, j + 1, p.nchan_out); nk2 = lut_need_kind(p.lut, j + 2, p.nchan_out);
    };
    seek(jj);

    auto take = [&](float (&dst)[4], uint32_t &inc, int s, int cb, int64_t ch) {
        const float4 v = *reinterpret_cast<const float4 *>(&sm.data[s][cb][xo]);
        const float raw[4] = {v.x, v.y, v.z, v.w};
        inc = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool m = mask_include<MODE>(p.mask, raw[k], ch, y, x + k);
            dst[k] = m ? raw[k] : p.fill;
            inc |= (m ? 1u : 0u) << k;
        }
        any_inc |= inc;
    };
    auto emit = [&](const float (&prev)[4], const float (&cur)[4], uint32_t mp, uint32_t mc, int64_t i) {
        while (nk0.x == (int32_t)i) {
            InterpEntry e;
            e.need = nk0.x; e.kind = nk0.y;
            if (jj & 1) { e.t_lo = w_odd.x; e.t_hi = w_odd.y; w_odd = lut_weights(p.lut, jj + 2, p.nchan_out); }
            else        { e.t_lo = w_even.x; e.t_hi = w_even.y; w_even = lut_weights(p.lut, jj + 2, p.nchan_out); }
            nk0 = nk1; nk1 = nk2; nk2 = lut_need_kind(p.lut, jj + 3, p.nchan_out);
            double r[4];
            uint32_t m4 = 0u;
            if (p.mode == 0) {
                if (e.kind == IK_INTERIOR) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) r[k] = interp_numpy(e.t_lo, e.t_hi, (double)prev[k], (double)cur[k]);
                    m4 = mp | mc;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) r[k] = (e.kind != IK_KNOT && p.has_fill_value) ? p.fill_value : (double)cur[k];
                    m4 = mc;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (e.kind == IK_LEFT || e.kind == IK_RIGHT) r[k] = p.has_fill_value ? p.fill_value : nan64();
                    else r[k] = interp_scipy(e.t_lo, (double)prev[k], (double)cur[k]);
                    m4 |= (r[k] == r[k] ? 1u : 0u) << k;
                }
            }
            const uint32_t mbits = (m4 * 0x00204081u) & 0x01010101u;     // bit k -> byte k
            const int64_t jo = p.out_reversed ? p.nchan_out - 1 - jj : jj;
            const int64_t jm = (p.mode == 1) ? jj : jo;
            if (active) {
                if (OUT64) {
                    double *o = interp_out_row_t<double, SCATTER>(p, jo, plane_out) + obase;
                    *reinterpret_cast<double2 *>(o) = make_double2(r[0], r[1]);
                    *reinterpret_cast<double2 *>(o + 2) = make_double2(r[2], r[3]);
                } else {
                    float *o = interp_out_row_t<float, SCATTER>(p, jo, plane_out) + obase;
                    *reinterpret_cast<float4 *>(o) = make_float4((float)r[0], (float)r[1], (float)r[2], (float)r[3]);
                }
                if (p.out_mask) *reinterpret_cast<uint32_t *>(p.out_mask + jm * plane_out + obase) = mbits;
            }
            ++jj;
        }
    };
    static_assert(IT_CB % 2 == 0, "the register sets swap roles with the channel parity");
    for (int q = 0; q < n_total; ++q) {
        const int s = q % IT_STAGES;
        const int it = q < n_first ? start + q : q - n_first;
        const int64_t i0 = (int64_t)it * IT_CB;
        const int nch = (start > 0 && q == n_total - 1) ? 1 : (int)min((int64_t)IT_CB, p.nchan - i0);
        if (start > 0 && q == n_first) seek(0);                          // second leg of the march: from channel 0 up to the start
        mbar_wait(&sm.full[s], (q / IT_STAGES) & 1);
        for (int cb = 0; cb < nch; cb += 2) {
            const int64_t i = i0 + cb;                                   // even ascending-order channel: into va, vb is the previous one
            take(va, inc_a, s, cb, p.in_reversed ? p.nchan - 1 - i : i);
            emit(vb, va, inc_b, inc_a, i);
            if (cb + 1 < nch) {
                take(vb, inc_b, s, cb + 1, p.in_reversed ? p.nchan - 2 - i : i + 1);
                emit(va, vb, inc_a, inc_b, i + 1);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
    }
    bool any_included[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) any_included[k] = (any_inc >> k) & 1u;
    if (p.mode == 0 && active && (p.fill == p.fill || p.has_fill_value)) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (any_included[k]) continue;
            for (int64_t j = 0; j < p.nchan_out; ++j) {
                if (OUT64) interp_out_row_t<double, SCATTER>(p, j, plane_out)[obase + k] = nan64();
                else       interp_out_row_t<float, SCATTER>(p, j, plane_out)[obase + k] = nan32();
                if (p.out_mask) p.out_mask[j * plane_out + obase + k] = 0;
            }
        }
    }
}

template <int MODE, int OUT64, int CB, int STAGES>
static cudaError_t launch_interp_tma_ring(const InterpParams &p, unsigned grid, int tiles_per_row, cudaStream_t s) {
    if (p.out_chan_ptrs) {
        auto kern = spectral_interp_tma_kernel<MODE, OUT64, CB, STAGES, true>;
        static unsigned long long configured = 0;    // per instantiation, one bit per device
        if (cudaError_t e = ensure_dyn_smem(kern, sizeof(InterpSmem<CB, STAGES>), &configured)) return e;
        kern<<<grid, IT_THREADS, sizeof(InterpSmem<CB, STAGES>), s>>>(p, tiles_per_row);
        return cudaGetLastError();
    }
    auto kern = spectral_interp_tma_kernel<MODE, OUT64, CB, STAGES, false>;
    static unsigned long long configured = 0;        // per instantiation, one bit per device
    if (cudaError_t e = ensure_dyn_smem(kern, sizeof(InterpSmem<CB, STAGES>), &configured)) return e;
    kern<<<grid, IT_THREADS, sizeof(InterpSmem<CB, STAGES>), s>>>(p, tiles_per_row);
    return cudaGetLastError();
}

template <int MODE, int OUT64>
static cudaError_t launch_interp_tma_one(const InterpParams &p, unsigned grid, int tiles_per_row, cudaStream_t s) {
    // the ring's geometry decides how many CTAs an SM holds: 8 channels x 3 slabs = 48 KB -> 4 CTAs (5.9 ms on a config-5
    // shard against 7.9 ms with 4 slabs = 3 CTAs); SC_INTERP_RING selects the other shapes for experiments
    switch (env_int("SC_INTERP_RING", 0)) {
        case 1:  return launch_interp_tma_ring<MODE, OUT64, 8, 4>(p, grid, tiles_per_row, s);     // 64 KB: 3 CTAs per SM (round 1)
        case 2:  return launch_interp_tma_ring<MODE, OUT64, 4, 6>(p, grid, tiles_per_row, s);     // finer hand-over, 48 KB
        case 3:  return launch_interp_tma_ring<MODE, OUT64, 4, 5>(p, grid, tiles_per_row, s);     // 40 KB, 80 registers: 5 CTAs per SM
        default: return launch_interp_tma_ring<MODE, OUT64, 8, 3>(p, grid, tiles_per_row, s);
    }
}

template <int OUT64>
static cudaError_t launch_interp_tma(const InterpParams &p, cudaStream_t s) {
    const int tiles_per_row = (int)cdiv(p.nx, IT_TILE);
    const unsigned grid = (unsigned)(tiles_per_row * p.ny);
    switch (p.mask.mode) {
        case MODE_NONE:     return launch_interp_tma_one<MODE_NONE, OUT64>(p, grid, tiles_per_row, s);
        case MODE_INTERVAL: return launch_interp_tma_one<MODE_INTERVAL, OUT64>(p, grid, tiles_per_row, s);
        default:            return launch_interp_tma_one<MODE_GENERIC, OUT64>(p, grid, tiles_per_row, s);
    }
}

template <int OUT64>
static cudaError_t launch_interp(const InterpParams &p, cudaStream_t s) {
    const unsigned grid = (unsigned)cdiv(p.ny * p.nx, 128);
    switch (p.mask.mode) {
        case MODE_NONE:     spectral_interp_kernel<MODE_NONE, OUT64><<<grid, 128, 0, s>>>(p); break;
        case MODE_INTERVAL: spectral_interp_kernel<MODE_INTERVAL, OUT64><<<grid, 128, 0, s>>>(p); break;
        default:            spectral_interp_kernel<MODE_GENERIC, OUT64><<<grid, 128, 0, s>>>(p); break;
    }
    return cudaGetLastError();
}

}  // namespace scb

using namespace scb;

static int spectral_interp_impl(const float *in, void *out, int out_dtype, uint8_t *out_mask,
                                int64_t nchan, int64_t ny, int64_t nx,
                                int64_t stride_c, int64_t stride_y,
                                int64_t nchan_out,
                                const sc_mask_desc *mask, double fill,
                                const double *in_axis, const double *grid,
                                int has_fill_value, double fill_value,
                                int in_reversed, int out_reversed, int mode,
                                const uint64_t *out_chan_ptrs, int phase, int nphases,
                                void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_cube_args(in, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(out != nullptr || out_chan_ptrs != nullptr, "out is NULL");
    SC_CHECK_ARG(out_dtype == SC_F32 || out_dtype == SC_F64, "out_dtype must be SC_F32 or SC_F64");
    SC_CHECK_ARG(nchan_out > 0 && nchan_out < ((int64_t)1 << 31) && nchan < ((int64_t)1 << 31), "bad channel counts");
    SC_CHECK_ARG(in_axis && grid, "in_axis / grid must not be NULL");
    SC_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 (numpy class) or 1 (dask class)");
    SC_CHECK_ARG(nchan >= 2, "need at least two input channels");
    for (int64_t i = 1; i < nchan; ++i)
        SC_CHECK_ARG(in_axis[i] > in_axis[i - 1], "in_axis must be strictly ascending (flip it and set in_reversed)");
    for (int64_t j = 1; j < nchan_out; ++j)
        SC_CHECK_ARG(grid[j] > grid[j - 1], "grid must be strictly ascending (flip it and set out_reversed)");
    const size_t need = (size_t)nchan_out * sizeof(InterpEntry) + 256;
    if (!workspace || workspace_bytes < need) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return SC_ERR_WORKSPACE;
    }
    // ---- host: classify every output sample (numpy: binary search with x == xp[j] and x == xp[-1]
    //      special cases; scipy: searchsorted(left) clipped to [1, n-1]) ----
    std::vector<InterpEntry> lut((size_t)nchan_out);
    const double x_first = in_axis[0], x_last = in_axis[nchan - 1];
    for (int64_t j = 0; j < nchan_out; ++j) {
        const double xv = grid[j];
        InterpEntry e{};
        double xlo = 0.0, xhi = 1.0;
        if (xv < x_first) { e.kind = IK_LEFT; e.need = 0; }
        else if (xv > x_last) { e.kind = IK_RIGHT; e.need = (int32_t)(nchan - 1); }
        else if (mode == 0) {
            // largest k with in_axis[k] <= xv
            const int64_t k = (std::upper_bound(in_axis, in_axis + nchan, xv) - in_axis) - 1;
            if (k == nchan - 1 || in_axis[k] == xv) { e.kind = IK_KNOT; e.need = (int32_t)k; }
            else { e.kind = IK_INTERIOR; e.need = (int32_t)(k + 1); xlo = in_axis[k]; xhi = in_axis[k + 1]; }
        } else {
            int64_t hi = std::lower_bound(in_axis, in_axis + nchan, xv) - in_axis;    // searchsorted(left)
            if (hi < 1) hi = 1;
            if (hi > nchan - 1) hi = nchan - 1;
            e.kind = IK_INTERIOR; e.need = (int32_t)hi; xlo = in_axis[hi - 1]; xhi = in_axis[hi];
        }
        e.t_lo = (xv - xlo) / (xhi - xlo);
        e.t_hi = (xv - xhi) / (xhi - xlo);
        lut[(size_t)j] = e;
    }
    // `need` must be non-decreasing for the single streaming pass (it is: grid and axis are sorted;
    // LEFT entries need channel 0 and precede everything, RIGHT entries need the last channel)
    cudaStream_t s = (cudaStream_t)stream;
    InterpEntry *lut_dev = (InterpEntry *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    SC_CUDA(cudaMemcpyAsync(lut_dev, lut.data(), (size_t)nchan_out * sizeof(InterpEntry), cudaMemcpyHostToDevice, s));
    InterpParams p{};
    p.in = in; p.out = out; p.out_mask = out_mask;
    p.nchan = nchan; p.ny = ny; p.nx = nx; p.stride_c = stride_c; p.stride_y = stride_y; p.nchan_out = nchan_out;
    p.lut = lut_dev; p.fill = (float)fill; p.has_fill_value = has_fill_value; p.fill_value = fill_value;
    p.in_reversed = in_reversed; p.out_reversed = out_reversed; p.mode = mode;
    p.out_chan_ptrs = out_chan_ptrs;
    p.phase = phase; p.nphases = nphases;
    for (int v = 0; v < 2; ++v) {
        const int64_t cb = v == 0 ? 8 : 4, n_iter = cdiv(nchan, cb);
        for (int ph = 0; ph < 16; ++ph) {
            const int64_t start = (nphases > 1 && ph < nphases) ? n_iter * ph / nphases : 0;
            // the first leg begins with a channel that serves only as the LOWER neighbour there: entries with
            // need <= start * cb are emitted by the second leg, which ends on that very channel
            int64_t j0 = 0;
            if (start > 0)
                j0 = std::upper_bound(lut.begin(), lut.end(), start * cb,
                                      [](int64_t c, const InterpEntry &e) { return c < (int64_t)e.need; }) - lut.begin();
            p.slab_start[v][ph] = (int)start; p.jj_start[v][ph] = (int)j0;
        }
    }
    rc = build_dev_mask(mask, in, stride_c, stride_y, &p.mask);
    if (rc) return rc;
    LaunchScope ls(SC_OP_SPECTRAL_INTERP, s);
    const bool aligned = ((uintptr_t)in % 16 == 0) && stride_c % 4 == 0 && stride_y % 4 == 0 && nx % 4 == 0 &&
                         (uintptr_t)out % 16 == 0 && (!out_mask || (uintptr_t)out_mask % 4 == 0);
    const int choice = env_int("SC_INTERP_KERNEL", 0);               // 0 auto, 1 direct, 2 tma
    cudaError_t e;
    if (aligned && choice != 1 && (choice == 2 || cdiv(nx, IT_TILE) * ny >= 148))
        e = out_dtype == SC_F64 ? launch_interp_tma<1>(p, s) : launch_interp_tma<0>(p, s);
    else
        e = out_dtype == SC_F64 ? launch_interp<1>(p, s) : launch_interp<0>(p, s);
    if (e != cudaSuccess) return cuda_fail(e, "spectral_interp_kernel launch");
    return SC_OK;
}

extern "C" int sc_spectral_interp(const float *in, void *out, int out_dtype, uint8_t *out_mask,
                                  int64_t nchan, int64_t ny, int64_t nx,
                                  int64_t stride_c, int64_t stride_y,
                                  int64_t nchan_out,
                                  const sc_mask_desc *mask, double fill,
                                  const double *in_axis, const double *grid,
                                  int has_fill_value, double fill_value,
                                  int in_reversed, int out_reversed, int mode,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    return spectral_interp_impl(in, out, out_dtype, out_mask, nchan, ny, nx, stride_c, stride_y, nchan_out, mask, fill,
                                in_axis, grid, has_fill_value, fill_value, in_reversed, out_reversed, mode, nullptr, 0, 1,
                                workspace, workspace_bytes, stream);
}

extern "C" int sc_spectral_interp_scatter(const float *in, const uint64_t *out_chan_ptrs, int out_dtype, uint8_t *out_mask,
                                          int64_t nchan, int64_t ny, int64_t nx,
                                          int64_t stride_c, int64_t stride_y,
                                          int64_t nchan_out,
                                          const sc_mask_desc *mask, double fill,
                                          const double *in_axis, const double *grid,
                                          int has_fill_value, double fill_value,
                                          int in_reversed, int out_reversed, int mode,
                                          int phase, int nphases,
                                          void *workspace, size_t workspace_bytes, void *stream) {
    SC_CHECK_ARG(out_chan_ptrs != nullptr, "out_chan_ptrs is NULL");
    SC_CHECK_ARG(nphases >= 1 && nphases <= 16 && phase >= 0 && phase < nphases, "phase must be in [0, nphases), nphases <= 16");
    return spectral_interp_impl(in, nullptr, out_dtype, out_mask, nchan, ny, nx, stride_c, stride_y, nchan_out, mask, fill,
                                in_axis, grid, has_fill_value, fill_value, in_reversed, out_reversed, mode, out_chan_ptrs,
                                phase, nphases, workspace, workspace_bytes, stream);
}
