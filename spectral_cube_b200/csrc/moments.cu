// moments.cu -- fused masked moment0/1/2 along the spectral axis, one pass over the cube.
//
// Replaces spectral_cube/_moments.py:30-202 (cube/slice/ray strategies) and
// dask_spectral_cube.py:1083-1101.  HBM-bound streaming reduction.
//
// Main kernel (moments_tma_kernel): a CTA owns a tile of TILE_X adjacent spaxels of one image
// row.  A producer warp streams spectral slabs -- CB channels x TILE_X floats, one
// `cp.async.bulk` (TMA, SASS UBLKCP) row copy per channel, plus the matching slice of the
// per-channel coordinate table -- through a STAGES-deep shared-memory ring guarded by
// full/empty mbarriers; four consumer warps read their float4 from the slab (conflict-free
// LDS.128) and accumulate.  Loads in flight are set by the ring depth, not by what the
// compiler schedules, and every global access is a >= 2 KB contiguous burst.
// Whole spectra belong to one thread, so no cross-lane reduction is needed.
//
// Fallback kernel (moments_axis0_kernel): direct vector loads; used for unaligned views and
// for planes too small to fill 148 SMs, where the spectral axis is additionally split over
// blockDim.y and partial sums are combined through shared memory in a fixed order.
//
// Arithmetic (both): accumulators are float64 (NumPy >= 2 promotes the reference to
// float64, _moments.py:176; dask casts explicitly, dask_spectral_cube.py:1083).  Sums are
// taken about a uniform origin K = x[nchan/2]:  S0 = sum w, S1 = sum w d, S2 = sum w d^2 with
// d = x - K from a per-channel {d, d^2} table, so the per-voxel float64 work is one convert,
// one add and two FMAs, all predicated on the include test (no selects).
// M1 = K + S1/S0, M2 = S2/S0 - (S1/S0)^2.  The mask predicate is evaluated in registers
// (common.cuh), NaNs are skipped (nansum), rays with nothing included give NaN
// (np_compat.py:20-24).
#include "common.cuh"
#include "tma.cuh"

namespace scb {

struct MomParams {
    const float *cube;
    int64_t nchan, ny, nx;
    int64_t stride_c, stride_y;
    int64_t ngroups;          // ny * (nx / VEC) for VEC>1 (nx % VEC == 0), ny*nx for VEC=1
    int64_t groups_per_row;
    const double2 *tab;       // {d, d^2} per channel (device)
    double K, pix_size, m1_offset;
    double *m0, *m1, *m2;
    int tiles_per_row;        // TMA kernel only
    DevMask mask;
};

// One voxel: test the include predicate, then accumulate under that predicate (excluded and
// NaN voxels never touch the float64 results; no select instructions are needed).
template <int MODE, int WANT>
__device__ __forceinline__ void accumulate(const DevMask &m, float f, int64_t c, int64_t y, int64_t x,
                                           const double2 &t, double &s0, double &s1, double &s2, int &cnt, float other = 0.0f) {
    bool inc;
    if (MODE == MODE_INTERVAL_OTHER) inc = (other > m.lo) & (other < m.hi) & (f == f);   // the mask lives on another cube
    else {
        inc = mask_include<(MODE == MODE_INTERVAL_OTHER ? MODE_GENERIC : MODE)>(m, f, c, y, x);
        if (MODE != MODE_INTERVAL) inc = inc & (f == f);
    }
    if (inc) {
        const double w = (double)f;
        s0 += w;
        if (WANT & (SC_WANT_M1 | SC_WANT_M2)) s1 = fma(w, t.x, s1);
        if (WANT & SC_WANT_M2) s2 = fma(w, t.y, s2);
        cnt += 1;
    }
}

template <int WANT>
__device__ __forceinline__ void finalize(const MomParams &p, int64_t o, double s0, double s1, double s2, int cnt) {
    const bool any = cnt > 0;
    const double mean = s1 / s0;                        // offset from K; 0/0 -> NaN
    if ((WANT & SC_WANT_M0) && p.m0) p.m0[o] = any ? s0 * p.pix_size : nan64();
    if ((WANT & SC_WANT_M1) && p.m1) p.m1[o] = any ? (p.K + mean) + p.m1_offset : nan64();
    // a ray with a single included voxel has zero spread by construction (the reference gets
    // ~1e-26 from (x - M1)^2); the raw-sum form would leave +-1 ulp of d^2 here
    if ((WANT & SC_WANT_M2) && p.m2) p.m2[o] = any ? ((cnt == 1 && s0 != 0.0) ? 0.0 : s2 / s0 - mean * mean) : nan64();
}

// ---------------------------------------------------------------------------------------------
// TMA-pipelined kernel
// ---------------------------------------------------------------------------------------------
constexpr int TMA_TILE_X = 512;          // floats per tile row (2 KB bursts)
constexpr int TMA_CONSUMERS = 128;       // 4 consumer warps, float4 each
constexpr int TMA_THREADS = TMA_CONSUMERS + 32;

template <int CB, int STAGES>
struct MomSmem {
    float  data[STAGES][CB][TMA_TILE_X];     // 16-byte aligned rows
    double2 tab[STAGES][CB];
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
};

template <int CB, int STAGES, int MODE, int WANT>
__global__ void __launch_bounds__(TMA_THREADS)
moments_tma_kernel(const __grid_constant__ MomParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MomSmem<CB, STAGES> &sm = *reinterpret_cast<MomSmem<CB, STAGES> *>(smem_raw);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t tile = blockIdx.x;
    const int64_t y = tile / p.tiles_per_row;
    const int64_t x0 = (tile - y * p.tiles_per_row) * TMA_TILE_X;
    const int width = (int)min((int64_t)TMA_TILE_X, p.nx - x0);      // multiple of 4
    const int n_iter = (int)((p.nchan + CB - 1) / CB);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], TMA_CONSUMERS / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == TMA_CONSUMERS / 32) {
        // ---------------- producer warp ----------------
        const float *src = p.cube + y * p.stride_y + x0;
        const uint64_t pol = l2_evict_first_policy();
        const uint32_t row_bytes = (uint32_t)width * 4u;
        for (int it = 0; it < n_iter; ++it) {
            const int s = it % STAGES;
            const int64_t c0 = (int64_t)it * CB;
            const int nch = (int)min((int64_t)CB, p.nchan - c0);
            if (it >= STAGES) mbar_wait(&sm.empty[s], ((it / STAGES) - 1) & 1);
            if (lane == 0) {
                uint32_t bytes = (uint32_t)nch * row_bytes;
                if (WANT & (SC_WANT_M1 | SC_WANT_M2)) bytes += (uint32_t)nch * 16u;
                mbar_expect_tx(&sm.full[s], bytes);
            }
            __syncwarp();
            if (lane < nch)
                tma_load_1d(&sm.data[s][lane][0], src + (c0 + lane) * p.stride_c, row_bytes, &sm.full[s], pol);
            if ((WANT & (SC_WANT_M1 | SC_WANT_M2)) && lane == 31)
                tma_load_1d_nohint(&sm.tab[s][0], p.tab + c0, (uint32_t)nch * 16u, &sm.full[s]);
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const int xo = tid * 4;
    const bool active = xo < width;
    double s0[4], s1[4], s2[4];
    int cnt[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { s0[j] = 0.0; s1[j] = 0.0; s2[j] = 0.0; cnt[j] = 0; }

    for (int it = 0; it < n_iter; ++it) {
        const int s = it % STAGES;
        const int64_t c0 = (int64_t)it * CB;
        const int nch = (int)min((int64_t)CB, p.nchan - c0);
        mbar_wait(&sm.full[s], (it / STAGES) & 1);
        if (active) {
            if (nch == CB) {
#pragma unroll
                for (int cb = 0; cb < CB; ++cb) {
                    const float4 v = *reinterpret_cast<const float4 *>(&sm.data[s][cb][xo]);
                    double2 t = make_double2(0.0, 0.0);
                    if (WANT & (SC_WANT_M1 | SC_WANT_M2)) t = sm.tab[s][cb];
                    accumulate<MODE, WANT>(p.mask, v.x, c0 + cb, y, x0 + xo + 0, t, s0[0], s1[0], s2[0], cnt[0]);
                    accumulate<MODE, WANT>(p.mask, v.y, c0 + cb, y, x0 + xo + 1, t, s0[1], s1[1], s2[1], cnt[1]);
                    accumulate<MODE, WANT>(p.mask, v.z, c0 + cb, y, x0 + xo + 2, t, s0[2], s1[2], s2[2], cnt[2]);
                    accumulate<MODE, WANT>(p.mask, v.w, c0 + cb, y, x0 + xo + 3, t, s0[3], s1[3], s2[3], cnt[3]);
                }
            } else {
                for (int cb = 0; cb < nch; ++cb) {
                    const float4 v = *reinterpret_cast<const float4 *>(&sm.data[s][cb][xo]);
                    double2 t = make_double2(0.0, 0.0);
                    if (WANT & (SC_WANT_M1 | SC_WANT_M2)) t = sm.tab[s][cb];
                    accumulate<MODE, WANT>(p.mask, v.x, c0 + cb, y, x0 + xo + 0, t, s0[0], s1[0], s2[0], cnt[0]);
                    accumulate<MODE, WANT>(p.mask, v.y, c0 + cb, y, x0 + xo + 1, t, s0[1], s1[1], s2[1], cnt[1]);
                    accumulate<MODE, WANT>(p.mask, v.z, c0 + cb, y, x0 + xo + 2, t, s0[2], s1[2], s2[2], cnt[2]);
                    accumulate<MODE, WANT>(p.mask, v.w, c0 + cb, y, x0 + xo + 3, t, s0[3], s1[3], s2[3], cnt[3]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
    }
    if (!active) return;
    const int64_t o = y * p.nx + x0 + xo;
#pragma unroll
    for (int j = 0; j < 4; ++j) finalize<WANT>(p, o + j, s0[j], s1[j], s2[j], cnt[j]);
}

// ---------------------------------------------------------------------------------------------
// direct-load kernel (fallback / small planes)
// ---------------------------------------------------------------------------------------------
template <int VEC, int UNROLL, int MODE, int WANT>
__global__ void __launch_bounds__(256)
moments_axis0_kernel(const __grid_constant__ MomParams p) {
    const int tx = threadIdx.x;
    const int sy = threadIdx.y;                 // spectral split index
    const int nsplit = blockDim.y;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + tx;
    const bool active = g < p.ngroups;

    double s0[VEC], s1[VEC], s2[VEC];
    int cnt[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { s0[j] = 0.0; s1[j] = 0.0; s2[j] = 0.0; cnt[j] = 0; }

    int64_t y = 0, x = 0;
    if (active) {
        y = g / p.groups_per_row;
        x = (g - y * p.groups_per_row) * VEC;
        const int64_t chunk = (p.nchan + nsplit - 1) / nsplit;
        const int64_t c_begin = (int64_t)sy * chunk;
        const int64_t c_end = min(p.nchan, c_begin + chunk);
        const float *ptr = p.cube + y * p.stride_y + x + c_begin * p.stride_c;
        constexpr bool OTHER = MODE == MODE_INTERVAL_OTHER;     // a second stream: the cube the mask is tested on
        const float *optr = OTHER ? p.mask.other + y * p.mask.other_sy + x + c_begin * p.mask.other_sc : ptr;
        const int64_t ostep = OTHER ? p.mask.other_sc : 0;

        int64_t c = c_begin;
        for (; c + UNROLL <= c_end; c += UNROLL) {
            float v[UNROLL][VEC], ov[OTHER ? UNROLL : 1][VEC];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                VecLoad<VEC>::load(ptr + (int64_t)u * p.stride_c, v[u]);
                if (OTHER) VecLoad<VEC>::load(optr + (int64_t)u * ostep, ov[OTHER ? u : 0]);
            }
            ptr += (int64_t)UNROLL * p.stride_c;
            optr += (int64_t)UNROLL * ostep;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                double2 t = make_double2(0.0, 0.0);
                if (WANT & (SC_WANT_M1 | SC_WANT_M2)) t = __ldg(p.tab + c + u);
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    accumulate<MODE, WANT>(p.mask, v[u][j], c + u, y, x + j, t, s0[j], s1[j], s2[j], cnt[j], OTHER ? ov[OTHER ? u : 0][j] : 0.0f);
            }
        }
        for (; c < c_end; ++c) {
            float v[VEC], ov[VEC];
            VecLoad<VEC>::load(ptr, v);
            if (OTHER) VecLoad<VEC>::load(optr, ov);
            ptr += p.stride_c;
            optr += ostep;
            double2 t = make_double2(0.0, 0.0);
            if (WANT & (SC_WANT_M1 | SC_WANT_M2)) t = __ldg(p.tab + c);
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                accumulate<MODE, WANT>(p.mask, v[j], c, y, x + j, t, s0[j], s1[j], s2[j], cnt[j], OTHER ? ov[j] : 0.0f);
        }
    }

    if (nsplit > 1) {
        // combine spectral splits in a fixed order: split 0 adds splits 1..n-1
        extern __shared__ double smem[];
        const int bx = blockDim.x;
        double *sh = smem;                                   // [nsplit-1][4][VEC][bx]
        if (sy > 0) {
            double *base = sh + (size_t)(sy - 1) * 4 * VEC * bx;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                base[(0 * VEC + j) * bx + tx] = s0[j];
                base[(1 * VEC + j) * bx + tx] = s1[j];
                base[(2 * VEC + j) * bx + tx] = s2[j];
                base[(3 * VEC + j) * bx + tx] = (double)cnt[j];
            }
        }
        __syncthreads();
        if (sy > 0) return;
        for (int s = 1; s < nsplit; ++s) {
            const double *base = sh + (size_t)(s - 1) * 4 * VEC * bx;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                s0[j] += base[(0 * VEC + j) * bx + tx];
                s1[j] += base[(1 * VEC + j) * bx + tx];
                s2[j] += base[(2 * VEC + j) * bx + tx];
                cnt[j] += (int)base[(3 * VEC + j) * bx + tx];
            }
        }
    }
    if (!active) return;
    const int64_t o = y * p.nx + x;
#pragma unroll
    for (int j = 0; j < VEC; ++j) finalize<WANT>(p, o + j, s0[j], s1[j], s2[j], cnt[j]);
}

__global__ void moment_table_kernel(const double *__restrict__ x, int64_t n, double K, double2 *tab) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { double d = x[i] - K; tab[i] = make_double2(d, d * d); }
}

// ---- central moments of arbitrary order about a given plane (second pass of the reference) ----
struct CentralParams {
    const float *cube;
    int64_t nchan, ny, nx, stride_c, stride_y;
    const double *x;          // per-channel offset from channel 0 (device)
    const double *centre;     // (ny, nx)
    int order;
    double *out;
    DevMask mask;
};

template <int MODE>
__global__ void __launch_bounds__(256)
moment_central_kernel(const __grid_constant__ CentralParams p) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.ny * p.nx) return;
    const int64_t y = g / p.nx, x = g - y * p.nx;
    const float *ptr = p.cube + y * p.stride_y + x;
    const double m = p.centre[g];
    double num = 0.0, den = 0.0;
    int cnt = 0;
    for (int64_t c = 0; c < p.nchan; ++c) {
        const float f = ldg_stream1(ptr + c * p.stride_c);
        bool inc = mask_include<MODE>(p.mask, f, c, y, x);
        if (MODE != MODE_INTERVAL) inc = inc & (f == f);
        if (inc) {
            const double w = (double)f;
            const double d = __ldg(p.x + c) - m;
            double pw = d;
            for (int k = 1; k < p.order; ++k) pw *= d;
            num = fma(w, pw, num);
            den += w;
            cnt += 1;
        }
    }
    p.out[g] = cnt > 0 ? num / den : nan64();
}

// ---- launch plumbing ------------------------------------------------------------------------
template <int VEC, int UNROLL, int MODE>
static cudaError_t launch_moments_want(const MomParams &p, int want, dim3 grid, dim3 block, size_t smem, cudaStream_t s) {
    const int hi = (want & SC_WANT_M2) ? 2 : (want & SC_WANT_M1) ? 1 : 0;
    // M1-only requests run the M0|M1|M2 instantiation: ptxas turns the two-op predicated body of
    // an M0|M1 variant into selects + spills, and it measured slower than the three-op one.
    if (hi >= 1) moments_axis0_kernel<VEC, UNROLL, MODE, 7><<<grid, block, smem, s>>>(p);
    else         moments_axis0_kernel<VEC, UNROLL, MODE, 1><<<grid, block, smem, s>>>(p);
    return cudaGetLastError();
}

template <int VEC, int UNROLL>
static cudaError_t launch_moments_mode(const MomParams &p, int want, dim3 grid, dim3 block, size_t smem, cudaStream_t s) {
    switch (p.mask.mode) {
        case MODE_NONE:     return launch_moments_want<VEC, UNROLL, MODE_NONE>(p, want, grid, block, smem, s);
        case MODE_INTERVAL: return launch_moments_want<VEC, UNROLL, MODE_INTERVAL>(p, want, grid, block, smem, s);
        case MODE_INTERVAL_OTHER: return launch_moments_want<VEC, UNROLL, MODE_INTERVAL_OTHER>(p, want, grid, block, smem, s);
        default:            return launch_moments_want<VEC, UNROLL, MODE_GENERIC>(p, want, grid, block, smem, s);
    }
}

template <int CB, int STAGES, int MODE, int WANT>
static cudaError_t launch_tma_one(const MomParams &p, unsigned grid, cudaStream_t s) {
    auto kern = moments_tma_kernel<CB, STAGES, MODE, WANT>;
    const size_t smem = sizeof(MomSmem<CB, STAGES>);
    static unsigned long long configured = 0;        // per instantiation, one bit per device
    if (cudaError_t e = ensure_dyn_smem(kern, smem, &configured)) return e;
    kern<<<grid, TMA_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

template <int CB, int STAGES, int MODE>
static cudaError_t launch_tma_want(const MomParams &p, int want, unsigned grid, cudaStream_t s) {
    const int hi = (want & SC_WANT_M2) ? 2 : (want & SC_WANT_M1) ? 1 : 0;
    if (hi >= 1) return launch_tma_one<CB, STAGES, MODE, 7>(p, grid, s);   // see launch_moments_want
    return launch_tma_one<CB, STAGES, MODE, 1>(p, grid, s);
}

template <int CB, int STAGES>
static cudaError_t launch_tma_mode(const MomParams &p, int want, unsigned grid, cudaStream_t s) {
    switch (p.mask.mode) {
        case MODE_NONE:     return launch_tma_want<CB, STAGES, MODE_NONE>(p, want, grid, s);
        case MODE_INTERVAL: return launch_tma_want<CB, STAGES, MODE_INTERVAL>(p, want, grid, s);
        default:            return launch_tma_want<CB, STAGES, MODE_GENERIC>(p, want, grid, s);
    }
}

// tuning knobs (overridable for experiments through the environment)
int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

int moments_axis0_device(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                         int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask,
                         const double *chan_offset_dev, double K,
                         double pix_size, double m1_offset, int want_bits,
                         double *out_m0, double *out_m1, double *out_m2,
                         double2 *tab_dev, cudaStream_t s) {
    MomParams p;
    p.cube = cube; p.nchan = nchan; p.ny = ny; p.nx = nx;
    p.stride_c = stride_c; p.stride_y = stride_y;
    p.tab = tab_dev; p.K = K; p.pix_size = pix_size; p.m1_offset = m1_offset;
    p.m0 = out_m0; p.m1 = out_m1; p.m2 = out_m2;
    p.tiles_per_row = 0; p.groups_per_row = 0; p.ngroups = 0;
    int rc = build_dev_mask(mask, cube, stride_c, stride_y, &p.mask, true);
    if (rc != SC_OK) return rc;

    if (want_bits & (SC_WANT_M1 | SC_WANT_M2)) {
        LaunchScope ls(0, s);
        moment_table_kernel<<<(unsigned)cdiv(nchan, 256), 256, 0, s>>>(chan_offset_dev, nchan, K, tab_dev);
        SC_CUDA(cudaGetLastError());
    }

    const bool al16 = ((uintptr_t)cube % 16 == 0) && stride_c % 4 == 0 && stride_y % 4 == 0 && nx % 4 == 0;
    const bool al8 = ((uintptr_t)cube % 8 == 0) && stride_c % 2 == 0 && stride_y % 2 == 0 && nx % 2 == 0;

    // ---- TMA-pipelined path: aligned cubes with enough tiles to fill the chip ----
    const int64_t tiles_per_row = cdiv(nx, TMA_TILE_X);
    const int64_t n_tiles = tiles_per_row * ny;
    const int kernel_choice = env_int("SC_MOM_KERNEL", 0);      // 0 auto, 1 direct, 2 tma
    // a mask on another cube streams that cube too: the direct kernel reads both (the TMA ring carries one cube)
    const bool other = p.mask.mode == MODE_INTERVAL_OTHER;
    const bool oal16 = !other || (((uintptr_t)p.mask.other % 16 == 0) && p.mask.other_sc % 4 == 0 && p.mask.other_sy % 4 == 0);
    const bool oal8 = !other || (((uintptr_t)p.mask.other % 8 == 0) && p.mask.other_sc % 2 == 0 && p.mask.other_sy % 2 == 0);
    const bool tma_ok = al16 && !other && n_tiles < (int64_t)1 << 31;
    if (tma_ok && kernel_choice != 1 && (kernel_choice == 2 || n_tiles >= 2 * 148 * 3)) {
        p.tiles_per_row = (int)tiles_per_row;
        const int cfg = env_int("SC_MOM_TMA_CFG", 0);
        LaunchScope ls(SC_OP_MOMENTS, s);
        cudaError_t e;
        switch (cfg) {
            case 1:  e = launch_tma_mode<4, 8>(p, want_bits, (unsigned)n_tiles, s); break;
            case 2:  e = launch_tma_mode<8, 6>(p, want_bits, (unsigned)n_tiles, s); break;
            case 3:  e = launch_tma_mode<16, 3>(p, want_bits, (unsigned)n_tiles, s); break;
            default: e = launch_tma_mode<8, 4>(p, want_bits, (unsigned)n_tiles, s); break;
        }
        if (e != cudaSuccess) return cuda_fail(e, "moments_tma_kernel launch");
        return SC_OK;
    }

    // ---- direct-load path ----
    int vec = 1;
    if (al16 && oal16) vec = 4; else if (al8 && oal8) vec = 2;
    int force_vec = env_int("SC_MOM_VEC", 0);
    if (force_vec == 1 || (force_vec == 2 && al8)) vec = force_vec;
    p.groups_per_row = nx / vec;
    p.ngroups = ny * p.groups_per_row;

    // spectral split: keep >= ~2 resident thread-waves of work on 148 SMs
    const int bx = env_int("SC_MOM_BX", 128);
    int nsplit = 1;
    const int64_t want_threads = (int64_t)148 * 2048 * 2;
    while (nsplit < 8 && p.ngroups * nsplit < want_threads && nchan / (nsplit * 2) >= 32 && bx * nsplit * 2 <= 256) nsplit *= 2;
    int force_split = env_int("SC_MOM_SPLIT", 0);
    if (force_split > 0 && bx * force_split <= 256) nsplit = force_split;
    const int unroll = env_int("SC_MOM_UNROLL", 4);

    dim3 block(bx, nsplit);
    dim3 grid((unsigned)cdiv(p.ngroups, bx));
    size_t smem = nsplit > 1 ? (size_t)(nsplit - 1) * 4 * vec * bx * sizeof(double) : 0;
    LaunchScope ls(SC_OP_MOMENTS, s);
    cudaError_t e;
    if (vec == 4)      e = unroll >= 8 ? launch_moments_mode<4, 8>(p, want_bits, grid, block, smem, s)
                                       : launch_moments_mode<4, 4>(p, want_bits, grid, block, smem, s);
    else if (vec == 2) e = launch_moments_mode<2, 4>(p, want_bits, grid, block, smem, s);
    else               e = launch_moments_mode<1, 8>(p, want_bits, grid, block, smem, s);
    if (e != cudaSuccess) return cuda_fail(e, "moments_axis0_kernel launch");
    return SC_OK;
}

int check_cube_args(const float *cube, int64_t nchan, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y) {
    SC_CHECK_ARG(cube != nullptr, "cube pointer is NULL");
    SC_CHECK_ARG(nchan > 0 && ny > 0 && nx > 0, "cube shape (%lld, %lld, %lld) must be positive", (long long)nchan, (long long)ny, (long long)nx);
    SC_CHECK_ARG(stride_y >= nx, "stride_y=%lld smaller than nx=%lld", (long long)stride_y, (long long)nx);
    SC_CHECK_ARG(stride_c >= stride_y, "stride_c=%lld smaller than stride_y=%lld (spectral axis must be the slowest)", (long long)stride_c, (long long)stride_y);
    return SC_OK;
}

}  // namespace scb

using namespace scb;

extern "C" int sc_moments_axis0(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                                int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask,
                                const double *chan_offset, double pix_size, double m1_offset,
                                int want_bits, double *out_m0, double *out_m1, double *out_m2,
                                void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_cube_args(cube, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(want_bits > 0 && want_bits <= 7, "want_bits=%d must be a combination of 1|2|4", want_bits);
    SC_CHECK_ARG(!(want_bits & SC_WANT_M0) || out_m0, "out_m0 is NULL but moment 0 was requested");
    SC_CHECK_ARG(!(want_bits & SC_WANT_M1) || out_m1, "out_m1 is NULL but moment 1 was requested");
    SC_CHECK_ARG(!(want_bits & SC_WANT_M2) || out_m2, "out_m2 is NULL but moment 2 was requested");
    cudaStream_t s = (cudaStream_t)stream;
    double K = 0.0;
    double *xdev = nullptr;
    double2 *tab = nullptr;
    if (want_bits & (SC_WANT_M1 | SC_WANT_M2)) {
        SC_CHECK_ARG(chan_offset != nullptr, "chan_offset is NULL but moment 1/2 was requested");
        const size_t need = sc_workspace_bytes(SC_OP_MOMENTS, nchan, ny, nx, 0);
        if (!workspace || workspace_bytes < need) {
            set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
            return SC_ERR_WORKSPACE;
        }
        K = chan_offset[nchan / 2];
        tab = (double2 *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
        xdev = (double *)(tab + nchan);
        SC_CUDA(cudaMemcpyAsync(xdev, chan_offset, (size_t)nchan * 8, cudaMemcpyHostToDevice, s));
    }
    return moments_axis0_device(cube, nchan, ny, nx, stride_c, stride_y, mask, xdev, K, pix_size,
                                m1_offset, want_bits, out_m0, out_m1, out_m2, tab, s);
}

extern "C" int sc_moment_central_axis0(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                                       int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask,
                                       const double *chan_offset, const double *centre, int order,
                                       double *out, void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_cube_args(cube, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(chan_offset && centre && out, "chan_offset/centre/out must not be NULL");
    SC_CHECK_ARG(order >= 1 && order <= 64, "order=%d out of range", order);
    const size_t need = (size_t)nchan * 8 + 256;
    if (!workspace || workspace_bytes < need) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return SC_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    double *xdev = (double *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    SC_CUDA(cudaMemcpyAsync(xdev, chan_offset, (size_t)nchan * 8, cudaMemcpyHostToDevice, s));
    CentralParams p;
    p.cube = cube; p.nchan = nchan; p.ny = ny; p.nx = nx; p.stride_c = stride_c; p.stride_y = stride_y;
    p.x = xdev; p.centre = centre; p.order = order; p.out = out;
    rc = build_dev_mask(mask, cube, stride_c, stride_y, &p.mask);
    if (rc) return rc;
    dim3 block(128), grid((unsigned)cdiv(ny * nx, 128));
    LaunchScope ls(SC_OP_MOMENTS, s);
    if (p.mask.mode == MODE_NONE)          moment_central_kernel<MODE_NONE><<<grid, block, 0, s>>>(p);
    else if (p.mask.mode == MODE_INTERVAL) moment_central_kernel<MODE_INTERVAL><<<grid, block, 0, s>>>(p);
    else                                   moment_central_kernel<MODE_GENERIC><<<grid, block, 0, s>>>(p);
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}
