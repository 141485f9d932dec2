// spectral_smooth.cu -- per-spaxel NaN-interpolating convolution along the spectral axis, and
// the fused spectral_smooth -> moment0/1/2 chain (BASELINE config 3).
//
// Replaces `convolve(spectrum, kernel, normalize_kernel=True)` applied to every spaxel
// (spectral_cube.py:3186-3222 through :3103-3159 and :147-158; dask_spectral_cube.py:880-917).
// astropy semantics: true convolution (kernel flipped), zero-filled boundary whose zeros are
// VALID samples, out = sum_k K[k] v[c-k] [v not NaN] / sum_k K[k] [v not NaN] in float64,
// denominator 0 keeps the input NaN.  Masked voxels are replaced by `fill` before convolving.
//
// Kernel (smooth_tma_kernel): a CTA owns SM_TILE adjacent spaxels of one image row (one per
// consumer thread).  A producer warp streams blocks of B channels x SM_TILE floats through a
// shared-memory ring with `cp.async.bulk` (TMA) + mbarriers.  For every block of B output
// channels a consumer thread pulls the B + 2H inputs it needs from three neighbouring ring
// stages into registers as float64 (mask/fill applied, NaN -> 0 with a bit recorded), then
// evaluates B independent n-tap FMA chains with compile-time register indices.  The
// denominator is the kernel sum unless the window's NaN bits are set (rare), in which case the
// missing taps are subtracted one by one.  Outputs are stored straight from registers
// (a warp writes 128 contiguous bytes per channel) or, in the fused variant, rounded to the
// smoothed cube's dtype and accumulated into the moment sums -- the smoothed cube is never
// written.  The float64 pipe is the co-limiter here: n FMAs per voxel against 8 B of traffic.
#include "common.cuh"
#include "tma.cuh"

namespace scb {

int check_cube_args(const float *cube, int64_t nchan, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y);
int env_int(const char *name, int dflt);

constexpr int SM_TILE = 128;                 // spaxels per CTA = consumer threads
constexpr int SM_THREADS = SM_TILE + 32;
constexpr int SM_MAX_TAPS = 2 * 16 + 1;

struct SmoothParams {
    const float *in;
    void *out;                               // float32 or float64 (OUT64)
    int64_t nchan, ny, nx;
    int64_t stride_c, stride_y, out_stride_c, out_stride_y;
    int tiles_per_row;
    int ntaps;                               // actual number of taps (<= 2H+1)
    float fill;
    double ksum;                             // sum of the normalised taps (~1)
    double taps[SM_MAX_TAPS];                // normalised, centred in a 2H+1 window, zero padded
    int passthrough_spaxels;                 // numpy class: spaxels with nothing included are copied through
    // fused moments
    const double2 *tab;                      // {d, d^2} per channel
    double K, pix_size, m1_offset;
    double *m0, *m1, *m2;
    int round_f32;                           // smoothed cube dtype is float32 (dask class)
    DevMask mask;
};

template <int B, int STAGES>
struct SmoothSmem {
    float data[STAGES][B][SM_TILE];
    double taps[SM_MAX_TAPS];
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
};

// float32 -> float64 through an opaque conversion so the NaN select stays on the float32 side
__device__ __forceinline__ double cvt_f64(float v) {
    double d;
    asm("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(v));
    return d;
}

// Rare path: some inputs of this output's window are NaN.  `wb` has bit j set when window entry j
// (input o + j, tap 2H - j) is NaN.  Subtract the missing taps from the full kernel sum; if every
// real tap is missing the denominator is exactly 0.
template <int H>
__device__ __noinline__ double smooth_bot_with_nans(const double *taps_sm, uint64_t wb, double ksum, int ntaps) {
    constexpr uint64_t FULL = (1ull << (2 * H + 1)) - 1ull;
    if (wb == FULL) return 0.0;                              // whole window missing (blank spectra)
    const int h = ntaps >> 1;
    if (__popcll(wb) > H) {
        // mostly missing: add up what is there (taps outside the real kernel are zero)
        double bot = 0.0;
        uint64_t good = ~wb & FULL;
        while (good) {
            const int j = __ffsll((long long)good) - 1;
            good &= good - 1;
            bot += taps_sm[2 * H - j];
        }
        return bot;
    }
    double bot = ksum;
    int nbad = 0;
    while (wb) {
        const int j = __ffsll((long long)wb) - 1;
        wb &= wb - 1;
        const int k = 2 * H - j;
        bot -= taps_sm[k];
        nbad += (k >= H - h && k <= H + h) ? 1 : 0;
    }
    return nbad >= ntaps ? 0.0 : bot;
}

// EPI 0: store float32, 1: store float64, 2: fused moments
template <int H, int B, int STAGES, int MODE, int EPI>
__global__ void __launch_bounds__(SM_THREADS)
smooth_tma_kernel(const __grid_constant__ SmoothParams p) {
    constexpr int NT = 2 * H + 1;
    constexpr int NIN = B + 2 * H;
    static_assert(H <= B, "window must fit in neighbouring stages");
    static_assert(NIN <= 64, "NaN bit mask is 64 bits");
    static_assert(STAGES >= 4, "three stages are read while one is being filled");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmoothSmem<B, STAGES> &sm = *reinterpret_cast<SmoothSmem<B, STAGES> *>(smem_raw);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t tile = blockIdx.x;
    const int64_t y = tile / p.tiles_per_row;
    const int64_t x0 = (tile - y * p.tiles_per_row) * SM_TILE;
    const int width = (int)min((int64_t)SM_TILE, p.nx - x0);
    const int nblk = (int)((p.nchan + B - 1) / B);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], SM_TILE / 32); }
        mbar_fence_init();
    }
    if (tid < NT) sm.taps[tid] = p.taps[tid];
    __syncthreads();

    if (warp == SM_TILE / 32) {
        // ---------------- producer warp ----------------
        const float *src = p.in + y * p.stride_y + x0;
        const uint64_t pol = l2_evict_first_policy();
        const uint32_t row_bytes = (uint32_t)width * 4u;
        for (int j = 0; j < nblk; ++j) {
            const int s = j % STAGES;
            const int64_t c0 = (int64_t)j * B;
            const int nch = (int)min((int64_t)B, p.nchan - c0);
            if (j >= STAGES) mbar_wait(&sm.empty[s], ((j / STAGES) - 1) & 1);
            if (lane == 0) mbar_expect_tx(&sm.full[s], (uint32_t)nch * row_bytes);
            __syncwarp();
            if (lane < nch)
                tma_load_1d(&sm.data[s][lane][0], src + (c0 + lane) * p.stride_c, row_bytes, &sm.full[s], pol);
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const bool active = tid < width;
    const int64_t x = x0 + tid;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    int cnt = 0;
    bool any_included = false;                                       // for the pass-through fix-up
    char *outp = reinterpret_cast<char *>(p.out) + ((EPI == 1) ? 8 : 4) * (y * p.out_stride_y + x);
    const int64_t out_step = ((EPI == 1) ? 8 : 4) * p.out_stride_c;

    int sprev = STAGES - 1, scur = 0, snext = 1;                     // ring slots of blocks i-1, i, i+1
    mbar_wait(&sm.full[0], 0);
    for (int i = 0; i < nblk; ++i) {
        const int64_t c0 = (int64_t)i * B;
        if (i + 1 < nblk) mbar_wait(&sm.full[snext], ((i + 1) / STAGES) & 1);
        const float *pprev = &sm.data[sprev][B - H][tid];
        const float *pcur = &sm.data[scur][0][tid];
        const float *pnext = &sm.data[snext][0][tid];
        const bool interior = (i >= 1) && (c0 + B + H <= p.nchan);   // every input channel exists

        // ---- gather the B + 2H inputs as float64, apply mask/fill, record NaN and include bits ----
        double w[NIN];
        uint64_t nanbits = 0, incbits = 0;
        float centre_filled[B];
#pragma unroll
        for (int q = 0; q < NIN; ++q) {
            const int64_t cc = c0 - H + q;                           // uniform across the CTA
            float v = 0.0f;
            if (interior || (cc >= 0 && cc < p.nchan)) {
                v = (q < H) ? pprev[q * SM_TILE] : (q < H + B ? pcur[(q - H) * SM_TILE] : pnext[(q - H - B) * SM_TILE]);
                const bool inc = mask_include<MODE>(p.mask, v, cc, y, x);
                if (EPI == 2 && inc && v == v) incbits |= 1ull << q;
                if (EPI != 2 && MODE != MODE_NONE && q >= H && q < H + B) any_included |= inc;
                v = inc ? v : p.fill;
            }
            if (q >= H && q < H + B) centre_filled[q - H] = v;
            const bool isn = v != v;
            if (isn) nanbits |= 1ull << q;
            w[q] = cvt_f64(isn ? 0.0f : v);
        }

        // ---- B outputs: independent n-tap chains with static register indices (taps come from the
        //      constant bank); kept free of branches so the chains interleave on the float64 pipe ----
        double res[B];
#pragma unroll
        for (int o = 0; o < B; ++o) {
            double top = 0.0;
#pragma unroll
            for (int k = 0; k < NT; ++k) top = fma(p.taps[k], w[o + 2 * H - k], top);   // out[c] += K[k] v[c + H - k]
            res[o] = top;
        }
        if (nanbits != 0) {
            // some inputs were NaN: redo the denominator of the outputs whose window saw one
#pragma unroll
            for (int o = 0; o < B; ++o) {
                const uint64_t wb = (nanbits >> o) & ((1ull << NT) - 1ull);
                if (wb != 0) {
                    const double bot = smooth_bot_with_nans<H>(sm.taps, wb, p.ksum, p.ntaps);
                    res[o] = (bot == 0.0) ? (double)centre_filled[o] : res[o] / bot;
                }
            }
        }
#pragma unroll
        for (int o = 0; o < B; ++o) {
            const int64_t c = c0 + o;
            if (interior || c < p.nchan) {                           // uniform
                if (EPI == 0) {
                    if (active) *reinterpret_cast<float *>(outp) = (float)res[o];
                    outp += out_step;
                } else if (EPI == 1) {
                    if (active) *reinterpret_cast<double *>(outp) = res[o];
                    outp += out_step;
                } else {
                    // fused moments: the smoothed value enters as the materialised dtype would hold it,
                    // under the include mask of the ORIGINAL data (spectral_cube.py:3043-3045)
                    const double sv = p.round_f32 ? (double)(float)res[o] : res[o];
                    const bool inc = ((incbits >> (o + H)) & 1ull) && sv == sv;
                    if (inc) {
                        const double2 t = __ldg(p.tab + c);
                        s0 += sv;
                        s1 = fma(sv, t.x, s1);
                        s2 = fma(sv, t.y, s2);
                        cnt += 1;
                    }
                }
            }
        }
        __syncwarp();
        if (i >= 1 && lane == 0) mbar_arrive(&sm.empty[sprev]);
        sprev = scur; scur = snext; snext = (snext + 1 == STAGES) ? 0 : snext + 1;
    }
    if (EPI != 2 && MODE != MODE_NONE && p.passthrough_spaxels && active && !any_included) {
        // `_apply_spectral_function` (spectral_cube.py:147-158): a spaxel with nothing included is
        // copied through (= the fill value everywhere).  Away from the spectral edges the convolution
        // already produced that; the first/last h channels saw valid zero padding and must be redone.
        const int h = p.ntaps >> 1;
        char *o0 = reinterpret_cast<char *>(p.out) + ((EPI == 1) ? 8 : 4) * (y * p.out_stride_y + x);
        for (int64_t c = 0; c < p.nchan; ++c) {
            if (c >= h && c < p.nchan - h) { c = p.nchan - h - 1; continue; }
            if (EPI == 0) *reinterpret_cast<float *>(o0 + c * out_step) = p.fill;
            else          *reinterpret_cast<double *>(o0 + c * out_step) = (double)p.fill;
        }
    }
    if (EPI == 2 && active) {
        const int64_t o = y * p.nx + x;
        const bool any = cnt > 0;
        const double mean = s1 / s0;
        if (p.m0) p.m0[o] = any ? s0 * p.pix_size : nan64();
        if (p.m1) p.m1[o] = any ? (p.K + mean) + p.m1_offset : nan64();
        if (p.m2) p.m2[o] = any ? ((cnt == 1 && s0 != 0.0) ? 0.0 : s2 / s0 - mean * mean) : nan64();
    }
}

// ---- generic fallback: any tap count, any alignment; one thread per output voxel column ----------
template <int MODE, int EPI>
__global__ void __launch_bounds__(128)
smooth_generic_kernel(const __grid_constant__ SmoothParams p, const double *__restrict__ taps, int ntaps) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.ny * p.nx) return;
    const int64_t y = g / p.nx, x = g - y * p.nx;
    const int h = ntaps >> 1;
    const float *src = p.in + y * p.stride_y + x;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    int cnt = 0;
    bool any_included = false;
    for (int64_t c = 0; c < p.nchan; ++c) {
        double top = 0.0, bot = 0.0;
        float centre = 0.0f;
        bool centre_inc = false;
        for (int k = 0; k < ntaps; ++k) {
            const int64_t cc = c + h - k;
            float v = 0.0f;
            if (cc >= 0 && cc < p.nchan) {
                v = __ldg(src + cc * p.stride_c);
                const bool inc = mask_include<MODE>(p.mask, v, cc, y, x);
                if (k == h) { centre_inc = inc && v == v; any_included |= inc; }
                v = inc ? v : p.fill;
            }
            if (k == h) centre = v;
            if (v == v) { top = fma(taps[k], (double)v, top); bot += taps[k]; }
        }
        double res = (bot == 0.0) ? (double)centre : top / bot;
        if (EPI == 0)      reinterpret_cast<float *>(p.out)[c * p.out_stride_c + y * p.out_stride_y + x] = (float)res;
        else if (EPI == 1) reinterpret_cast<double *>(p.out)[c * p.out_stride_c + y * p.out_stride_y + x] = res;
        else {
            const double sv = p.round_f32 ? (double)(float)res : res;
            if (centre_inc && sv == sv) {
                const double2 t = __ldg(p.tab + c);
                s0 += sv; s1 = fma(sv, t.x, s1); s2 = fma(sv, t.y, s2); cnt += 1;
            }
        }
    }
    if (EPI != 2 && MODE != MODE_NONE && p.passthrough_spaxels && !any_included) {
        for (int64_t c = 0; c < p.nchan; ++c) {
            if (EPI == 0) reinterpret_cast<float *>(p.out)[c * p.out_stride_c + y * p.out_stride_y + x] = p.fill;
            else          reinterpret_cast<double *>(p.out)[c * p.out_stride_c + y * p.out_stride_y + x] = (double)p.fill;
        }
    }
    if (EPI == 2) {
        const bool any = cnt > 0;
        const double mean = s1 / s0;
        if (p.m0) p.m0[g] = any ? s0 * p.pix_size : nan64();
        if (p.m1) p.m1[g] = any ? (p.K + mean) + p.m1_offset : nan64();
        if (p.m2) p.m2[g] = any ? ((cnt == 1 && s0 != 0.0) ? 0.0 : s2 / s0 - mean * mean) : nan64();
    }
}

__global__ void moment_table_kernel(const double *__restrict__ x, int64_t n, double K, double2 *tab);

// ---- launch plumbing ---------------------------------------------------------------------------
template <int H, int B, int STAGES, int MODE, int EPI>
static cudaError_t launch_smooth_one(const SmoothParams &p, unsigned grid, cudaStream_t s) {
    auto kern = smooth_tma_kernel<H, B, STAGES, MODE, EPI>;
    const size_t smem = sizeof(SmoothSmem<B, STAGES>);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kern<<<grid, SM_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

template <int H, int B, int STAGES, int EPI>
static cudaError_t launch_smooth_mode(const SmoothParams &p, unsigned grid, cudaStream_t s) {
    switch (p.mask.mode) {
        case MODE_NONE:     return launch_smooth_one<H, B, STAGES, MODE_NONE, EPI>(p, grid, s);
        case MODE_INTERVAL: return launch_smooth_one<H, B, STAGES, MODE_INTERVAL, EPI>(p, grid, s);
        default:            return launch_smooth_one<H, B, STAGES, MODE_GENERIC, EPI>(p, grid, s);
    }
}

template <int EPI>
static cudaError_t launch_smooth_h(const SmoothParams &p, int h, unsigned grid, cudaStream_t s) {
    if (h <= 2)  return launch_smooth_mode<2, 16, 5, EPI>(p, grid, s);
    if (h <= 4)  return launch_smooth_mode<4, 16, 5, EPI>(p, grid, s);
    if (h <= 6)  return launch_smooth_mode<6, 16, 5, EPI>(p, grid, s);
    if (h <= 8)  return launch_smooth_mode<8, 16, 5, EPI>(p, grid, s);
    if (h <= 12) return launch_smooth_mode<12, 16, 5, EPI>(p, grid, s);
    return launch_smooth_mode<16, 16, 5, EPI>(p, grid, s);
}

template <int EPI>
static cudaError_t launch_generic(const SmoothParams &p, const double *taps_dev, int ntaps, cudaStream_t s) {
    const unsigned grid = (unsigned)cdiv(p.ny * p.nx, 128);
    switch (p.mask.mode) {
        case MODE_NONE:     smooth_generic_kernel<MODE_NONE, EPI><<<grid, 128, 0, s>>>(p, taps_dev, ntaps); break;
        case MODE_INTERVAL: smooth_generic_kernel<MODE_INTERVAL, EPI><<<grid, 128, 0, s>>>(p, taps_dev, ntaps); break;
        default:            smooth_generic_kernel<MODE_GENERIC, EPI><<<grid, 128, 0, s>>>(p, taps_dev, ntaps); break;
    }
    return cudaGetLastError();
}

// Normalise the taps like astropy does (kernel /= kernel.sum()) and centre them in a 2H+1 window.
static int prepare_taps(const double *taps, int ntaps, SmoothParams &p, int *h_out, double *norm) {
    SC_CHECK_ARG(taps != nullptr, "taps is NULL");
    SC_CHECK_ARG(ntaps >= 1 && (ntaps & 1), "Kernel size must be odd in all axes. (got %d taps)", ntaps);
    double sum = 0.0;
    for (int k = 0; k < ntaps; ++k) sum += taps[k];
    SC_CHECK_ARG(fabs(sum) > 1e-8, "The kernel can't be normalized, because its sum is close to zero.");
    const int h = ntaps >> 1;
    *h_out = h;
    double ksum = 0.0;
    for (int k = 0; k < ntaps; ++k) { norm[k] = taps[k] / sum; ksum += norm[k]; }
    p.ntaps = ntaps;
    p.ksum = ksum;
    if (h <= 16) {
        const int H = h <= 2 ? 2 : h <= 4 ? 4 : h <= 6 ? 6 : h <= 8 ? 8 : h <= 12 ? 12 : 16;
        for (int k = 0; k < SM_MAX_TAPS; ++k) p.taps[k] = 0.0;
        for (int k = 0; k < ntaps; ++k) p.taps[k + (H - h)] = norm[k];
    }
    return SC_OK;
}

template <int EPI>
static int run_smooth(SmoothParams &p, const sc_mask_desc *mask, const double *taps, int ntaps,
                      void *workspace, size_t workspace_bytes, size_t ws_offset, cudaStream_t s, int op) {
    int h = 0;
    double norm[1024];
    SC_CHECK_ARG(ntaps <= 1023, "at most 1023 taps are supported");
    int rc = prepare_taps(taps, ntaps, p, &h, norm);
    if (rc) return rc;
    rc = build_dev_mask(mask, p.in, p.stride_c, p.stride_y, &p.mask);
    if (rc) return rc;
    const bool aligned = ((uintptr_t)p.in % 16 == 0) && p.stride_c % 4 == 0 && p.stride_y % 4 == 0 && p.nx % 4 == 0;
    const int choice = env_int("SC_SMOOTH_KERNEL", 0);            // 0 auto, 1 generic, 2 tma
    const int64_t tiles_per_row = cdiv(p.nx, SM_TILE);
    const int64_t n_tiles = tiles_per_row * p.ny;
    LaunchScope ls(op, s);
    if (h <= 16 && aligned && choice != 1 && n_tiles < ((int64_t)1 << 31)) {
        p.tiles_per_row = (int)tiles_per_row;
        cudaError_t e = launch_smooth_h<EPI>(p, h, (unsigned)n_tiles, s);
        if (e != cudaSuccess) return cuda_fail(e, "smooth_tma_kernel launch");
        return SC_OK;
    }
    const size_t need = ws_offset + (size_t)ntaps * 8 + 256;
    if (!workspace || workspace_bytes < need) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return SC_ERR_WORKSPACE;
    }
    double *taps_dev = (double *)((((uintptr_t)workspace + ws_offset) + 255) & ~(uintptr_t)255);
    SC_CUDA(cudaMemcpyAsync(taps_dev, norm, (size_t)ntaps * 8, cudaMemcpyHostToDevice, s));
    cudaError_t e = launch_generic<EPI>(p, taps_dev, ntaps, s);
    if (e != cudaSuccess) return cuda_fail(e, "smooth_generic_kernel launch");
    return SC_OK;
}

}  // namespace scb

using namespace scb;

extern "C" int sc_spectral_smooth(const float *in, void *out, int out_dtype,
                                  int64_t nchan, int64_t ny, int64_t nx,
                                  int64_t stride_c, int64_t stride_y,
                                  int64_t out_stride_c, int64_t out_stride_y,
                                  const sc_mask_desc *mask, double fill,
                                  const double *taps, int ntaps, int spaxel_passthrough,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_cube_args(in, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(out != nullptr, "out is NULL");
    SC_CHECK_ARG(out_dtype == SC_F32 || out_dtype == SC_F64, "out_dtype must be SC_F32 or SC_F64");
    SC_CHECK_ARG(out_stride_y >= nx && out_stride_c >= out_stride_y, "bad output strides");
    SC_CHECK_ARG((const void *)in != out || out_dtype == SC_F32, "in-place smoothing needs a float32 output");
    cudaStream_t s = (cudaStream_t)stream;
    SmoothParams p{};
    p.in = in; p.out = out; p.nchan = nchan; p.ny = ny; p.nx = nx;
    p.stride_c = stride_c; p.stride_y = stride_y; p.out_stride_c = out_stride_c; p.out_stride_y = out_stride_y;
    p.fill = (float)fill;
    // `_apply_spectral_function` copies a spaxel with nothing included (spectral_cube.py:155-158)
    p.passthrough_spaxels = spaxel_passthrough ? 1 : 0;
    if (out_dtype == SC_F32) return run_smooth<0>(p, mask, taps, ntaps, workspace, workspace_bytes, 0, s, SC_OP_SPECTRAL_SMOOTH);
    return run_smooth<1>(p, mask, taps, ntaps, workspace, workspace_bytes, 0, s, SC_OP_SPECTRAL_SMOOTH);
}

extern "C" int sc_smooth_moments_axis0(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                                       int64_t stride_c, int64_t stride_y,
                                       const sc_mask_desc *mask, double fill,
                                       const double *taps, int ntaps, int smooth_dtype,
                                       const double *chan_offset, double pix_size, double m1_offset,
                                       int want_bits, double *out_m0, double *out_m1, double *out_m2,
                                       void *workspace, size_t workspace_bytes, void *stream) {
    int rc = check_cube_args(cube, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(want_bits > 0 && want_bits <= 7, "want_bits=%d must be a combination of 1|2|4", want_bits);
    SC_CHECK_ARG(!(want_bits & SC_WANT_M0) || out_m0, "out_m0 is NULL but moment 0 was requested");
    SC_CHECK_ARG(!(want_bits & SC_WANT_M1) || out_m1, "out_m1 is NULL but moment 1 was requested");
    SC_CHECK_ARG(!(want_bits & SC_WANT_M2) || out_m2, "out_m2 is NULL but moment 2 was requested");
    SC_CHECK_ARG(chan_offset != nullptr, "chan_offset is NULL");
    SC_CHECK_ARG(smooth_dtype == SC_F32 || smooth_dtype == SC_F64, "smooth_dtype must be SC_F32 or SC_F64");
    const size_t need = (size_t)nchan * 24 + 512;
    if (!workspace || workspace_bytes < need) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return SC_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const double K = chan_offset[nchan / 2];
    double2 *tab = (double2 *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    double *xdev = (double *)(tab + nchan);
    SC_CUDA(cudaMemcpyAsync(xdev, chan_offset, (size_t)nchan * 8, cudaMemcpyHostToDevice, s));
    {
        LaunchScope ls(0, s);
        moment_table_kernel<<<(unsigned)cdiv(nchan, 256), 256, 0, s>>>(xdev, nchan, K, tab);
        SC_CUDA(cudaGetLastError());
    }
    SmoothParams p{};
    p.in = cube; p.out = nullptr; p.nchan = nchan; p.ny = ny; p.nx = nx;
    p.stride_c = stride_c; p.stride_y = stride_y;
    p.fill = (float)fill;
    p.tab = tab; p.K = K; p.pix_size = pix_size; p.m1_offset = m1_offset;
    p.m0 = (want_bits & SC_WANT_M0) ? out_m0 : nullptr;
    p.m1 = (want_bits & SC_WANT_M1) ? out_m1 : nullptr;
    p.m2 = (want_bits & SC_WANT_M2) ? out_m2 : nullptr;
    p.round_f32 = smooth_dtype == SC_F32;
    return run_smooth<2>(p, mask, taps, ntaps, workspace, workspace_bytes, need, s, SC_OP_SMOOTH_MOMENTS);
}
