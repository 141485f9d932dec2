// spectral_smooth.cu -- per-spaxel NaN-interpolating convolution along the spectral axis, and
// the fused spectral_smooth -> moment0/1/2 chain (BASELINE config 3).
//
// Replaces `convolve(spectrum, kernel, normalize_kernel=True)` applied to every spaxel
// (spectral_cube.py:3186-3222 through :3103-3159 and :147-158; dask_spectral_cube.py:880-917).
// astropy semantics: true convolution (kernel flipped), zero-filled boundary whose zeros are
// VALID samples, out = sum_k K[k] v[c-k] [v not NaN] / sum_k K[k] [v not NaN] in float64,
// denominator 0 keeps the input value.  Masked voxels are replaced by `fill` before convolving.
//
// Kernel (smooth_tma_kernel).  A CTA owns SM_TILE (128) adjacent spaxels of one image row, one per
// thread.  Blocks of B channels x SM_TILE floats (one tiled TMA request per block,
// `cp.async.bulk.tensor.3d`) stream through a STAGES-deep shared-memory ring; arrival is an mbarrier per slot, and
// the warp that is LAST to finish with a slot refills it at once (no producer role: a producer
// that also computes paces the whole CTA, see the kernel).  Per block of B output
// channels a thread
//   * takes its B NEW inputs from the ring (every sample is loaded from shared memory, masked
//     and widened to float64 exactly once; the 2H older inputs it still needs stay in registers:
//     for 2H <= B they are simply the tail of the previous block's array -- the loop is unrolled
//     by two and the two arrays swap roles, so no register is moved),
//   * widens float32 -> float64 WITHOUT the conversion unit: the float's bits are placed in the
//     double's fields (one 64-bit multiply-shift + one AND), which yields v * 2^-896 exactly for
//     every finite v including zeros and denormals; the taps are pre-scaled by 2^896 on the host
//     so every product and sum is bit-identical to the unscaled arithmetic.  (F2F runs at ~9
//     lanes/clk/SM on this part and was the top limiter of the first version.)
//   * runs B independent n-tap FMA chains with compile-time register indices and taps read
//     from the constant bank, branch-free so the chains interleave on the float64 pipe,
//   * fixes up the denominator only where the rolling NaN bit mask says a window saw a NaN:
//     sum of the PRESENT taps from a 6-bit-chunk look-up table in shared memory.
// Outputs are stored straight from registers (a warp writes 128 contiguous bytes per channel) or,
// in the fused variant, rounded to the smoothed cube's dtype and accumulated into the moment sums
// under the include mask of the ORIGINAL data -- the smoothed cube is never written.
#include "common.cuh"
#include "tma.cuh"
#include <type_traits>

namespace scb {

int check_cube_args(const float *cube, int64_t nchan, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y);
int env_int(const char *name, int dflt);

constexpr int SM_TILE = 128;                 // spaxels per CTA = threads (512-byte rows; 256 measured 5 % slower, 64 8 % slower)
constexpr int SM_THREADS = SM_TILE;          // warp 0 also issues the TMA copies
constexpr int SM_MIN_CTAS = 4;               // register budget: 128 per thread; four 4-warp rings per SM wait less on their slowest warp than two 8-warp ones
constexpr int SM_MAX_TAPS = 2 * 16 + 1;
constexpr int SM_LUT_CHUNKS = (SM_MAX_TAPS + 5) / 6;
constexpr int SM_B = 16;                     // output channels per block
constexpr int SM_STAGES = 6;
constexpr int SM_G = 8;                      // outputs whose chains run side by side (register budget)

struct SmoothParams {
    alignas(64) CUtensorMap tmap;            // the input cube; box = {tile, 1, block channels}
    const float *in;
    void *out;                               // float32 or float64
    int64_t nchan, ny, nx;
    int64_t stride_c, stride_y, out_stride_c, out_stride_y;
    int tiles_per_row;
    int ntaps;                               // actual number of taps (<= 2H+1)
    float fill;
    double ksum;                             // sum of the normalised taps (~1)
    double taps[SM_MAX_TAPS];                // normalised, centred in a 2H+1 window, zero padded
    double taps_scaled[SM_MAX_TAPS];         // the same times 2^896 (see place_scaled)
    int passthrough_spaxels;                 // numpy class: spaxels with nothing included are copied through
    // fused moments
    const double2 *tab;                      // {d, d^2} per channel
    double K, pix_size, m1_offset;
    double *m0, *m1, *m2;
    int round_f32;                           // smoothed cube dtype is float32 (dask class)
    DevMask mask;
};

struct SmoothSmem {
    float data[SM_STAGES][SM_B][SM_TILE];
    double taps[SM_MAX_TAPS];
    double lut[SM_LUT_CHUNKS][64];           // lut[ch][m] = sum of taps whose window bits (6 ch + b) are set in m
    double tmul[SM_MAX_TAPS + 2 * SM_B];     // tmul[B + j] = 1 / (sum of all taps but the one under window bit j), 1 outside [0, 2H]
    int rmul_degenerate;                     // some such sum is 0: the exact path handles every window with a NaN
    uint64_t full[SM_STAGES];
    unsigned int done[SM_STAGES];            // warps that have finished with the block in the slot
};

// v * 2^-896 as a double, exactly, for every finite v (zeros and denormals included): the float's
// exponent and mantissa fields are dropped into the double's.  hi:lo = (int64)bits << 29, then the
// three sign-extension bits below the sign are cleared.  +-inf keeps being +-inf when KEEP_INF.
template <bool KEEP_INF>
__device__ __forceinline__ double place_scaled(float v) {
    const int b = __float_as_int(v);
    const long long t = (long long)b << 29;
    int hi = (int)(t >> 32) & 0x8FFFFFFF;
    if (KEEP_INF && (b & 0x7F800000) == 0x7F800000) hi |= 0x7FF00000;
    return __hiloint2double(hi, (int)t);
}

// Denominator of a window with missing (NaN) inputs: `present` has bit j set when window entry j
// (input o + j, tap 2H - j) is a number.  Sum of the present taps, six window bits per table look-up;
// exactly 0 when nothing is present.
template <int H>
__device__ __forceinline__ double smooth_bot_present(const double (*lut)[64], uint64_t present) {
    constexpr int NCH = (2 * H + 1 + 5) / 6;
    double bot = 0.0;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) bot += lut[ch][(present >> (6 * ch)) & 63u];
    return bot;
}

// In the smoothing kernels MODE_INTERVAL means "interval mask AND the fill value is NaN" (the host
// sends finite fills to MODE_GENERIC), so "excluded" and "NaN" collapse into one test.
// TAIL = false: the caller guarantees the channel exists (`valid` is ignored and folds away).
template <int MODE, int EPI, bool TAIL, typename bits_t>
__device__ __forceinline__ double smooth_take(const SmoothParams &p, float v, bool valid, int q, int64_t cc, int64_t y, int64_t x,
                                              bits_t &nanbits, bits_t &incbits, bool &any_included) {
    // `valid` (uniform): the channel exists; beyond the cube the sample is a VALID zero that no mask touches
    constexpr bool TRACK_ANY = TAIL || MODE != MODE_INTERVAL;        // full INTERVAL blocks derive it from nanbits
    const bits_t bit = (bits_t)1 << q;
    if (MODE == MODE_INTERVAL) {
        const bool inr = (v > p.mask.lo) & (v < p.mask.hi);
        const bool use = TAIL ? (inr & valid) : inr;                 // a real, included, finite sample
        if (EPI == 2) { if (use) incbits |= bit; }
        else if (TRACK_ANY) any_included |= use;
        if (!(TAIL ? (inr | !valid) : inr)) nanbits |= bit;
        return place_scaled<false>(use ? v : 0.0f);
    }
    const bool ok = TAIL ? valid : true;
    const bool inc = ok && mask_include<MODE>(p.mask, v, cc, y, x);
    if (EPI == 2) { if (inc && v == v) incbits |= bit; }
    else if (MODE != MODE_NONE) any_included |= inc;
    v = ok ? (inc ? v : p.fill) : 0.0f;
    const bool isn = v != v;
    if (isn) nanbits |= bit;
    return place_scaled<true>(isn ? 0.0f : v);
}

struct SmoothAcc { double s0, s1, s2; int cnt; };

// Rare path, out of line (one copy per kernel, a rolled loop: the hot loop must stay small enough for
// the instruction cache): the windows of a group of SM_G outputs saw several missing (NaN) inputs.
// `nb` = nanbits >> g: bit j of (nb >> oo) is set when window entry j of output oo is missing.
// res[oo] <- top / (sum of the present taps), or the (filled) input value when nothing under the
// kernel is valid.  `centre` points at the ring sample of output 0 of the group (row stride SM_TILE);
// outputs oo >= nvalid lie beyond the last channel.
template <int H, int MODE, typename bits_t>
__device__ __noinline__ void smooth_fix_group(const SmoothParams &p, const double (*lut)[64], double *res, bits_t nb,
                                              const float *centre, int nvalid, int64_t c, int64_t y, int64_t x) {
    constexpr bits_t FULL = (bits_t)((1ull << (2 * H + 1)) - 1ull);
#pragma unroll 1
    for (int oo = 0; oo < SM_G; ++oo) {
        const bits_t wb = (bits_t)(nb >> oo) & FULL;
        if (wb == 0) continue;
        const double bot = wb == FULL ? 0.0 : smooth_bot_present<H>(lut, (uint64_t)(bits_t)(~wb & FULL));
        if (bot != 0.0) { res[oo] = res[oo] / bot; continue; }
        float cv = 0.0f;                                         // nothing valid under the kernel
        if (oo < nvalid) {
            cv = centre[oo * SM_TILE];
            if (!mask_include<MODE>(p.mask, cv, c + oo, y, x)) cv = p.fill;
        }
        res[oo] = (double)cv;
    }
}

// Blank block, out of line: every sample of every lane's window is missing (the blanked frame of a
// mosaic): the outputs are the filled inputs themselves.
template <int MODE, int EPI>
__device__ __noinline__ void smooth_blank_block(const SmoothParams &p, const float *pcur, char *outp, int64_t out_step,
                                                int nvalid, int64_t c0, int64_t y, int64_t x) {
#pragma unroll 1
    for (int o = 0; o < nvalid; ++o) {
        float cv = pcur[o * SM_TILE];
        if (!mask_include<MODE>(p.mask, cv, c0 + o, y, x)) cv = p.fill;
        if (EPI == 0) *reinterpret_cast<float *>(outp + o * out_step) = cv;
        else          *reinterpret_cast<double *>(outp + o * out_step) = (double)cv;
    }
}

// Outputs [g, g + SM_G) of a block: window entry q is car[car_off + q] for q < 2H, cur[q - 2H] after.
template <int H, int MODE, int EPI, bool TAIL, typename bits_t, int NCAR>
__device__ __forceinline__ void smooth_emit(const SmoothParams &p, SmoothSmem &sm, const double (&car)[NCAR], int car_off,
                                            const double (&cur)[SM_B], const float *pcur, int g,
                                            int64_t c0, int64_t y, int64_t x, bool active,
                                            bits_t nanbits, bits_t incbits, const double *tq, char *outp, int64_t out_step, SmoothAcc &acc) {
    constexpr int NT = 2 * H + 1;
    double res[SM_G];
#pragma unroll
    for (int oo = 0; oo < SM_G; ++oo) {
        const int o = g + oo;
        double top = 0.0;
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            const int q = o + 2 * H - k;                         // out[c] += K[k] v[c + H - k]
            top = fma(p.taps_scaled[k], q < 2 * H ? car[car_off + q] : cur[q - 2 * H], top);
        }
        res[oo] = top;
    }
    constexpr bits_t GROUP = (bits_t)((1ull << (NT + SM_G - 1)) - 1ull);
    if (((nanbits >> g) & GROUP) != 0) {
        // Some inputs under these windows were NaN: rescale the outputs that saw one.  Usual case
        // (`tq` set by the caller): the thread has ONE missing input, window entry Q, which sits under
        // window bit Q - o of output o; the multiplier 1 / (sum of the other taps) is tmul[Q - o],
        // padded with ones where the input is outside the window: one LDS + one DMUL per output.
        // Several missing inputs (or a degenerate table): the exact path.
        if (tq != nullptr) {
#pragma unroll
            for (int oo = 0; oo < SM_G; ++oo) res[oo] *= tq[-(g + oo)];
        } else if (MODE != MODE_GENERIC && ((nanbits >> g) & GROUP) == GROUP) {
            // every input under every window of the group is missing (a blank spaxel in a warp that
            // also holds data): the output is the filled input, NaN in these two modes
#pragma unroll
            for (int oo = 0; oo < SM_G; ++oo) res[oo] = nan64();
        } else {
            double tmp[SM_G];                                        // through local memory on this path only
#pragma unroll
            for (int oo = 0; oo < SM_G; ++oo) tmp[oo] = res[oo];
            const int nvalid = TAIL ? (int)max((int64_t)0, min((int64_t)SM_G, p.nchan - (c0 + g))) : SM_G;
            smooth_fix_group<H, MODE, bits_t>(p, sm.lut, tmp, (bits_t)(nanbits >> g), pcur + g * SM_TILE, nvalid, c0 + g, y, x);
#pragma unroll
            for (int oo = 0; oo < SM_G; ++oo) res[oo] = tmp[oo];
        }
    }
#pragma unroll
    for (int oo = 0; oo < SM_G; ++oo) {
        const int o = g + oo;
        const int64_t c = c0 + o;
        if (!TAIL || c < p.nchan) {                              // uniform
            if (EPI == 0) {
                if (active) *reinterpret_cast<float *>(outp + o * out_step) = (float)res[oo];
            } else if (EPI == 1) {
                if (active) *reinterpret_cast<double *>(outp + o * out_step) = res[oo];
            } else {
                // fused moments: the smoothed value enters as the materialised dtype would hold it,
                // under the include mask of the ORIGINAL data (spectral_cube.py:3043-3045)
                const double sv = p.round_f32 ? (double)(float)res[oo] : res[oo];
                const bool inc = ((incbits >> (o + H)) & 1u) && sv == sv;
                if (inc) {
                    const double2 t = __ldg(p.tab + c);
                    acc.s0 += sv;
                    acc.s1 = fma(sv, t.x, acc.s1);
                    acc.s2 = fma(sv, t.y, acc.s2);
                    acc.cnt += 1;
                }
            }
        }
    }
}

// One block of B outputs.  `car` holds the 2H inputs before this block's new ones at
// car[car_off .. car_off + 2H) (window entries [0, 2H)); `cur` receives the B new inputs
// (entries [2H, 2H+B)).  TAIL = false: every channel this block reads or writes exists.
template <int H, int MODE, int EPI, bool TAIL, typename bits_t, int NCAR>
__device__ __forceinline__ void smooth_block(const SmoothParams &p, SmoothSmem &sm, const double (&car)[NCAR], int car_off,
                                             double (&cur)[SM_B], const float *pcur, const float *pnext,
                                             int64_t c0, int64_t y, int64_t x, bool active,
                                             bits_t &nanbits, bits_t &incbits, bool &any_included,
                                             char *&outp, int64_t out_step, SmoothAcc &acc) {
    constexpr int B = SM_B, NT = 2 * H + 1;
    // ---- the B new inputs: channels [c0+H, c0+B+H) = rows [H, B) of this stage, rows [0, H) of the next ----
#pragma unroll
    for (int n = 0; n < B; ++n) {
        const int q = 2 * H + n;
        const int64_t cc = c0 + H + n;                           // uniform across the CTA
        // the ring row is always addressable; beyond the cube it holds stale bytes that `valid` discards
        const float v = (n < B - H) ? pcur[(n + H) * SM_TILE] : pnext[(n + H - B) * SM_TILE];
        cur[n] = smooth_take<MODE, EPI, TAIL, bits_t>(p, v, TAIL ? cc < p.nchan : true, q, cc, y, x, nanbits, incbits, any_included);
    }
    constexpr bits_t NEWBITS = (bits_t)((((bits_t)1 << B) - 1) << (2 * H));
    if (!TAIL && MODE == MODE_INTERVAL && EPI != 2) any_included |= (nanbits & NEWBITS) != NEWBITS;
    // ---- blank spectra (every sample of every lane's window missing, e.g. the blanked frame of a mosaic):
    //      the result is the filled input itself, no arithmetic needed ----
    constexpr bits_t ALLBITS = (bits_t)(~(bits_t)0) >> (8 * sizeof(bits_t) - (B + 2 * H));
    if (EPI != 2 && __all_sync(0xffffffffu, nanbits == ALLBITS)) {
        if (active)
            smooth_blank_block<MODE, EPI>(p, pcur, outp, out_step, TAIL ? (int)max((int64_t)0, min((int64_t)B, p.nchan - c0)) : B, c0, y, x);
        outp += B * out_step;
        nanbits >>= B;
        incbits >>= B;
        return;
    }
    // ---- NaN bookkeeping: one missing input in the whole window -> table pointer, more -> exact path ----
    const double *tq = nullptr;
    if (nanbits != 0 && (nanbits & (nanbits - 1)) == 0 && !sm.rmul_degenerate) {
        const int Q = (8 * (int)sizeof(bits_t) - 1) - (sizeof(bits_t) == 4 ? __clz((int)nanbits) : __clzll((long long)nanbits));
        tq = &sm.tmul[Q + B];
    }
    // ---- the outputs, in groups of SM_G: SM_G independent n-tap chains, NaN fix-up, store ----
#pragma unroll
    for (int g = 0; g < B; g += SM_G)
        smooth_emit<H, MODE, EPI, TAIL, bits_t, NCAR>(p, sm, car, car_off, cur, pcur, g, c0, y, x, active, nanbits, incbits, tq, outp, out_step, acc);
    outp += B * out_step;
    nanbits >>= B;
    incbits >>= B;
}

// EPI 0: store float32, 1: store float64, 2: fused moments
template <int H, int MODE, int EPI>
__global__ void __launch_bounds__(SM_THREADS, SM_MIN_CTAS)
smooth_tma_kernel(const __grid_constant__ SmoothParams p) {
    constexpr int B = SM_B, STAGES = SM_STAGES, TILE_PX = SM_TILE;
    constexpr int NT = 2 * H + 1;
    constexpr int NIN = B + 2 * H;
    constexpr int NWARPS = SM_THREADS / 32;
    using bits_t = typename std::conditional<(NIN <= 32), uint32_t, uint64_t>::type;
    static_assert(H <= B, "window must fit in two neighbouring stages");
    static_assert(NIN <= 64, "NaN bit mask is 64 bits");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmoothSmem &sm = *reinterpret_cast<SmoothSmem *>(smem_raw);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int64_t tile = blockIdx.x;
    const int64_t y = tile / p.tiles_per_row;
    const int64_t x0 = (tile - y * p.tiles_per_row) * SM_TILE;
    const int width = (int)min((int64_t)SM_TILE, p.nx - x0);
    const int nblk = (int)((p.nchan + B - 1) / B);
    // blocks [0, nfull) touch existing channels only: (i + 1) B + H <= nchan
    const int nfull = p.nchan >= H ? (int)((p.nchan - H) / B) : 0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); sm.done[s] = 0; }
        sm.rmul_degenerate = 0;
        mbar_fence_init();
    }
    if (tid < NT) sm.taps[tid] = p.taps[tid];
    __syncthreads();
    for (int e = tid; e < ((NT + 5) / 6) * 64; e += SM_THREADS) {
        const int ch = e >> 6, m = e & 63;
        double a = 0.0;
        for (int b = 0; b < 6; ++b) {
            const int j = 6 * ch + b;                                // window bit j <-> tap 2H - j
            if (((m >> b) & 1) && j < NT) a += sm.taps[2 * H - j];
        }
        sm.lut[ch][m] = a;
    }
    for (int u = tid; u < NT + 2 * B; u += SM_THREADS) {
        const int t = u - B;                                         // window bit
        double r = 1.0;
        if (t >= 0 && t < NT) {
            double a = 0.0;                                          // same summation order as the LUT path
            for (int j = 0; j < NT; ++j) if (j != t) a += sm.taps[2 * H - j];
            if (a != 0.0) r = 1.0 / a; else atomicOr(&sm.rmul_degenerate, 1);
        }
        sm.tmul[u] = r;
    }
    __syncthreads();

    // Ring refill without a producer role: block j lives in slot j % STAGES.  Every warp bumps the
    // slot's `done` counter when it has read the block; the warp that arrives LAST knows the slot is
    // free and issues block j + STAGES into it at once.  (A fixed producer warp that also computes
    // throttles the whole CTA to its own single-warp pace: the others run ahead, drain the ring and
    // wait for it.)
    // one tiled TMA request per ring slot (rows past the last channel / columns past nx arrive as zeros)
    const uint64_t pol = l2_evict_first_policy();
    auto issue_block = [&](int j) {
        const int s = j % STAGES;
        if (lane == 0) {
            mbar_expect_tx(&sm.full[s], (uint32_t)(B * TILE_PX * 4));
            tma_load_box3d(&sm.data[s][0][0], &p.tmap, (int)x0, (int)y, j * B, &sm.full[s], pol);
        }
    };
    if (warp == 0)
        for (int j = 0; j < STAGES && j < nblk; ++j) issue_block(j);

    const bool active = tid < width;
    const int64_t x = x0 + tid;
    SmoothAcc acc{0.0, 0.0, 0.0, 0};
    bool any_included = false;                                       // for the pass-through fix-up
    char *outp = reinterpret_cast<char *>(p.out) + ((EPI == 1) ? 8 : 4) * (y * p.out_stride_y + x);
    const int64_t out_step = ((EPI == 1) ? 8 : 4) * p.out_stride_c;
    bits_t nanbits = 0, incbits = 0;

    // Window storage.  2H <= B: two B-entry arrays that swap roles block by block (the carried 2H
    // inputs are simply the tail of the previous block's array: no register is moved).  2H > B:
    // a 2H-entry carry array that is shifted by B after every block.
    constexpr bool SWAP = 2 * H <= B;
    constexpr int NCAR = SWAP ? B : 2 * H;
    constexpr int CAR_OFF = SWAP ? B - 2 * H : 0;
    double wa[NCAR], wb[B];

    // prologue: channels [-H, 0) are (valid) zeros, channels [0, H) come from the first stage; they
    // form window entries [0, 2H) of block 0
    mbar_wait(&sm.full[0], 0);
#pragma unroll
    for (int q = 0; q < 2 * H; ++q) {
        wa[CAR_OFF + q] = 0.0;
        if (q >= H)
            wa[CAR_OFF + q] = smooth_take<MODE, EPI, true, bits_t>(p, sm.data[0][q - H][tid], q - H < p.nchan, q, q - H, y, x, nanbits, incbits, any_included);
    }

    int scur = 0, snext = 1 % STAGES;
    // after a block: hand the slot back and, as the last warp to do so, refill it
    auto release = [&](int i) {
        __syncwarp();
        int last = 0;
        // (no fence: the ring rows were consumed by arithmetic that precedes this point in issue order)
        if (lane == 0) {
            last = atomicAdd(&sm.done[scur], 1u) == (unsigned)(NWARPS - 1);
            if (last) sm.done[scur] = 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && i + STAGES < nblk) issue_block(i + STAGES);
        scur = snext; snext = (snext + 1 == STAGES) ? 0 : snext + 1;
    };

    int i = 0;
    if constexpr (SWAP) {
        // pairs of full blocks: wa -> wb, then wb -> wa
        for (; i + 2 <= nfull; i += 2) {
            mbar_wait(&sm.full[snext], ((i + 1) / STAGES) & 1);
            smooth_block<H, MODE, EPI, false, bits_t, NCAR>(p, sm, wa, CAR_OFF, wb, &sm.data[scur][0][tid], &sm.data[snext][0][tid],
                                                            (int64_t)i * B, y, x, active, nanbits, incbits, any_included, outp, out_step, acc);
            release(i);
            if (i + 2 < nblk) mbar_wait(&sm.full[snext], ((i + 2) / STAGES) & 1);
            smooth_block<H, MODE, EPI, false, bits_t, B>(p, sm, wb, B - 2 * H, wa, &sm.data[scur][0][tid], &sm.data[snext][0][tid],
                                                         (int64_t)(i + 1) * B, y, x, active, nanbits, incbits, any_included, outp, out_step, acc);
            release(i + 1);
        }
    } else {
        for (; i < nfull; ++i) {
            if (i + 1 < nblk) mbar_wait(&sm.full[snext], ((i + 1) / STAGES) & 1);
            smooth_block<H, MODE, EPI, false, bits_t, NCAR>(p, sm, wa, CAR_OFF, wb, &sm.data[scur][0][tid], &sm.data[snext][0][tid],
                                                            (int64_t)i * B, y, x, active, nanbits, incbits, any_included, outp, out_step, acc);
#pragma unroll
            for (int q = 0; q < NCAR; ++q) wa[q] = (q + B < NCAR) ? wa[q + B] : wb[q + B - NCAR];
            release(i);
        }
    }
    // the remaining blocks (an odd full one, and those that run past the last channel)
    for (; i < nblk; ++i) {
        if (i + 1 < nblk) mbar_wait(&sm.full[snext], ((i + 1) / STAGES) & 1);
        smooth_block<H, MODE, EPI, true, bits_t, NCAR>(p, sm, wa, CAR_OFF, wb, &sm.data[scur][0][tid], &sm.data[snext][0][tid],
                                                       (int64_t)i * B, y, x, active, nanbits, incbits, any_included, outp, out_step, acc);
        // slide the window by B channels
#pragma unroll
        for (int q = 0; q < NCAR; ++q) wa[q] = (q + B < NCAR) ? wa[q + B] : wb[q + B - NCAR];
        release(i);
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
        const bool any = acc.cnt > 0;
        const double mean = acc.s1 / acc.s0;
        if (p.m0) p.m0[o] = any ? acc.s0 * p.pix_size : nan64();
        if (p.m1) p.m1[o] = any ? (p.K + mean) + p.m1_offset : nan64();
        if (p.m2) p.m2[o] = any ? ((acc.cnt == 1 && acc.s0 != 0.0) ? 0.0 : acc.s2 / acc.s0 - mean * mean) : nan64();
    }
}

// ---- generic fallback: any tap count, any alignment; one thread per spaxel ---------------------------
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
template <int H, int MODE, int EPI>
static cudaError_t launch_smooth_one(const SmoothParams &p, unsigned grid, cudaStream_t s) {
    auto kern = smooth_tma_kernel<H, MODE, EPI>;
    const size_t smem = sizeof(SmoothSmem);
    static unsigned long long configured = 0;        // per instantiation, one bit per device
    if (cudaError_t e = ensure_dyn_smem(kern, smem, &configured)) return e;
    kern<<<grid, SM_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

// kernel-side mask mode: the interval form is only used when the fill value is NaN
static int smooth_mode(const SmoothParams &p) {
    if (p.mask.mode == MODE_INTERVAL && p.fill == p.fill) return MODE_GENERIC;
    return p.mask.mode;
}

template <int H, int EPI>
static cudaError_t launch_smooth_mode(const SmoothParams &p, unsigned grid, cudaStream_t s) {
    switch (smooth_mode(p)) {
        case MODE_NONE:     return launch_smooth_one<H, MODE_NONE, EPI>(p, grid, s);
        case MODE_INTERVAL: return launch_smooth_one<H, MODE_INTERVAL, EPI>(p, grid, s);
        default:            return launch_smooth_one<H, MODE_GENERIC, EPI>(p, grid, s);
    }
}

static int padded_half(int h) { return h <= 2 ? 2 : h <= 4 ? 4 : h <= 8 ? 8 : 16; }

template <int EPI>
static cudaError_t launch_smooth_h(const SmoothParams &p, int h, unsigned grid, cudaStream_t s) {
    switch (padded_half(h)) {
        case 2:  return launch_smooth_mode<2, EPI>(p, grid, s);
        case 4:  return launch_smooth_mode<4, EPI>(p, grid, s);
        case 8:  return launch_smooth_mode<8, EPI>(p, grid, s);
        default: return launch_smooth_mode<16, EPI>(p, grid, s);
    }
}

template <int EPI>
static cudaError_t launch_generic(const SmoothParams &p, const double *taps_dev, int ntaps, cudaStream_t s) {
    const unsigned grid = (unsigned)cdiv(p.ny * p.nx, 128);
    switch (smooth_mode(p)) {
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
        const int H = padded_half(h);
        for (int k = 0; k < SM_MAX_TAPS; ++k) { p.taps[k] = 0.0; p.taps_scaled[k] = 0.0; }
        for (int k = 0; k < ntaps; ++k) {
            p.taps[k + (H - h)] = norm[k];
            p.taps_scaled[k + (H - h)] = ldexp(norm[k], 896);       // exact: undoes place_scaled's 2^-896
        }
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
        int trc = make_cube_tensor_map(&p.tmap, p.in, p.nchan, p.ny, p.nx, p.stride_c, p.stride_y,
                                       SM_TILE, SM_B);
        if (trc) return trc;
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
