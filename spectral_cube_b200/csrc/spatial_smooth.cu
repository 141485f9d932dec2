// spatial_smooth.cu -- per-channel 2-D NaN-interpolating convolution.
//
// Replaces `convolve(image, kernel2d, normalize_kernel=True)` applied to every channel
// (spectral_cube.py:2808-2842 through :3049-3101 and :161-172; dask_spectral_cube.py:962-993).
// Same astropy semantics as spectral_smooth.cu: true convolution, zero-filled boundary whose
// zeros are valid samples, top/bot NaN interpolation, float64 numerator.
//
// Separable kernel (sep_march_kernel), used for outer-product kernels (Gaussian2DKernel):
//   top = conv_y(conv_x(v * ok)), bot = conv_y(conv_x(ok)) are both exactly separable.
//   A CTA owns a strip of SP_TX columns of one channel and MARCHES down the rows in blocks of
//   R = 8: a producer warp streams the raw rows (strip + horizontal halo) into a shared-memory
//   ring with cp.async.bulk (TMA) + mbarriers; the compute warps (1) apply mask/fill in place,
//   (2) run the row pass with register blocking (a thread makes 8 adjacent outputs from 8+2H
//   inputs read as float4) into a ring of row-passed lines {top f64, bot f32, centre f32},
//   (3) run the column pass for the block that now has its 2H neighbours (a thread makes 8
//   vertically adjacent outputs from 8+2H ring lines) and stores.  Every input row is
//   row-passed once, so the float64 work is 2(2H+1) FMAs per voxel -- the float64 pipe, not HBM,
//   bounds this kernel.  The denominator runs on the float32 pipe (it only enters the result
//   as a ratio, so its rounding is a pure relative error of ~1e-7).
//   Row sharding: rows above/below the shard come from `halo_top` / `halo_bot` (the
//   neighbouring ranks' filled edge rows) or are zero at the image boundary.
//
// Direct kernel (direct2d_kernel): any odd x odd kernel (Tophat2DKernel, rotated beams), one
// thread per output, taps in shared memory, input through L1/L2.
#include "common.cuh"
#include "tma.cuh"
#include <stddef.h>

namespace scb {

int check_cube_args(const float *cube, int64_t nchan, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y);
int env_int(const char *name, int dflt);

constexpr int SP_TX = 128;               // strip width = compute threads
constexpr int SP_HP = 16;                // horizontal halo carried in the raw ring (>= H, multiple of 4)
constexpr int SP_W = SP_TX + 2 * SP_HP;  // raw row width in floats
constexpr int SP_R = 8;                  // rows per block
constexpr int SP_RS = 3;                 // raw ring stages
constexpr int SP_THREADS = SP_TX + 32;
constexpr int SP_MAX_TAPS = 2 * 16 + 1;

struct SpatialParams {
    const float *in;
    void *out;
    int64_t nchan, ny, nx;
    int64_t stride_c, stride_y, out_stride_c, out_stride_y;
    const float *halo_top, *halo_bot;    // (nchan, halo_rows, nx) filled rows or NULL
    int halo_rows;
    int strips_per_row;                  // strips across x
    int rows_per_cta;                    // rows of the shard one CTA marches over
    int chunks;                          // row chunks per (channel, strip)
    float fill;
    double ksum;                         // sum of the normalised 2-D kernel (~1)
    double ty[SP_MAX_TAPS], tx[SP_MAX_TAPS];     // normalised factors, centred in 2H+1, zero padded
    float tyf[SP_MAX_TAPS], txf[SP_MAX_TAPS];
    double tx_scaled[SP_MAX_TAPS];        // tx * 2^896 (the row pass widens float32 by bit placement)
    uint32_t qy[SP_MAX_TAPS], qx[SP_MAX_TAPS];    // round(t * 2^31): the denominator's exact integer factors (sparse kernel)
    unsigned long long qall;              // (sum qy) * (sum qx): the denominator with nothing missing
    uint32_t cqy[SP_MAX_TAPS + 1];        // prefix sums of qy: cqy[k] = qy[0] + .. + qy[k-1] (pipe kernel, blank blocks)
    uint32_t cqx[SP_MAX_TAPS + 1];        // prefix sums of qx (pipe kernel, runs of missing inputs in crowded blocks)
    uint32_t qsx;                         // sum qx
    double qscale31;                      // qall * 2^(896-31): out = top * qall / present; reciprocal of present >> 31, widened by bit placement
    float lo_closed, hi_closed;          // interval mask as a closed float32 interval
    // pipe kernel -> fix-up kernel: outputs whose window holds almost nothing valid (0 < present < fix_thresh) are redone
    // exactly; {count, overflow}, a bit per 32-column x 64-row tile of the output (set by the first column warp that flags an
    // output in it) and the list of (channel, first row, first column) of the tiles whose bit was set
    unsigned int *fix_state;
    uint32_t *fix_list;
    uint32_t *fix_bitmap;
    uint32_t fix_cap;
    int hpad;                            // half-width of the padded taps (2 hpad + 1 entries in ty / tx)
    int64_t fix_bands, fix_segs;         // tiles per column of the plane, per row
    unsigned long long fix_thresh;
    const uint8_t *passthrough;          // (nchan) 1 = copy the filled plane through; may be NULL
    const unsigned int *sel;             // {missing, total} of a sample of the cube, or NULL: picks the denominator strategy
    DevMask mask;
};

template <int NB>
struct SpatialSmem {
    double top[NB * SP_R][SP_TX];
    float bot[NB * SP_R][SP_TX];
    float ctr[NB * SP_R][SP_TX];
    float raw[SP_RS][SP_R][SP_W];
    uint64_t full[SP_RS];
    uint64_t empty[SP_RS];
};

__device__ __forceinline__ bool mask_include_rt(const DevMask &m, float v, int64_t c, int64_t y, int64_t x) {
    if (m.mode == MODE_NONE) return true;
    if (m.mode == MODE_INTERVAL) return (v > m.lo) & (v < m.hi);
    return eval_mask_generic(m.prog, v, c, y, x);
}

// v * 2^-896 as a double, exactly, for finite v (see spectral_smooth.cu); +-inf stays +-inf
__device__ __forceinline__ double place_scaled_sp(float v) {
    const int b = __float_as_int(v);
    const long long t = (long long)b << 29;
    int hi = (int)(t >> 32) & 0x8FFFFFFF;
    if ((b & 0x7F800000) == 0x7F800000) hi |= 0x7FF00000;
    return __hiloint2double(hi, (int)t);
}

// Strategy switch, decided on the device so that the call stays asynchronous: a sampling kernel classifies the
// 8-row x 128-column blocks of up to 16 planes; a block is CROWDED when more than 1/64 of its samples are missing but
// not all of them (the pipe kernel then scatters tens of deficits per block or builds them from runs of missing samples:
// measured 14.5 ms per shard at 1 % scattered NaNs, 22.6 at 3 %, 25 at 40 %; clean or blank blocks cost 8.5).  With more
// than a third of the blocks crowded the float32 denominator convolved alongside the numerator (sep_march_kernel, flat
// 17.5 ms) wins.  The round-1 rule counted missing SAMPLES (> 10 %): it sent the shards
// under config 4's blank frame (25 % missing, nearly all of it in blank blocks) to the march, 16.5 ms instead of 12.
// Both kernels are launched; the one not selected returns at once.
__device__ __forceinline__ bool sel_wants_march(const unsigned int *sel) {
    return (unsigned long long)sel[0] * 3ull > (unsigned long long)sel[1];
}

__device__ __forceinline__ void compute_bar() {          // barrier among the SP_TX compute threads only
    asm volatile("bar.sync 1, %0;" :: "n"(SP_TX) : "memory");
}

template <int H, int OUT64>
__global__ void __launch_bounds__(SP_THREADS)
sep_march_kernel(const __grid_constant__ SpatialParams p) {
    if (p.sel && !sel_wants_march(p.sel)) return;        // the sparse-denominator kernel serves this cube
    constexpr int NT = 2 * H + 1;
    constexpr int HB = (H + SP_R - 1) / SP_R;            // halo in blocks
    constexpr int NB = 2 * HB + 2;                       // ring of row-passed blocks
    constexpr int NIN_X = SP_R + 2 * SP_HP;              // 40 inputs per row-pass thread (aligned superset)
    constexpr int NIN_Y = SP_R + 2 * H;
    static_assert(H <= SP_HP, "horizontal halo too small");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SpatialSmem<NB> &sm = *reinterpret_cast<SpatialSmem<NB> *>(smem_raw);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // blockIdx.x = ((c * chunks) + chunk) * strips + strip
    int64_t bid = blockIdx.x;
    const int strip = (int)(bid % p.strips_per_row); bid /= p.strips_per_row;
    const int chunk = (int)(bid % p.chunks);
    const int64_t c = bid / p.chunks;
    const int64_t x0 = (int64_t)strip * SP_TX;
    const int64_t ya = (int64_t)chunk * p.rows_per_cta;
    const int64_t yb = min(p.ny, ya + p.rows_per_cta);
    const int nout_blk = (int)((yb - ya + SP_R - 1) / SP_R);
    const int nblk = nout_blk + 2 * HB;                  // blocks to row-pass
    const int64_t y_first = ya - (int64_t)HB * SP_R;     // first row of block 0

    // clipped horizontal window of the raw rows
    const int64_t xl = max((int64_t)0, x0 - SP_HP), xr = min(p.nx, x0 + SP_TX + SP_HP);
    const int col_off = (int)(xl - (x0 - SP_HP));        // where the copy lands in the raw row
    const uint32_t row_bytes = (uint32_t)(xr - xl) * 4u;

    if (tid == 0) {
        for (int s = 0; s < SP_RS; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], SP_TX / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == SP_TX / 32) {
        // ---------------- producer warp: raw rows of block b -> raw[b % RS] ----------------
        const uint64_t pol = l2_evict_first_policy();
        for (int b = 0; b < nblk; ++b) {
            const int s = b % SP_RS;
            if (b >= SP_RS) mbar_wait(&sm.empty[s], ((b / SP_RS) - 1) & 1);
            const float *src = nullptr;
            if (lane < SP_R) {
                const int64_t y = y_first + (int64_t)b * SP_R + lane;
                if (y >= 0 && y < p.ny) src = p.in + c * p.stride_c + y * p.stride_y + xl;
                else if (y < 0 && p.halo_top && y >= -p.halo_rows)
                    src = p.halo_top + (c * p.halo_rows + (p.halo_rows + y)) * p.nx + xl;
                else if (y >= p.ny && p.halo_bot && y < p.ny + p.halo_rows)
                    src = p.halo_bot + (c * p.halo_rows + (y - p.ny)) * p.nx + xl;
            }
            const unsigned have = __ballot_sync(0xffffffffu, src != nullptr);
            if (lane == 0) mbar_expect_tx(&sm.full[s], (uint32_t)__popc(have) * row_bytes);
            __syncwarp();
            if (src) tma_load_1d(&sm.raw[s][lane][col_off], src, row_bytes, &sm.full[s], pol);
        }
        return;
    }

    // ---------------- compute warps ----------------
    // float32 denominator of a window with nothing missing, accumulated in the same order as below
    float botx_full = 0.0f;
#pragma unroll
    for (int k = 0; k < NT; ++k) botx_full = fmaf(p.txf[k], 1.0f, botx_full);
    float bot_full = 0.0f;
#pragma unroll
    for (int k = 0; k < NT; ++k) bot_full = fmaf(p.tyf[k], botx_full, bot_full);
    // res = top / (ksum * bot / bot_full) = top * recip_scale * (1 / bot); the float reciprocal is widened
    // by bit placement (x 2^-896), which recip_scale undoes
    const double recip_scale = ((double)bot_full / p.ksum) * 0x1p+896;
    const bool pass = p.passthrough && p.passthrough[c];
    const bool strip_clipped = xl != x0 - SP_HP || xr != x0 + SP_TX + SP_HP;
    // masks that are not "closed interval with NaN fill" are applied by the in-place sweep
    const bool mask_by_sweep = p.mask.mode == MODE_GENERIC || (p.mask.mode == MODE_INTERVAL && p.fill == p.fill);

    const int rrow = tid >> 4;            // row-pass mapping: row of the block
    const int roct = tid & 15;            // ... and octet of columns

    for (int b = 0; b < nblk; ++b) {
        const int s = b % SP_RS;
        const int64_t yblk = y_first + (int64_t)b * SP_R;
        mbar_wait(&sm.full[s], (b / SP_RS) & 1);

        // (1) Only where needed (uniform per block): zero what the copy did not cover and apply masks that
        //     the row pass cannot fold into its validity test.  Interior blocks skip this and its barrier.
        const bool rows_inside = yblk >= (p.halo_top ? -(int64_t)p.halo_rows : 0) &&
                                 yblk + SP_R <= p.ny + (p.halo_bot ? (int64_t)p.halo_rows : 0);
        const bool sweep = strip_clipped || !rows_inside || mask_by_sweep;
        float lo_c = -INFINITY, hi_c = INFINITY;                     // closed validity interval of the row pass
        if (sweep) {
#pragma unroll 1
            for (int r = 0; r < SP_R; ++r) {
                const int64_t y = yblk + r;                          // uniform
                const bool own = y >= 0 && y < p.ny;
                const bool halo = (y < 0 && p.halo_top && y >= -p.halo_rows) || (y >= p.ny && p.halo_bot && y < p.ny + p.halo_rows);
                const bool need_mask = own && p.mask.mode != MODE_NONE;
                for (int col = tid; col < SP_W; col += SP_TX) {
                    const int64_t x = x0 - SP_HP + col;
                    float v = 0.0f;                                  // outside the image: a valid zero
                    if (x >= xl && x < xr && (own || halo)) {
                        v = sm.raw[s][r][col];
                        if (need_mask && !mask_include_rt(p.mask, v, c, y, x)) v = p.fill;
                    }
                    sm.raw[s][r][col] = v;
                }
            }
            compute_bar();
        } else if (p.mask.mode == MODE_INTERVAL) {
            lo_c = p.lo_closed; hi_c = p.hi_closed;                  // excluded == NaN-filled == "missing"
        }

        // (2) row pass: 8 adjacent outputs of row `rrow` from 40 aligned inputs
        {
            double w[NIN_X];
            float okf[NIN_X];
            bool anybad = false;
#pragma unroll
            for (int q = 0; q < NIN_X / 4; ++q) {
                const float4 f = *reinterpret_cast<const float4 *>(&sm.raw[s][rrow][roct * 8 + q * 4]);
                const float a[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                    const bool ok = (a[z] >= lo_c) & (a[z] <= hi_c);  // false for NaN
                    anybad |= !ok;
                    w[q * 4 + z] = place_scaled_sp(ok ? a[z] : 0.0f);
                    okf[q * 4 + z] = ok ? 1.0f : 0.0f;
                }
            }
            const int slot = (b % NB) * SP_R + rrow;
            double top[SP_R];
#pragma unroll
            for (int j = 0; j < SP_R; ++j) top[j] = 0.0;
#pragma unroll
            for (int k = 0; k < NT; ++k) {                           // tap-outer: each tap is fetched once
                const double t = p.tx_scaled[k];
#pragma unroll
                for (int j = 0; j < SP_R; ++j) top[j] = fma(t, w[j + SP_HP + H - k], top[j]);
            }
            float bot[SP_R];
#pragma unroll
            for (int j = 0; j < SP_R; ++j) bot[j] = botx_full;
            if (anybad) {
#pragma unroll
                for (int j = 0; j < SP_R; ++j) bot[j] = 0.0f;
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    const float t = p.txf[k];
#pragma unroll
                    for (int j = 0; j < SP_R; ++j) bot[j] = fmaf(t, okf[j + SP_HP + H - k], bot[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < SP_R; j += 2) {
                *reinterpret_cast<double2 *>(&sm.top[slot][roct * 8 + j]) = make_double2(top[j], top[j + 1]);
                *reinterpret_cast<float2 *>(&sm.bot[slot][roct * 8 + j]) = make_float2(bot[j], bot[j + 1]);
            }
            // centre values (filled input) for the bot == 0 / pass-through cases
#pragma unroll
            for (int j = 0; j < SP_R; ++j) {
                const float cv = sm.raw[s][rrow][SP_HP + roct * 8 + j];
                sm.ctr[slot][roct * 8 + j] = ((cv >= lo_c) & (cv <= hi_c)) || cv != cv || sweep ? cv : p.fill;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
        compute_bar();

        // (3) column pass for output block j = b - 2 HB (its rows are block j + HB of the march)
        if (b >= 2 * HB) {
            const int jb = b - HB;                                   // march-block holding the output rows
            const int64_t yout = y_first + (int64_t)jb * SP_R;
            double w[NIN_Y];
            float bt[NIN_Y];
            // ring slots of the 2 HB + 1 march-blocks this output block reads (block b - 2HB .. b)
            int slot_of[2 * HB + 1];
            {
                int sl = (b - 2 * HB) % NB;
#pragma unroll
                for (int t = 0; t < 2 * HB + 1; ++t) { slot_of[t] = sl * SP_R; sl = (sl + 1 == NB) ? 0 : sl + 1; }
            }
#pragma unroll
            for (int i = 0; i < NIN_Y; ++i) {
                const int rel = HB * SP_R - H + i;                   // compile-time: row relative to block b - 2HB
                const int slot = slot_of[rel / SP_R] + (rel % SP_R);
                w[i] = sm.top[slot][tid];
                bt[i] = sm.bot[slot][tid];
            }
            const int64_t x = x0 + tid;
            double top[SP_R];
            float bot[SP_R];
#pragma unroll
            for (int r = 0; r < SP_R; ++r) { top[r] = 0.0; bot[r] = 0.0f; }
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                const double t = p.ty[k];
                const float tf = p.tyf[k];
#pragma unroll
                for (int r = 0; r < SP_R; ++r) {
                    top[r] = fma(t, w[r + 2 * H - k], top[r]);
                    bot[r] = fmaf(tf, bt[r + 2 * H - k], bot[r]);
                }
            }
            char *op = reinterpret_cast<char *>(p.out) + (OUT64 ? 8 : 4) * (c * p.out_stride_c + yout * p.out_stride_y + x);
            const int64_t ostep = (OUT64 ? 8 : 4) * p.out_stride_y;
#pragma unroll
            for (int r = 0; r < SP_R; ++r) {
                if (yout + r < yb && x < p.nx) {
                    double res = top[r];
                    if (bot[r] != bot_full || pass) {
                        const float centre = sm.ctr[slot_of[HB] + r][tid];
                        if (bot[r] == 0.0f || pass) res = (double)centre;
                        else res = top[r] * recip_scale * place_scaled_sp(__frcp_rn(bot[r]));
                    }
                    if (OUT64) *reinterpret_cast<double *>(op + r * ostep) = res;
                    else       *reinterpret_cast<float *>(op + r * ostep) = (float)res;
                }
            }
        }
    }
}

// ================================================================================================
// sep_sparse_kernel -- the separable path for non-negative factors (every Gaussian).
//
// Same march as sep_march_kernel for the float64 numerator (row pass -> ring -> column pass, 2(2H+1)
// FMAs per voxel), but the DENOMINATOR sum_k K[k] ok[..] is no longer convolved: with zero padding
// counting as valid it equals the full kernel sum everywhere except under a missing input, and missing
// inputs are rare.  It is therefore kept as an exact integer DEFICIT: the factors are quantised once
// (qy, qx = round(t 2^31), so every product qy qx is an exact 62-bit integer and sums are order-free),
// each missing input found while a block is widened is put on a short list, and when an output block
// is finished every thread adds, for the listed inputs of the 2 HB + 1 blocks around it, qy[ky] qx[kx]
// to the deficits of its own 8 outputs in registers (zero-padded tables, no branches, no atomics:
// ~18 instructions per listed input against 2(2H+1) float FMAs per voxel for the convolution).  Blocks
// with more than SQ_CAP missing inputs (blanked frames, masked regions) switch to the same sum as a
// separable INTEGER convolution (row deficits in a ring, column gather), which adds exactly the same
// integers -- results do not depend on which path a block took.  present = qall - deficit == 0 is the exact
// "nothing valid under the kernel" test; otherwise out = top * qall / present through a float
// reciprocal (relative error ~1e-7, as before).
//
// Other changes: every input is widened to float64 ONCE into a padded shared-memory row (the row pass
// of sep_march_kernel widened each input five times); rows of the widened block and of the ring are
// padded by 16 bytes so that the eight threads of an LDS.128/STS.128 phase (eight different rows, same
// octet) hit eight different bank groups; the centre values for the rare "nothing valid" outputs are
// re-read from global memory instead of being carried in a third ring.
// ================================================================================================
constexpr int SQ_CAP = 8;                // missing inputs per block handled one by one
constexpr int SQ_FULL = SP_R * SP_W;     // a blank block: every sample of the 8 x 160 window is missing
constexpr int SQ_QXOFF = SP_W;           // kx = xo - cin + SP_HP + H lies in (-SP_W, SP_TX + 2 SP_HP)
constexpr int SQ_QXP = SP_W + SP_TX + 2 * SP_HP + 8;
constexpr int SQ_QYOFF = 32;             // ky = H - dy + ro lies in [H - 2 HB 8 - 7, H + 2 HB 8 + 7]
constexpr int SQ_QYP = 96;
constexpr int SQ_WD = SP_W + 2;          // padded row of the widened block (doubles)
constexpr int SQ_TOP = SP_TX + 2;        // padded ring row (doubles)

// Rare outputs of the sparse kernel, out of line: a plane that is copied through, nothing valid under the
// kernel (present == 0: the filled input), or almost nothing valid (present < 2^55: exact division).
__device__ __noinline__ double sparse_rare_output(const SpatialParams &p, double top, unsigned long long present, bool pass,
                                                  int64_t c, int64_t y, int64_t x) {
    if (present == 0ull || pass) {
        float cv = __ldg(p.in + c * p.stride_c + y * p.stride_y + x);
        if (!mask_include_rt(p.mask, cv, c, y, x)) cv = p.fill;
        return (double)cv;
    }
    return top * ((double)p.qall / (double)present);
}

template <int NB>
struct SparseSmem {
    double top[NB * SP_R][SQ_TOP];
    double wd[SP_R][SQ_WD];
    uint32_t dxi[NB * SP_R][SP_TX];
    float raw[SP_RS][SP_R][SP_W];
    uint32_t bad[SP_R][SP_W];                     // 1 = missing (words: the crowded row pass reads them as vectors)
    uint32_t qxp[SQ_QXP];                         // qx[k] at index k + SQ_QXOFF, zero elsewhere
    uint32_t qyp[SQ_QYP];                         // qy[k] at index k + SQ_QYOFF, zero elsewhere
    uint64_t full[SP_RS];
    uint64_t empty[SP_RS];
    int count[NB];
    uint16_t list[NB][SQ_CAP];
};
static_assert(offsetof(SparseSmem<6>, dxi) % 16 == 0 && offsetof(SparseSmem<6>, raw) % 16 == 0 &&
              offsetof(SparseSmem<6>, bad) % 16 == 0 && offsetof(SparseSmem<4>, bad) % 16 == 0, "vector accesses need 16-byte aligned members");

constexpr int SQ_CT = 256;               // compute threads: 8 warps share one 8 x 128 block
constexpr int SQ_THREADS = SQ_CT + 32;   // + the producer warp

__device__ __forceinline__ void sparse_bar() {           // barrier among the SQ_CT compute threads only
    asm volatile("bar.sync 1, %0;" :: "n"(SQ_CT) : "memory");
}

// Work split (256 compute threads per 8 x 128 block, half the serial work per warp of a 128-thread
// CTA -- the kernel is bound by the latency of its phases, not by issue slots):
//   widening   thread -> column tid & 127 of rows 2i + (tid >> 7), plus one sample of the right-hand halo;
//   row pass   thread -> row tid & 7 (fastest: eight rows = eight bank groups), columns 4 (tid >> 3) .. + 3;
//   column pass thread -> column tid & 127, output rows 4 (tid >> 7) .. + 3.
template <int H, int OUT64>
__global__ void __launch_bounds__(SQ_THREADS, 2)
sep_sparse_kernel(const __grid_constant__ SpatialParams p) {
    if (p.sel && sel_wants_march(p.sel)) return;         // too many missing samples: the convolved denominator is cheaper
    constexpr int NT = 2 * H + 1;
    constexpr int HB = (H + SP_R - 1) / SP_R;            // halo in blocks
    constexpr int NB = 2 * HB + 2;                       // ring of row-passed blocks
    constexpr int RQ = 4;                                // outputs per thread in both passes
    constexpr int NIN_X = RQ + 2 * SP_HP;                // 36 inputs per row-pass thread (aligned superset)
    constexpr int NIN_Y = RQ + 2 * H;
    static_assert(H <= SP_HP, "horizontal halo too small");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SparseSmem<NB> &sm = *reinterpret_cast<SparseSmem<NB> *>(smem_raw);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // blockIdx.x = ((c * chunks) + chunk) * strips + strip
    int64_t bid = blockIdx.x;
    const int strip = (int)(bid % p.strips_per_row); bid /= p.strips_per_row;
    const int chunk = (int)(bid % p.chunks);
    const int64_t c = bid / p.chunks;
    const int64_t x0 = (int64_t)strip * SP_TX;
    const int64_t ya = (int64_t)chunk * p.rows_per_cta;
    const int64_t yb = min(p.ny, ya + p.rows_per_cta);
    const int nout_blk = (int)((yb - ya + SP_R - 1) / SP_R);
    const int nblk = nout_blk + 2 * HB;                  // blocks to row-pass
    const int64_t y_first = ya - (int64_t)HB * SP_R;     // first row of block 0

    const int64_t xl = max((int64_t)0, x0 - SP_HP), xr = min(p.nx, x0 + SP_TX + SP_HP);
    const int col_off = (int)(xl - (x0 - SP_HP));
    const uint32_t row_bytes = (uint32_t)(xr - xl) * 4u;

    if (tid == 0) {
        for (int s = 0; s < SP_RS; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], SQ_CT / 32); }
        mbar_fence_init();
    }
    if (tid < NB) sm.count[tid] = 0;
    for (int i = tid; i < SQ_QXP; i += SQ_THREADS) { const int k = i - SQ_QXOFF; sm.qxp[i] = (k >= 0 && k < NT) ? p.qx[k] : 0u; }
    for (int i = tid; i < SQ_QYP; i += SQ_THREADS) { const int k = i - SQ_QYOFF; sm.qyp[i] = (k >= 0 && k < NT) ? p.qy[k] : 0u; }
    __syncthreads();

    if (warp == SQ_CT / 32) {
        // ---------------- producer warp: raw rows of block b -> raw[b % RS] ----------------
        const uint64_t pol = l2_evict_first_policy();
        for (int b = 0; b < nblk; ++b) {
            const int s = b % SP_RS;
            if (b >= SP_RS) mbar_wait(&sm.empty[s], ((b / SP_RS) - 1) & 1);
            const float *src = nullptr;
            if (lane < SP_R) {
                const int64_t y = y_first + (int64_t)b * SP_R + lane;
                if (y >= 0 && y < p.ny) src = p.in + c * p.stride_c + y * p.stride_y + xl;
                else if (y < 0 && p.halo_top && y >= -p.halo_rows)
                    src = p.halo_top + (c * p.halo_rows + (p.halo_rows + y)) * p.nx + xl;
                else if (y >= p.ny && p.halo_bot && y < p.ny + p.halo_rows)
                    src = p.halo_bot + (c * p.halo_rows + (y - p.ny)) * p.nx + xl;
            }
            const unsigned have = __ballot_sync(0xffffffffu, src != nullptr);
            if (lane == 0) mbar_expect_tx(&sm.full[s], (uint32_t)__popc(have) * row_bytes);
            __syncwarp();
            if (src) tma_load_1d(&sm.raw[s][lane][col_off], src, row_bytes, &sm.full[s], pol);
        }
        return;
    }

    // ---------------- compute warps ----------------
    const bool pass = p.passthrough && p.passthrough[c];
    const bool strip_clipped = xl != x0 - SP_HP || xr != x0 + SP_TX + SP_HP;
    const bool mask_by_sweep = p.mask.mode == MODE_GENERIC || (p.mask.mode == MODE_INTERVAL && p.fill == p.fill);
    const int rrow = tid & 7;             // row pass: row of the block ...
    const int rq = tid >> 3;              // ... and quad of columns
    const int ccol = tid & (SP_TX - 1);   // column pass: column ...
    const int chalf = tid >> 7;           // ... and half of the output block (rows 4 chalf .. + 3)

    for (int b = 0; b < nblk; ++b) {
        const int s = b % SP_RS;
        const int rb = b % NB;
        const int64_t yblk = y_first + (int64_t)b * SP_R;
        mbar_wait(&sm.full[s], (b / SP_RS) & 1);

        // (1) Only where needed (uniform per block): zero what the copy did not cover and apply masks that
        //     the validity test below cannot express.
        const bool rows_inside = yblk >= (p.halo_top ? -(int64_t)p.halo_rows : 0) &&
                                 yblk + SP_R <= p.ny + (p.halo_bot ? (int64_t)p.halo_rows : 0);
        const bool sweep = strip_clipped || !rows_inside || mask_by_sweep;
        float lo_c = -INFINITY, hi_c = INFINITY;                     // closed validity interval
        if (sweep) {
#pragma unroll 1
            for (int r = 0; r < SP_R; ++r) {
                const int64_t y = yblk + r;                          // uniform
                const bool own = y >= 0 && y < p.ny;
                const bool halo = (y < 0 && p.halo_top && y >= -p.halo_rows) || (y >= p.ny && p.halo_bot && y < p.ny + p.halo_rows);
                const bool need_mask = own && p.mask.mode != MODE_NONE;
                for (int col = tid; col < SP_W; col += SQ_CT) {
                    const int64_t x = x0 - SP_HP + col;
                    float v = 0.0f;                                  // outside the image: a valid zero
                    if (x >= xl && x < xr && (own || halo)) {
                        v = sm.raw[s][r][col];
                        if (need_mask && !mask_include_rt(p.mask, v, c, y, x)) v = p.fill;
                    }
                    sm.raw[s][r][col] = v;
                }
            }
            sparse_bar();
        } else if (p.mask.mode == MODE_INTERVAL) {
            lo_c = p.lo_closed; hi_c = p.hi_closed;                  // excluded == NaN-filled == "missing"
        }

        // (2) widen every input once; list the missing ones (all indices but tid are compile-time)
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int row = i < 4 ? 2 * i + chalf : warp;
            const int col = i < 4 ? ccol : SP_TX + lane;
            const float v = sm.raw[s][row][col];
            const bool ok = (v >= lo_c) & (v <= hi_c);               // false for NaN
            sm.wd[row][col] = place_scaled_sp(ok ? v : 0.0f);
            sm.bad[row][col] = ok ? 0u : 1u;
            // count (and list) the missing ones: one shared-memory atomic per warp, not per sample -- a blank
            // block has 1280 of them
            const unsigned missing = __ballot_sync(0xffffffffu, !ok);
            if (missing != 0u) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&sm.count[rb], __popc(missing));
                base = __shfl_sync(0xffffffffu, base, 0);
                const int idx = base + __popc(missing & ((1u << lane) - 1u));
                if (!ok && idx < SQ_CAP) sm.list[rb][idx] = (uint16_t)((row << 8) | col);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);                    // the raw rows are not read again
        sparse_bar();
        // next block's counter: its old block was last read in step b - 1, which every warp has left by now
        if (tid == 0) sm.count[(b + 1) % NB] = 0;

        // (3) row pass: 4 adjacent outputs of row `rrow` from 36 widened inputs
        const int cnt_b = sm.count[rb];                              // uniform
        const int slot = rb * SP_R + rrow;
        if (cnt_b == SQ_FULL) {
            // a blank block (every sample missing): the numerator rows are zero, the deficit is known in closed form
            *reinterpret_cast<double2 *>(&sm.top[slot][rq * RQ]) = make_double2(0.0, 0.0);
            *reinterpret_cast<double2 *>(&sm.top[slot][rq * RQ + 2]) = make_double2(0.0, 0.0);
        } else {
            double w[NIN_X];
#pragma unroll
            for (int q = 0; q < NIN_X / 2; ++q) {
                const double2 d2 = *reinterpret_cast<const double2 *>(&sm.wd[rrow][rq * RQ + q * 2]);
                w[q * 2] = d2.x; w[q * 2 + 1] = d2.y;
            }
            double top[RQ];
#pragma unroll
            for (int j = 0; j < RQ; ++j) top[j] = 0.0;
#pragma unroll
            for (int k = 0; k < NT; ++k) {                           // tap-outer: each tap is fetched once
                const double t = p.tx_scaled[k];
#pragma unroll
                for (int j = 0; j < RQ; ++j) top[j] = fma(t, w[j + SP_HP + H - k], top[j]);
            }
#pragma unroll
            for (int j = 0; j < RQ; j += 2)
                *reinterpret_cast<double2 *>(&sm.top[slot][rq * RQ + j]) = make_double2(top[j], top[j + 1]);
            if (cnt_b > SQ_CAP) {
                // crowded block: row deficits as an integer convolution of the missing flags
                uint32_t f[NIN_X];
#pragma unroll
                for (int q = 0; q < NIN_X / 4; ++q) {
                    const uint4 u = *reinterpret_cast<const uint4 *>(&sm.bad[rrow][rq * RQ + q * 4]);
                    f[q * 4] = u.x; f[q * 4 + 1] = u.y; f[q * 4 + 2] = u.z; f[q * 4 + 3] = u.w;
                }
                uint32_t dx[RQ];
#pragma unroll
                for (int j = 0; j < RQ; ++j) dx[j] = 0u;
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    const uint32_t q = p.qx[k];
#pragma unroll
                    for (int j = 0; j < RQ; ++j) dx[j] += q * f[j + SP_HP + H - k];
                }
                *reinterpret_cast<uint4 *>(&sm.dxi[slot][rq * RQ]) = make_uint4(dx[0], dx[1], dx[2], dx[3]);
            }
        }
        sparse_bar();

        // (4) output block j = b - 2 HB (its rows are block b - HB of the march)
        if (b >= 2 * HB) {
            const int jb = b - HB;
            const int64_t yout = y_first + (int64_t)jb * SP_R + chalf * RQ;      // this thread's first output row
            // ---- deficit of this thread's 4 outputs from the 2 HB + 1 input blocks around them:
            //      exact integers in registers, no atomics ----
            unsigned long long def[RQ];
#pragma unroll
            for (int ro = 0; ro < RQ; ++ro) def[ro] = 0ull;
            bool all_blank = true;                                   // uniform: every block around the outputs is blank
#pragma unroll
            for (int t = 0; t < 2 * HB + 1; ++t) {
                const int st = (b - 2 * HB + t) % NB;
                const int cnt = sm.count[st];                        // uniform
                all_blank &= cnt == SQ_FULL;
                if (cnt == 0) continue;
                if (cnt == SQ_FULL) {
                    // blank block: every row contributes (sum of qx) x qy[ky]
#pragma unroll
                    for (int r_in = 0; r_in < SP_R; ++r_in) {
                        const uint32_t *qyr = &sm.qyp[H - ((t - HB) * SP_R + r_in) + chalf * RQ + SQ_QYOFF];
#pragma unroll
                        for (int ro = 0; ro < RQ; ++ro) def[ro] += (unsigned long long)qyr[ro] * p.qsx;
                    }
                } else if (cnt <= SQ_CAP) {
                    // few: one listed input at a time; tables padded with zeros make every product valid
#pragma unroll 1
                    for (int n = 0; n < cnt; ++n) {
                        const int ent = sm.list[st][n];              // uniform
                        const int r_in = ent >> 8, cin = ent & 255;
                        const uint32_t qxv = sm.qxp[ccol - cin + SP_HP + H + SQ_QXOFF];                          // qx[kx], kx = xo - xi + H
                        const uint32_t *qyr = &sm.qyp[H - ((t - HB) * SP_R + r_in) + chalf * RQ + SQ_QYOFF];     // qy[ky], ky = H - (yi - yo)
#pragma unroll
                        for (int ro = 0; ro < RQ; ++ro) def[ro] += (unsigned long long)qyr[ro] * qxv;
                    }
                } else {
                    // crowded: gather this block's row deficits down the column
#pragma unroll
                    for (int r_in = 0; r_in < SP_R; ++r_in) {
                        const uint32_t d = sm.dxi[st * SP_R + r_in][ccol];
                        const uint32_t *qyr = &sm.qyp[H - ((t - HB) * SP_R + r_in) + chalf * RQ + SQ_QYOFF];
#pragma unroll
                        for (int ro = 0; ro < RQ; ++ro) def[ro] += (unsigned long long)d * qyr[ro];      // one 32 x 32 + 64 multiply-add
                    }
                }
            }

            // ---- column pass: 4 outputs from 4 + 2H ring rows starting at ring row r0 ----
            const int64_t x = x0 + ccol;
            double top[RQ];
#pragma unroll
            for (int r = 0; r < RQ; ++r) top[r] = 0.0;
            if (!all_blank) {                                        // (a blank neighbourhood has present == 0: no numerator needed)
                double w[NIN_Y];
                int rr = ((b - 2 * HB) % NB) * SP_R + HB * SP_R - H + chalf * RQ;     // ring row of w[0]
                if (rr >= NB * SP_R) rr -= NB * SP_R;
#pragma unroll
                for (int i = 0; i < NIN_Y; ++i) {
                    w[i] = sm.top[rr][ccol];
                    rr = (rr + 1 == NB * SP_R) ? 0 : rr + 1;
                }
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    const double t = p.ty[k];
#pragma unroll
                    for (int r = 0; r < RQ; ++r) top[r] = fma(t, w[r + 2 * H - k], top[r]);
                }
            }
            char *op = reinterpret_cast<char *>(p.out) + (OUT64 ? 8 : 4) * (c * p.out_stride_c + yout * p.out_stride_y + x);
            const int64_t ostep = (OUT64 ? 8 : 4) * p.out_stride_y;
            // out = top * qall / present.  Common path without branches: float reciprocal of the top 32 bits of
            // `present` (exact to 2^-24 when present >= 2^55), widened by bit placement; a zero deficit keeps
            // `top` as it is.  Rare, out of line: a plane copied through, nothing valid (present == 0: the
            // filled input) or almost nothing valid (present < 2^55: exact division).
#pragma unroll
            for (int r = 0; r < RQ; ++r) {
                const unsigned long long present = p.qall - def[r];
                const uint32_t hi = (uint32_t)(present >> 31);
                const double rcp = place_scaled_sp(__frcp_rn((float)hi)) * p.qscale31;
                double res = def[r] != 0ull ? top[r] * rcp : top[r];
                const bool live = yout + r < yb && x < p.nx;
                if ((hi < (1u << 24) || pass) && live) res = sparse_rare_output(p, top[r], present, pass, c, yout + r, x);
                if (live) {
                    if (OUT64) *reinterpret_cast<double *>(op + r * ostep) = res;
                    else       *reinterpret_cast<float *>(op + r * ostep) = (float)res;
                }
            }
        }
    }
}

// ================================================================================================
// sep_pipe_kernel -- the separable path as a warp-specialised pipeline (round 2; replaces sep_sparse_kernel).
//
// Why: ncu on sep_sparse_kernel showed ~290 thread-instructions per voxel for the 58 DFMAs a 29 x 29 separable
// kernel needs -- the FP64 pipe (2 issue cycles per DFMA) sat at 27 %.  The budget of an FP64-bound kernel is ONE
// other instruction per DFMA.  This kernel gets there by (a) 16 outputs per thread in BOTH passes, inputs streamed
// through registers once (input-outer order: a loaded value feeds up to 16 independent accumulator chains), (b) taps
// held in registers for the whole march (DFMA with a constant-bank operand issues 10 % slower), (c) no phase
// barriers: row warps and column warps are different warps that meet through mbarriers on a ring of row-passed
// blocks, (d) compile-time ring offsets from three base pointers instead of modular index arithmetic, (e) the
// integer deficit of the denominator kept as a separable quantity: missing inputs scatter qx[.] into a 32-bit row
// plane with shared-memory atomics (29 per missing input), the column warps fold only the DIRTY rows with qy[.].
//
// CTA = one 128-column strip of one channel marching down the rows in blocks of 16; 9 warps:
//   warp 8      producer: raw rows (strip + 16-column halo) of block b -> raw[b % 2] with cp.async.bulk (TMA);
//   warps 0-3   row warps: widen the block to float64 once (bit placement, missing -> 0, listed), row pass
//               (thread = row r, 16-column segment g) -> top[b % 4], row deficits -> rowdef[b % 4];
//   warps 4-7   column warps: output block b - 1 from ring blocks b - 2 .. b (thread = column), deficit, epilogue, store.
// One CTA per SM (166 KB of shared memory, up to 224 registers): one row warp and one column warp per scheduler.
// Denominator semantics are those of sep_sparse_kernel (exact integers, zero padding valid, present == 0 <=> nothing
// valid); blocks with more than PP_CAP missing inputs compute their row deficits densely, blank blocks in closed form.
// ================================================================================================
constexpr int PP_R = 16;                   // rows per block = outputs per thread in both passes
constexpr int PP_RS = 2;                   // raw ring stages
constexpr int PP_NB = 4;                   // ring of row-passed blocks
constexpr int PP_WD = SP_W + 2;            // padded widened row (doubles): the 8 rows of an LDS.128 phase hit 8 bank groups
constexpr int PP_TOP = SP_TX + 2;          // padded ring row (doubles)
constexpr int PP_CAP = 64;                 // listed missing inputs per block; more -> dense integer row deficits
constexpr int PP_FULL = PP_R * SP_W;       // every sample of the block's window missing
constexpr int PP_DEFAULT_J = 8;            // outputs per thread in both passes (see sep_pipe_kernel): 8.4 ms against 9.3-10.6 ms with 16
enum { PM_CLEAN = 0, PM_SPARSE = 1, PM_CROWDED = 2, PM_BLANK = 3 };

struct PipeSmem {
    double top[PP_NB * PP_R][PP_TOP];
    double wd[2][PP_R][PP_WD];
    uint32_t rowdef[PP_NB * PP_R][SP_TX];
    float raw[PP_RS][PP_R][SP_W];
    uint32_t qyp[96];                      // qy[k] at index k + 32, zero elsewhere (the column warps' deficit fold)
    uint32_t cqxp[160];                    // cqx[clamp(t - 64, 0, NT)] at index t (runs of missing inputs, crowded blocks)
    uint32_t rowmask[PP_R][8];             // crowded blocks: bit c of row r = window sample (r, c) is missing (5 words used)
    uint16_t list[3][PP_CAP];
    int count[3];
    int mode[PP_NB];
    uint32_t dirty[PP_NB][4];              // per 32-column segment: bit r = row r of the block holds non-zero row deficits there
    uint64_t full[PP_RS], empty[PP_RS], rfull[PP_NB], rempty[PP_NB];
};
static_assert(offsetof(PipeSmem, wd) % 16 == 0 && offsetof(PipeSmem, rowdef) % 16 == 0 && offsetof(PipeSmem, raw) % 16 == 0 &&
              offsetof(PipeSmem, qyp) % 16 == 0, "vector accesses need 16-byte aligned members");

template <int NTHR>
__device__ __forceinline__ void pipe_bar(int id) {          // barrier among the row threads
    asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(NTHR) : "memory");
}

// a register copy of a kernel parameter that the compiler cannot fold back into a constant-bank operand
__device__ __forceinline__ double reg_copy(double v) {
    double r;
    asm volatile("mov.b64 %0, %1;" : "=d"(r) : "d"(v));
    return r;
}

struct PipeBlock {                         // what the out-of-line paths need to know about the block being processed
    int64_t c, yblk, x0, xl, xr;
};

// the filled sample at window position (row, col) of the block: zero outside the image (a valid sample), the
// neighbouring shard's filled row in the halo, the masked-out value replaced by the fill value
__device__ __noinline__ float pipe_filled(const SpatialParams &p, const float (*raw)[SP_W], int row, int col, const PipeBlock &k) {
    const int64_t y = k.yblk + row, x = k.x0 - SP_HP + col;
    const bool own = y >= 0 && y < p.ny;
    const bool halo = (y < 0 && p.halo_top && y >= -p.halo_rows) || (y >= p.ny && p.halo_bot && y < p.ny + p.halo_rows);
    float v = 0.0f;
    if (x >= k.xl && x < k.xr && (own || halo)) {
        v = raw[row][col];
        if (own && p.mask.mode != MODE_NONE && !mask_include_rt(p.mask, v, k.c, y, x)) v = p.fill;
    }
    return v;
}

// v * 2^-896 as a double, exactly, for FINITE v (no inf / NaN special case)
__device__ __forceinline__ double place_scaled_finite(float v) {
    const int b = __float_as_int(v);
    return __hiloint2double((b >> 3) & 0x8FFFFFFF, b << 29);
}

// (1, fast) widen an interior block: 16 rows x 80 pairs, warp w takes rows RW w .. RW w + RW - 1 (RW = 4 with four row
// warps, 2 with eight); returns the missing-element bits (bit 2i + e: element e of this thread's pair i)
template <bool FINITE, int RW>
__device__ __forceinline__ uint32_t pipe_widen_fast(PipeSmem &sm, int s, int par, int warp, int lane, float lo_c, float hi_c) {
    uint32_t miss = 0u;
#pragma unroll
    for (int i = 0; i < RW * 5 / 2; ++i) {
        // passes 0 .. 2RW-1: row RW w + i/2, pairs 32 (i&1) + lane; the RW/2 passes after them: the last 16 pairs of two rows each
        const int row = i < 2 * RW ? RW * warp + (i >> 1) : RW * warp + 2 * (i - 2 * RW) + (lane >> 4);
        const int pc = i < 2 * RW ? 32 * (i & 1) + lane : 64 + (lane & 15);
        const float2 v = *reinterpret_cast<const float2 *>(&sm.raw[s][row][2 * pc]);
        const bool ok0 = (v.x >= lo_c) & (v.x <= hi_c);           // false for NaN
        const bool ok1 = (v.y >= lo_c) & (v.y <= hi_c);
        const float w0 = ok0 ? v.x : 0.0f, w1 = ok1 ? v.y : 0.0f;
        *reinterpret_cast<double2 *>(&sm.wd[par][row][2 * pc]) =
            FINITE ? make_double2(place_scaled_finite(w0), place_scaled_finite(w1)) : make_double2(place_scaled_sp(w0), place_scaled_sp(w1));
        miss |= (ok0 ? 0u : 1u << (2 * i)) | (ok1 ? 0u : 2u << (2 * i));
    }
    return miss;
}

__device__ __forceinline__ void pipe_pair_of(int i, int rw, int warp, int lane, int &row, int &pc) {
    row = i < 2 * rw ? rw * warp + (i >> 1) : rw * warp + 2 * (i - 2 * rw) + (lane >> 4);
    pc = i < 2 * rw ? 32 * (i & 1) + lane : 64 + (lane & 15);
}

// (1, slow) widen a block that touches the image border, halo rows or a mask the interval test cannot express
__device__ __noinline__ void pipe_widen_slow(const SpatialParams &p, PipeSmem &sm, int s, int par, int l3, int rw, int warp, int lane,
                                             const PipeBlock &k) {
#pragma unroll 1
    for (int i = 0; i < rw * 5 / 2; ++i) {
        int row, pc;
        pipe_pair_of(i, rw, warp, lane, row, pc);
        const float vx = pipe_filled(p, sm.raw[s], row, 2 * pc, k);
        const float vy = pipe_filled(p, sm.raw[s], row, 2 * pc + 1, k);
        const bool ok0 = vx == vx, ok1 = vy == vy;
        *reinterpret_cast<double2 *>(&sm.wd[par][row][2 * pc]) = make_double2(place_scaled_sp(ok0 ? vx : 0.0f), place_scaled_sp(ok1 ? vy : 0.0f));
        if (!ok0) { const int idx = atomicAdd(&sm.count[l3], 1); if (idx < PP_CAP) sm.list[l3][idx] = (uint16_t)((row << 8) | (2 * pc)); }
        if (!ok1) { const int idx = atomicAdd(&sm.count[l3], 1); if (idx < PP_CAP) sm.list[l3][idx] = (uint16_t)((row << 8) | (2 * pc + 1)); }
    }
}

// (4) crowded block (more than PP_CAP missing inputs: blank frames, masked regions): the row deficits of one thread's J
// outputs from the RUNS of missing inputs in its window -- a run [i0, i1] adds qx[j + 2H - i1] + .. + qx[j + 2H - i0] to output j,
// two look-ups in the clamped prefix-sum table whatever its length.  Step A (all row threads, then a barrier): per-row missing
// bits by warp ballot; step B: this thread's J + 2H window bits, run by run.
__device__ __noinline__ void pipe_crowded_masks(const SpatialParams &p, PipeSmem &sm, int s, int rw, int warp, int lane,
                                                bool slow, float lo_c, float hi_c, const PipeBlock &k) {
#pragma unroll 1
    for (int rr = 0; rr < rw; ++rr) {
        const int row = rw * warp + rr;
#pragma unroll 1
        for (int q = 0; q < SP_W / 32; ++q) {
            const int col = 32 * q + lane;
            const float v = slow ? pipe_filled(p, sm.raw[s], row, col, k) : sm.raw[s][row][col];
            const unsigned word = __ballot_sync(0xffffffffu, !((v >= lo_c) & (v <= hi_c)));
            if (lane == 0) sm.rowmask[row][q] = word;
        }
        if (lane < 3) sm.rowmask[row][SP_W / 32 + lane] = 0u;
    }
}

template <int J>
__device__ __noinline__ void pipe_crowded_rowdef(PipeSmem &sm, int slot, int rrow, int seg, int h) {
    uint32_t dx[J];
#pragma unroll
    for (int j = 0; j < J; ++j) dx[j] = 0u;
    const int nin = J + 2 * h;
    const int w0 = seg * J + SP_HP - h;                           // first window column
    const uint32_t *m = &sm.rowmask[rrow][w0 >> 5];
    const int sh = w0 & 31;
    const uint32_t lo = __funnelshift_r(m[0], m[1], sh), hi = __funnelshift_r(m[1], m[2], sh);
    unsigned long long W = ((unsigned long long)hi << 32 | lo) & ((1ull << nin) - 1ull);
#pragma unroll 1
    while (W) {
        const int i0 = __ffsll((long long)W) - 1;
        const unsigned long long rest = ~(W >> i0);
        const int len = __ffsll((long long)rest) - 1;                 // trailing ones of W >> i0 (at least 1, at most 48)
        const int i1 = i0 + len - 1;
        W &= ~(((1ull << len) - 1ull) << i0);
        // sum over i in [i0, i1] of qx[j + 2h - i] = cqx[j + 2h - i0 + 1] - cqx[j + 2h - i1], indices clamped to [0, nt]
        const uint32_t *hi_t = &sm.cqxp[64 + 2 * h - i0 + 1], *lo_t = &sm.cqxp[64 + 2 * h - i1];
#pragma unroll
        for (int j = 0; j < J; ++j) dx[j] += hi_t[j] - lo_t[j];
    }
#pragma unroll
    for (int j = 0; j < J; j += 4)
        *reinterpret_cast<uint4 *>(&sm.rowdef[slot * PP_R + rrow][seg * J + j]) = make_uint4(dx[j], dx[j + 1], dx[j + 2], dx[j + 3]);
}

// The integer denominator carries the taps quantised to 2^-31.  That is exact enough (< 1e-6) whenever a tap of ordinary size
// takes part, but an output just inside a blank region -- only the outermost one or two kernel columns / rows reach valid
// data, present / qall < 2^-13 -- hangs on taps of ~2e-5 whose quantum is 1e-5 of THEM: the weighted mean then misses the
// reference by up to 2e-5.  Such outputs (two pixels deep along the edges of blank regions) are stored as a NaN with this
// payload and listed tile by tile; sep_fixup_kernel recomputes them from the input with the float64 taps.
constexpr uint32_t PP_SENTINEL32 = 0x7FC5CB20u;
constexpr unsigned long long PP_SENTINEL64 = 0x7FF85CB200000000ull;
constexpr uint32_t PP_FIX_CAP = 1u << 19;

constexpr int FX_TX = 32, FX_MAXROWS = 64, FX_W = FX_TX + 2 * SP_HP, FX_H = FX_MAXROWS + 2 * SP_HP, FX_THREADS = 128;
constexpr int FX_WARPS = FX_THREADS / 32, FX_BATCH = 6;                   // window rows a warp has in flight
constexpr size_t PP_FIX_BITMAP_BYTES = (size_t)4 << 20;                   // 2^25 tiles = 6.9e10 voxels: more than fits the HBM

// Rows [y, y + n) x the 32 columns from xs hold a flagged output: list the 64-row tile(s) they fall into, once each.
__device__ __forceinline__ void pipe_list_tile(const SpatialParams &p, int64_t c, int64_t y, int n, int64_t xs) {
    const int64_t seg = xs / FX_TX;
    for (int64_t band = y / FX_MAXROWS; band <= (y + n - 1) / FX_MAXROWS; ++band) {
        const int64_t t = (c * p.fix_bands + band) * p.fix_segs + seg;
        if (t >= (int64_t)(PP_FIX_BITMAP_BYTES * 8)) { p.fix_state[1] = 1u; return; }
        const uint32_t bit = 1u << (t & 31);
        if (atomicOr(&p.fix_bitmap[t >> 5], bit) & bit) continue;         // some other warp listed the tile
        const unsigned int at = atomicAdd(&p.fix_state[0], 1u);
        if (at < p.fix_cap) *reinterpret_cast<uint4 *>(p.fix_list + 4 * (size_t)at) = make_uint4((uint32_t)c, (uint32_t)(band * FX_MAXROWS), (uint32_t)(seg * FX_TX), 0u);
        else p.fix_state[1] = 1u;                                         // list full: the fix-up scans the whole output
    }
}

// One CTA per listed tile.  The tile's input window is staged ONCE in shared memory with a validity bit mask per window row;
// the flagged outputs are compacted and each one, four lanes on it, walks ONLY the valid samples of its window with the float64
// factors -- the exact weighted mean the reference computes (a flagged output has few valid samples by construction).
// History (config-4 interior shard, 1 million flagged outputs along the blank side columns): a warp per output reading global
// memory 20 ms; a thread per output position over a staged window 10 ms (5 active threads per instruction); compaction and
// valid-sample walk 2.1 ms; source row resolved per row instead of per sample 1.6 ms; tiles 64 rows tall instead of 8
// (the 28 halo rows of the window are amortised) and a rolled staging loop that fits the instruction cache: see DESIGN.md.
template <int OUT64>
__device__ __forceinline__ bool fixup_is_flagged(const void *out, int64_t at) {
    if (OUT64) return reinterpret_cast<const unsigned long long *>(out)[at] == PP_SENTINEL64;
    return reinterpret_cast<const uint32_t *>(out)[at] == PP_SENTINEL32;
}

template <int OUT64>
__global__ void __launch_bounds__(FX_THREADS)
sep_fixup_kernel(const __grid_constant__ SpatialParams p) {
    if (p.sel && sel_wants_march(p.sel)) return;
    const unsigned int listed = p.fix_state[0];
    const bool overflow = p.fix_state[1] != 0u;
    if (listed == 0u && !overflow) return;                                // the usual case: nothing to redo
    __shared__ float win[FX_H][FX_W + 1];
    __shared__ unsigned long long rmask[FX_H];
    __shared__ unsigned short flist[FX_MAXROWS * FX_TX];
    __shared__ int nflag, oy_lo, oy_hi;
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int h = p.hpad, nt = 2 * h + 1;
    const int64_t segs = (p.nx + FX_TX - 1) / FX_TX, bands = (p.ny + FX_MAXROWS - 1) / FX_MAXROWS;
    const int64_t ntiles = overflow ? p.nchan * bands * segs : (int64_t)min(listed, p.fix_cap);
    const float qnan = __int_as_float(0x7fc00000);
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int64_t c, y0, xs;
        if (overflow) {
            xs = (tile % segs) * FX_TX; const int64_t cb = tile / segs; y0 = (cb % bands) * FX_MAXROWS; c = cb / bands;
        } else {
            const uint4 e = *reinterpret_cast<const uint4 *>(p.fix_list + 4 * (size_t)tile);
            c = e.x; y0 = e.y; xs = e.z;
        }
        const int rows = (int)min((int64_t)FX_MAXROWS, p.ny - y0);
        if (tid == 0) { nflag = 0; oy_lo = FX_MAXROWS; oy_hi = -1; }
        __syncthreads();                                                  // (also orders the previous tile's reads of `win`)
        if (xs + lane < p.nx) {
            const int64_t obase = c * p.out_stride_c + y0 * p.out_stride_y + xs + lane;
            int lo = FX_MAXROWS, hi = -1;
            for (int oy = wrp; oy < rows; oy += FX_WARPS)
                if (fixup_is_flagged<OUT64>(p.out, obase + oy * p.out_stride_y)) {
                    flist[atomicAdd(&nflag, 1)] = (unsigned short)(oy * FX_TX + lane);
                    lo = min(lo, oy); hi = max(hi, oy);
                }
            if (hi >= 0) { atomicMin(&oy_lo, lo); atomicMax(&oy_hi, hi); }
        }
        __syncthreads();
        const int nf = nflag;
        if (nf == 0) continue;
        const int ylo = oy_lo;                                            // only the rows the flagged outputs reach are staged
        const int64_t ytop = y0 + ylo - h;                                // image row of window row 0
        {
            // stage the window a row per warp pass, FX_BATCH rows in flight; the ballots give each row's validity mask (the
            // window is at most 64 wide).  The row decides where a sample comes from (image, a neighbour shard's halo rows,
            // zero padding -- a valid sample): resolved once per row, not once per sample.
            const int nrows = oy_hi - ylo + 1 + 2 * h;
            const int64_t xa = xs - h + lane, xb = xa + 32;
            const bool ina = xa >= 0 && xa < p.nx, inb = lane < 2 * h && xb < p.nx;       // (xb >= 0 always: h <= 16)
            const float *plane = p.in + c * p.stride_c;
            const int mode = p.mask.mode;
#pragma unroll 1
            for (int r0 = wrp; r0 < nrows; r0 += FX_WARPS * FX_BATCH) {
                float v0[FX_BATCH], v1[FX_BATCH];
#pragma unroll
                for (int i = 0; i < FX_BATCH; ++i) {
                    const int r = r0 + FX_WARPS * i;
                    const int64_t yy = ytop + r;
                    const float *src = nullptr;
                    if (r < nrows) {
                        if (yy >= 0 && yy < p.ny) src = plane + yy * p.stride_y;
                        else if (yy < 0 && p.halo_top && yy >= -p.halo_rows) src = p.halo_top + (c * p.halo_rows + (p.halo_rows + yy)) * p.nx;
                        else if (yy >= p.ny && p.halo_bot && yy < p.ny + p.halo_rows) src = p.halo_bot + (c * p.halo_rows + (yy - p.ny)) * p.nx;
                    }
                    v0[i] = (src && ina) ? __ldg(src + xa) : 0.0f;
                    v1[i] = (src && inb) ? __ldg(src + xb) : 0.0f;
                }
                if (mode == MODE_INTERVAL) {
#pragma unroll
                    for (int i = 0; i < FX_BATCH; ++i) {
                        const int64_t yy = ytop + r0 + FX_WARPS * i;
                        if (yy >= 0 && yy < p.ny) {                       // (halo rows arrive filled)
                            if (ina && !((v0[i] > p.mask.lo) & (v0[i] < p.mask.hi))) v0[i] = p.fill;
                            if (inb && !((v1[i] > p.mask.lo) & (v1[i] < p.mask.hi))) v1[i] = p.fill;
                        }
                    }
                } else if (mode != MODE_NONE) {
#pragma unroll 1
                    for (int i = 0; i < FX_BATCH; ++i) {
                        const int64_t yy = ytop + r0 + FX_WARPS * i;
                        float a = 0.0f, b2 = 0.0f;
#pragma unroll
                        for (int k = 0; k < FX_BATCH; ++k) if (k == i) { a = v0[k]; b2 = v1[k]; }
                        if (yy >= 0 && yy < p.ny) {
                            if (ina && !eval_mask_generic(p.mask.prog, a, c, yy, xa)) a = p.fill;
                            if (inb && !eval_mask_generic(p.mask.prog, b2, c, yy, xb)) b2 = p.fill;
                        }
#pragma unroll
                        for (int k = 0; k < FX_BATCH; ++k) if (k == i) { v0[k] = a; v1[k] = b2; }
                    }
                }
#pragma unroll
                for (int i = 0; i < FX_BATCH; ++i) {
                    const int r = r0 + FX_WARPS * i;
                    if (r < nrows) {                                      // (warp-uniform)
                        const float w1 = lane < 2 * h ? v1[i] : qnan;     // columns past the window: never valid
                        win[r][lane] = v0[i]; win[r][32 + lane] = w1;
                        const unsigned m0 = __ballot_sync(0xffffffffu, v0[i] == v0[i]), m1 = __ballot_sync(0xffffffffu, w1 == w1);
                        if (lane == 0) rmask[r] = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
                    }
                }
            }
        }
        __syncthreads();
        // four lanes per output, kernel rows dealt round-robin, two shuffles to combine
        const int sub = tid & 3;
        const unsigned long long span = (1ull << nt) - 1ull;              // nt <= 33
        for (int q0 = 0; q0 < nf; q0 += FX_THREADS / 4) {
            const int q = q0 + (tid >> 2);
            const bool active = q < nf;
            const int o = active ? flist[q] : 0, oy = o / FX_TX, ox = o % FX_TX;
            double top = 0.0, bot = 0.0;
            if (active) {
                for (int ky = sub; ky < nt; ky += 4) {
                    const int r = oy - ylo + 2 * h - ky;                  // out(y, x) reads in(y + h - ky, x + h - kx)
                    unsigned long long bits = (rmask[r] >> ox) & span;    // bit j: window column ox + j, tap kx = 2 h - j
                    if (bits == 0ull) continue;
                    double t = 0.0, b = 0.0;
                    while (bits) {
                        const int j = __ffsll((long long)bits) - 1;
                        bits &= bits - 1ull;
                        const double w = p.tx[2 * h - j];
                        t = fma(w, (double)win[r][ox + j], t); b += w;
                    }
                    const double wy = p.ty[ky];
                    top = fma(wy, t, top); bot = fma(wy, b, bot);
                }
            }
            top += __shfl_xor_sync(0xffffffffu, top, 1); bot += __shfl_xor_sync(0xffffffffu, bot, 1);
            top += __shfl_xor_sync(0xffffffffu, top, 2); bot += __shfl_xor_sync(0xffffffffu, bot, 2);
            if (active && sub == 0) {
                const double res = bot == 0.0 ? (double)win[oy - ylo + h][ox + h] : top / bot;
                const int64_t at = c * p.out_stride_c + (y0 + oy) * p.out_stride_y + xs + ox;
                if (OUT64) reinterpret_cast<double *>(p.out)[at] = res;
                else       reinterpret_cast<float *>(p.out)[at] = (float)res;
            }
        }
    }
}

// Which of the two exact treatments an output with almost nothing valid gets must not depend on how the image is cut into
// blocks or shards (the two round differently in the last bit): it is decided per output, from its own window -- is one of
// the two outermost window rows at least half populated?  Then the weight sits on well-populated rows (a horizontal edge
// of a blank region) and the float64 re-fold is exact; otherwise (vertical edges, corners) the fix-up kernel redoes it.
// (The column warps pace the pipeline: along a vertical edge two lanes of every block come here, and running the whole
// re-fold for them cost 8 ms per shard; these two shared-memory reads do not.)
template <int H>
__device__ __forceinline__ bool pipe_outer_row_populated(const SpatialParams &p, const uint32_t *rdf, int b, int wrow0, int j0,
                                                         int m0, int m1, int m2, int ccol) {
    bool full_row = false;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int rel = j0 + (e ? 2 * H : 0) + wrow0, tb = rel >> 4;
        const int mt = tb == 0 ? m0 : tb == 1 ? m1 : m2;
        if (mt == PM_BLANK) continue;
        const uint32_t d = mt == PM_CLEAN ? 0u : rdf[(((b - 2 + tb) % PP_NB) * PP_R + (rel & 15)) * SP_TX + ccol];
        full_row |= d < (p.qsx >> 1);
    }
    return full_row;
}

// An output whose integer `present` is below fix_thresh: the kernel-row factors are folded again in float64 over the row
// presences (qsx - rowdef) / qsx.  That is exact whenever the rows that carry the weight are themselves well populated -- the
// horizontal edge of a blank region: full rows, reached only by the outermost kernel rows -- and the output is finished
// here.  Where the weight sits on rows with almost nothing valid in them (vertical edges, corners) the quantised kernel-column
// factors are the problem: *ok = false, and the output goes to sep_fixup_kernel.
template <int H>
__device__ __noinline__ double pipe_present_f64(const SpatialParams &p, const uint32_t *rdf, int b, int wrow0, int j0,
                                                int m0, int m1, int m2, int ccol, bool *ok) {
    *ok = false;
    double P = 0.0, E = 0.0;
    const double inv_qsx = 1.0 / (double)p.qsx;
    const uint32_t tiny = p.qsx >> 11;
    for (int k = 0; k <= 2 * H; ++k) {
        const int rel = j0 + 2 * H - k + wrow0;                            // row counted from the first row of block b - 2
        const int tb = rel >> 4;
        const int mt = tb == 0 ? m0 : tb == 1 ? m1 : m2;
        if (mt == PM_BLANK) continue;
        uint32_t d = 0u;
        if (mt != PM_CLEAN) d = rdf[(((b - 2 + tb) % PP_NB) * PP_R + (rel & 15)) * SP_TX + ccol];
        const uint32_t pres = p.qsx - d;
        const double w = p.ty[k] * ((double)pres * inv_qsx);
        P += w;
        if (pres < tiny) E += w;
    }
    *ok = P > 0.0 && E * 1024.0 <= P;
    return P;
}

// J = outputs per thread in both passes: 16 -> 4 row warps + 4 column warps (fewest shared-memory loads per output),
// 8 -> 8 + 8 (two warps per role and scheduler: more latency hiding, 1.6x the shared-memory traffic of the passes)
template <int H, int OUT64, int J>
__global__ void __launch_bounds__(2 * (PP_R * SP_TX / J) + 32, 1)
sep_pipe_kernel(const __grid_constant__ SpatialParams p) {
    if (p.sel && sel_wants_march(p.sel)) return;         // too many missing samples: the convolved denominator is cheaper
    constexpr int NT = 2 * H + 1;
    constexpr int NIN = J + 2 * H;                       // inputs of J outputs
    constexpr int GROUP = PP_R * SP_TX / J;              // threads per role
    constexpr int NWARP = GROUP / 32;                    // warps per role
    constexpr int RW = PP_R / NWARP;                     // rows a row warp widens
    constexpr int NTHREADS = 2 * GROUP + 32;
    static_assert(H <= SP_HP && H <= PP_R && (H % 2) == 0, "half-width");
    static_assert(J == 8 || J == 16, "outputs per thread");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PipeSmem &sm = *reinterpret_cast<PipeSmem *>(smem_raw);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    int64_t bid = blockIdx.x;
    const int strip = (int)(bid % p.strips_per_row); bid /= p.strips_per_row;
    const int chunk = (int)(bid % p.chunks);
    const int64_t c = bid / p.chunks;
    const int64_t x0 = (int64_t)strip * SP_TX;
    const int64_t ya = (int64_t)chunk * p.rows_per_cta;
    const int64_t yb = min(p.ny, ya + p.rows_per_cta);
    const int nout_blk = (int)((yb - ya + PP_R - 1) / PP_R);
    const int nblk = nout_blk + 2;                       // one run-in and one run-out block
    const int64_t y_first = ya - PP_R;

    const int64_t xl = max((int64_t)0, x0 - SP_HP), xr = min(p.nx, x0 + SP_TX + SP_HP);
    const int col_off = (int)(xl - (x0 - SP_HP));
    const uint32_t row_bytes = (uint32_t)(xr - xl) * 4u;

    if (tid == 0) {
        for (int s = 0; s < PP_RS; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], NWARP); }
        for (int q = 0; q < PP_NB; ++q) { mbar_init(&sm.rfull[q], NWARP); mbar_init(&sm.rempty[q], NWARP); }
        mbar_fence_init();
        sm.count[0] = sm.count[1] = sm.count[2] = 0;
    }
    if (tid < 96) { const int k = tid - 32; sm.qyp[tid] = (k >= 0 && k < NT) ? p.qy[k] : 0u; }
    if (tid >= 96 && tid < 256) { const int t = tid - 96, k = t - 64; sm.cqxp[t] = p.cqx[k < 0 ? 0 : (k > NT ? NT : k)]; }

    // the columns of the raw rows that lie outside the image are never written by the copies: valid zeros, set once
    if (col_off > 0 || (int)(xr - xl) < SP_W) {
        for (int i = tid; i < PP_RS * PP_R * SP_W; i += NTHREADS) {
            const int col = i % SP_W;
            if (col < col_off || col >= col_off + (int)(xr - xl)) (&sm.raw[0][0][0])[i] = 0.0f;
        }
    }
    __syncthreads();

    if (warp == 2 * NWARP) {
        // ---------------- producer warp ----------------
        const uint64_t pol = l2_evict_first_policy();
        for (int b = 0; b < nblk; ++b) {
            const int s = b % PP_RS;
            if (b >= PP_RS) mbar_wait(&sm.empty[s], ((b / PP_RS) - 1) & 1);
            const float *src = nullptr;
            if (lane < PP_R) {
                const int64_t y = y_first + (int64_t)b * PP_R + lane;
                if (y >= 0 && y < p.ny) src = p.in + c * p.stride_c + y * p.stride_y + xl;
                else if (y < 0 && p.halo_top && y >= -p.halo_rows)
                    src = p.halo_top + (c * p.halo_rows + (p.halo_rows + y)) * p.nx + xl;
                else if (y >= p.ny && p.halo_bot && y < p.ny + p.halo_rows)
                    src = p.halo_bot + (c * p.halo_rows + (y - p.ny)) * p.nx + xl;
            }
            const unsigned have = __ballot_sync(0xffffffffu, src != nullptr);
            if (have != 0xFFFFu) {
                // rows above / below the image (and no neighbouring shard): zero padding, written here with plain stores
                for (int r = 0; r < PP_R; ++r) {
                    if ((have >> r) & 1u) continue;
                    for (int q = lane; q < SP_W / 4; q += 32) *reinterpret_cast<float4 *>(&sm.raw[s][r][4 * q]) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                __syncwarp();
            }
            if (lane == 0) mbar_expect_tx(&sm.full[s], (uint32_t)__popc(have) * row_bytes);      // (release: orders the zero rows too)
            __syncwarp();
            if (src) tma_load_1d(&sm.raw[s][lane][col_off], src, row_bytes, &sm.full[s], pol);
        }
        return;
    }

    const bool strip_clipped = xl != x0 - SP_HP || xr != x0 + SP_TX + SP_HP;
    const bool mask_by_sweep = p.mask.mode == MODE_GENERIC || (p.mask.mode == MODE_INTERVAL && p.fill == p.fill);
    // zero padding is a VALID sample: the fast path's interval test may run over it only if the interval contains 0
    const bool zero_ok = p.mask.mode != MODE_INTERVAL || (p.lo_closed <= 0.0f && p.hi_closed >= 0.0f);

    if (warp < NWARP) {
        // ================= row warps =================
        const int t = tid;                               // 0 .. GROUP - 1
        const int rrow = t & 15, seg = t >> 4;           // row of the block, segment of J columns
        double T[NT];
#pragma unroll
        for (int k = 0; k < NT; ++k) T[k] = reg_copy(p.tx_scaled[k]);
        for (int b = 0; b < nblk; ++b) {
            const int s = b % PP_RS, par = b & 1, l3 = b % 3, slot = b % PP_NB;
            PipeBlock blk;
            blk.c = c; blk.yblk = y_first + (int64_t)b * PP_R; blk.x0 = x0; blk.xl = xl; blk.xr = xr;
            const bool rows_inside = blk.yblk >= (p.halo_top ? -(int64_t)p.halo_rows : 0) &&
                                     blk.yblk + PP_R <= p.ny + (p.halo_bot ? (int64_t)p.halo_rows : 0);
            const bool slow = mask_by_sweep || ((strip_clipped || !rows_inside) && !zero_ok);     // uniform per block
            float lo_c = -INFINITY, hi_c = INFINITY;
            if (!slow && p.mask.mode == MODE_INTERVAL) { lo_c = p.lo_closed; hi_c = p.hi_closed; }
            if (t == 0) sm.count[(b + 1) % 3] = 0;       // last read two blocks ago, next written after this block's barrier
            // the ring slot of this block: free once the column warps have released block b - 4.  Its row-deficit plane is
            // cleared HERE, ahead of the barrier that follows the widening, so that a sparse block can scatter right after
            // that barrier (clearing after it cost a second barrier per block)
            if (b >= PP_NB) mbar_wait(&sm.rempty[slot], ((b / PP_NB) - 1) & 1);
            for (int i = t; i < PP_R * SP_TX / 4; i += GROUP)
                *reinterpret_cast<uint4 *>(&sm.rowdef[slot * PP_R + (i >> 5)][4 * (i & 31)]) = make_uint4(0u, 0u, 0u, 0u);
            if (t < 4) sm.dirty[slot][t] = 0u;
            mbar_wait(&sm.full[s], (b / PP_RS) & 1);

            // ---- (1) widen every input once: 16 rows x 80 pairs ----
            if (slow) {
                pipe_widen_slow(p, sm, s, par, l3, RW, warp, lane, blk);
            } else {
                // bit 2i + e of `miss`: element e of this thread's pair i is missing (listed after the loop: the common
                // iteration is branch-free).  Under an interval mask the bounds are finite, so +-inf is "missing" and the
                // placement needs no special case for it.
                uint32_t miss = 0u;
                if (p.mask.mode == MODE_INTERVAL) miss = pipe_widen_fast<true, RW>(sm, s, par, warp, lane, lo_c, hi_c);
                else                              miss = pipe_widen_fast<false, RW>(sm, s, par, warp, lane, lo_c, hi_c);
                if (__ballot_sync(0xffffffffu, miss != 0u)) {           // warp-uniform: ONE atomic per warp whatever the count
                    const int n = __popc(miss);
                    int incl = n;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int up = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += up; }
                    int base = 0;
                    if (lane == 31) base = atomicAdd(&sm.count[l3], incl);
                    int idx = __shfl_sync(0xffffffffu, base, 31) + incl - n;
                    while (miss && idx < PP_CAP) {
                        const int bit = __ffs((int)miss) - 1;
                        miss &= miss - 1u;
                        int row, pc;
                        pipe_pair_of(bit >> 1, RW, warp, lane, row, pc);
                        sm.list[l3][idx++] = (uint16_t)((row << 8) | (2 * pc + (bit & 1)));
                    }
                }
            }
            pipe_bar<GROUP>(1);
            const int cnt = sm.count[l3];                // uniform
            const int mode = cnt == 0 ? PM_CLEAN : cnt == PP_FULL ? PM_BLANK : cnt <= PP_CAP ? PM_SPARSE : PM_CROWDED;
            if (mode != PM_CROWDED) {                    // (a crowded block re-reads the raw rows for its dense deficits)
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[s]);
            }
            if (t < 4 && mode == PM_CROWDED) sm.dirty[slot][t] = 0xFFFFu;
            if (t == 4) sm.mode[slot] = mode;

            // ---- (2) sparse block: scatter the listed inputs' qx[.] into the (zeroed) row-deficit plane ----
            if (mode == PM_SPARSE) {
                for (int e = warp; e < cnt; e += NWARP) {
                    const int ent = sm.list[l3][e];
                    const int r_in = ent >> 8, cin = ent & 255;
                    const int xo0 = cin - SP_HP - H;                       // out(xo) reads in(xo + H - k): xo = xo0 + k
                    for (int k = lane; k < NT; k += 32) {
                        const int xo = xo0 + k;
                        if (xo >= 0 && xo < SP_TX) atomicAdd(&sm.rowdef[slot * PP_R + r_in][xo], p.qx[k]);
                    }
                    if (lane < 2) {                                        // the one or two 32-column segments the taps reach
                        const int xe = lane == 0 ? max(xo0, 0) : min(xo0 + NT - 1, SP_TX - 1);
                        if (xo0 + NT - 1 >= 0 && xo0 < SP_TX) atomicOr(&sm.dirty[slot][xe >> 5], 1u << r_in);
                    }
                }
            }

            // ---- (3) row pass: J adjacent outputs of row `rrow` from J + 2H widened inputs ----
            double *trow = &sm.top[slot * PP_R + rrow][seg * J];
            if (mode == PM_BLANK) {
#pragma unroll
                for (int j = 0; j < J; j += 2) *reinterpret_cast<double2 *>(trow + j) = make_double2(0.0, 0.0);
            } else {
                const double2 *src = reinterpret_cast<const double2 *>(&sm.wd[par][rrow][seg * J + SP_HP - H]);
                double acc[J];
#pragma unroll
                for (int j = 0; j < J; ++j) acc[j] = 0.0;
#pragma unroll
                for (int q = 0; q < NIN / 2; ++q) {
                    const double2 v2 = src[q];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double v = e ? v2.y : v2.x;
                        const int i = 2 * q + e;
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const int k = j + 2 * H - i;                   // out(j) reads in(j + 2H - k) of the window
                            if (k >= 0 && k < NT) acc[j] = fma(T[k], v, acc[j]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < J; j += 2) *reinterpret_cast<double2 *>(trow + j) = make_double2(acc[j], acc[j + 1]);
            }
            if (mode == PM_CROWDED) {
                pipe_crowded_masks(p, sm, s, RW, warp, lane, slow, lo_c, hi_c, blk);
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[s]);
                pipe_bar<GROUP>(2);
                pipe_crowded_rowdef<J>(sm, slot, rrow, seg, H);
                pipe_bar<GROUP>(3);                      // (the masks are rewritten by the next crowded block)
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.rfull[slot]);
        }
        return;
    }

    // ================= column warps =================
    {
        const int ct = tid - GROUP;                      // 0 .. GROUP - 1
        const int ccol = ct & (SP_TX - 1);               // column of the strip
        const int half = ct >> 7;                        // which J rows of the 16-row output block (0 when J = 16)
        const int cw = ccol >> 5;                        // this warp's 32-column segment
        const int64_t x = x0 + ccol;
        const bool pass = p.passthrough && p.passthrough[c];
        double T[NT];
#pragma unroll
        for (int k = 0; k < NT; ++k) T[k] = reg_copy(p.ty[k]);
        const float inv_q = 1.0f / (float)(uint32_t)(p.qall >> 31);
        const float qall_f = __ull2float_rn(p.qall);
        const double *topf = &sm.top[0][0];
        const uint32_t *rdf = &sm.rowdef[0][0];
        const int wrow0 = PP_R - H + half * J;           // first window row, counted from the first row of block b - 2
        for (int b = 0; b < nblk; ++b) {
            mbar_wait(&sm.rfull[b % PP_NB], (b / PP_NB) & 1);
            if (b < 2) continue;
            // ring blocks of the window: tb = 0, 1, 2 <-> march blocks b - 2, b - 1, b
            const int s0 = (b - 2) % PP_NB, s1 = (b - 1) % PP_NB, s2 = b % PP_NB;
            const int m0 = sm.mode[s0], m1 = sm.mode[s1], m2 = sm.mode[s2];          // uniform
            const int64_t yout = y_first + (int64_t)(b - 1) * PP_R + half * J;
            const int nlive = (int)max((int64_t)0, min((int64_t)J, yb - yout));
            const bool all_blank = m0 == PM_BLANK && m1 == PM_BLANK && m2 == PM_BLANK;

            // ---- numerators: J vertically adjacent outputs from J + 2H ring rows (window row i <-> row yout - H + i) ----
            double acc[J];
#pragma unroll
            for (int j = 0; j < J; ++j) acc[j] = 0.0;
            if (!all_blank) {
                // element offsets, in this thread's column, of the 8-row half blocks its window touches (integers: the loads
                // stay LDS): hb[m] = half block m + half of the three blocks' six
                int hb[6];
#pragma unroll
                for (int m = 0; m < 6; ++m) {
                    const int mm = m + half;                               // half = 0 or 1
                    const int sl = mm < 2 ? s0 : mm < 4 ? s1 : s2;
                    hb[m] = (sl * PP_R + (mm & 1) * 8) * PP_TOP + ccol;
                }
#pragma unroll
                for (int i = 0; i < NIN; ++i) {
                    const int rel = PP_R - H + i;                          // compile time: row relative to the window's own half-block origin
                    const double v = topf[hb[rel >> 3] + (rel & 7) * PP_TOP];
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const int k = j + 2 * H - i;
                        if (k >= 0 && k < NT) acc[j] = fma(T[k], v, acc[j]);
                    }
                }
            }

            // ---- deficits: exact integers; only the rows that are dirty in this warp's 32 columns are folded in ----
            unsigned long long def[J];
#pragma unroll
            for (int j = 0; j < J; ++j) def[j] = 0ull;
            bool any_def = false;                                          // uniform
            unsigned long long dirty48 = 0ull;                             // bit 16 tb + r: ring row r of block tb
#pragma unroll
            for (int tb = 0; tb < 3; ++tb) {
                const int mt = tb == 0 ? m0 : tb == 1 ? m1 : m2;
                const int st = tb == 0 ? s0 : tb == 1 ? s1 : s2;
                if (mt == PM_BLANK) {
                    any_def = true;
                    // every row of the block contributes (sum of qx) x qy[k]: prefix sums of qy give the block's share
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        // rows rel = 16 tb .. 16 tb + 15 <-> window rows i = rel - wrow0 <-> k = j + 2H - i
                        const int khi = j + 2 * H - (PP_R * tb - wrow0);                   // k of the block's first row
                        const int klo = khi - (PP_R - 1);                                   // k of its last row
                        const int a = klo < 0 ? 0 : (klo > NT ? NT : klo), e = khi + 1 < 0 ? 0 : (khi + 1 > NT ? NT : khi + 1);
                        if (e > a) def[j] += (unsigned long long)(p.cqy[e] - p.cqy[a]) * p.qsx;
                    }
                } else if (mt != PM_CLEAN) {
                    dirty48 |= (unsigned long long)sm.dirty[st][cw] << (16 * tb);
                }
            }
            // window rows i = rel - wrow0, i in [0, NIN)
            unsigned long long todo = (dirty48 >> wrow0) & ((1ull << NIN) - 1ull);
            any_def |= todo != 0ull;
            if (J == 16 && __popcll(todo) > 10) {
                // crowded neighbourhood: every row of the non-clean blocks, unrolled (compile-time taps, no loop overhead);
                // rows of sparse blocks that nothing was scattered into hold zeros
#pragma unroll
                for (int tb = 0; tb < 3; ++tb) {
                    const int mt = tb == 0 ? m0 : tb == 1 ? m1 : m2;
                    const int st = tb == 0 ? s0 : tb == 1 ? s1 : s2;
                    if (mt != PM_SPARSE && mt != PM_CROWDED) continue;
                    const int ord = st * PP_R * SP_TX + ccol;
#pragma unroll
                    for (int r_in = 0; r_in < PP_R; ++r_in) {
                        const int i = PP_R * tb + r_in - (PP_R - H);       // window row (half = 0 here)
                        if (i < 0 || i >= NIN) continue;
                        const uint32_t d = rdf[ord + r_in * SP_TX];
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const int k = j + 2 * H - i;
                            if (k >= 0 && k < NT) def[j] += (unsigned long long)d * p.qy[k];
                        }
                    }
                }
                todo = 0ull;
            }
#pragma unroll 1
            while (todo) {
                const int i = __ffsll((long long)todo) - 1;                // uniform
                todo &= todo - 1ull;
                const int rel = i + wrow0;
                const int st = (b - 2 + (rel >> 4)) % PP_NB;
                const uint32_t d = rdf[(st * PP_R + (rel & 15)) * SP_TX + ccol];
                const uint32_t *q = &sm.qyp[32 + 2 * H - i];               // q[j] = qy[j + 2H - i], zero outside the kernel
#pragma unroll
                for (int j = 0; j < J; ++j) def[j] += (unsigned long long)d * q[j];
            }

            // ---- epilogue: out = top * qall / present ----
            bool flagged = false;                                          // some output of this thread goes to the fix-up kernel
            char *op = reinterpret_cast<char *>(p.out) + (OUT64 ? 8 : 4) * (c * p.out_stride_c + yout * p.out_stride_y + x);
            const int64_t ostep = (OUT64 ? 8 : 4) * p.out_stride_y;
            if (x < p.nx && nlive > 0) {
                if (!any_def && !pass) {
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        if (j < nlive) {
                            if (OUT64) *reinterpret_cast<double *>(op + j * ostep) = acc[j];
                            else       *reinterpret_cast<float *>(op + j * ostep) = (float)acc[j];
                        }
                    }
                } else if (OUT64) {
#pragma unroll 1
                    for (int j0 = 0; j0 < J; j0 += 1) {
                        // (float64 output is the rarely used form: compact, register arrays indexed through a switch-free select chain)
                        double a = 0.0; unsigned long long dj = 0ull;
#pragma unroll
                        for (int j = 0; j < J; ++j) if (j == j0) { a = acc[j]; dj = def[j]; }
                        if (j0 < nlive) {
                            const unsigned long long present = p.qall - dj;
                            const uint32_t hi = (uint32_t)(present >> 31);
                            double res;
                            bool ok = false;
                            double pf = 1.0;
                            const bool low = present != 0ull && present < p.fix_thresh && !pass;
                            if (low && pipe_outer_row_populated<H>(p, rdf, b, wrow0, j0, m0, m1, m2, ccol))
                                pf = pipe_present_f64<H>(p, rdf, b, wrow0, j0, m0, m1, m2, ccol, &ok);
                            if (low && ok) res = a / pf;
                            else if (low) { res = __longlong_as_double((long long)PP_SENTINEL64); flagged = true; }
                            else if (hi < (1u << 24) || pass) res = sparse_rare_output(p, a, present, pass, c, yout + j0, x);
                            else res = dj != 0ull ? a * (place_scaled_sp(__frcp_rn((float)hi)) * p.qscale31) : a;
                            *reinterpret_cast<double *>(op + j0 * ostep) = res;
                        }
                    }
                } else {
                    // float32 output.  Common case: x = deficit / qall from the top 32 bits (2^-31 of the kernel sum) and
                    // out = top / (1 - x) in float32 (relative error < 5e-7 while x <= 1/2).  More than half of the weight
                    // missing: out = top * qall / present with `present` converted from its 64 bits (same accuracy however
                    // small it is).  Nothing valid at all (present == 0, common inside blank regions) or a plane that is
                    // copied through: the filled input.
                    uint32_t rare = 0u;
                    float *o32 = reinterpret_cast<float *>(op);
                    if (all_blank || pass) {
                        rare = (1u << J) - 1u;                             // (every deficit equals qall there)
                    } else if (nlive == J) {
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const float xdef = (float)(uint32_t)(def[j] >> 31) * inv_q;
                            rare |= xdef > 0.5f ? 1u << j : 0u;
                            *o32 = __fdividef((float)acc[j], 1.0f - xdef);
                            o32 += p.out_stride_y;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const float xdef = (float)(uint32_t)(def[j] >> 31) * inv_q;
                            rare |= xdef > 0.5f ? 1u << j : 0u;
                            if (j < nlive) *o32 = __fdividef((float)acc[j], 1.0f - xdef);
                            o32 += p.out_stride_y;
                        }
                    }
                    rare &= (1u << nlive) - 1u;
                    if (rare == (1u << nlive) - 1u) {
                        // every output of this thread is flagged: inside a blank region (or a plane copied through) nothing at
                        // all is valid under the kernel and the result is the filled input -- J INDEPENDENT loads, not a
                        // chain of them (the loop below waits a full memory latency per output)
                        bool nothing = true;
#pragma unroll
                        for (int j = 0; j < J; ++j) nothing &= def[j] == p.qall;
                        if (nothing || pass) {
                            const float *ip = p.in + c * p.stride_c + yout * p.stride_y + x;
                            float cv[J];
#pragma unroll
                            for (int j = 0; j < J; ++j) cv[j] = j < nlive ? __ldg(ip + j * p.stride_y) : 0.0f;
                            float *o = reinterpret_cast<float *>(op);
#pragma unroll
                            for (int j = 0; j < J; ++j) {
                                if (j < nlive) {
                                    if (!mask_include_rt(p.mask, cv[j], c, yout + j, x)) cv[j] = p.fill;
                                    o[j * p.out_stride_y] = cv[j];
                                }
                            }
                            rare = 0u;
                        }
                    }
                    if (rare) {
                        // out of the common path (edges of blank regions): the values move to thread-local arrays so that the
                        // loop over the flagged outputs can index them
                        float fa[J];
                        unsigned long long dd[J];
#pragma unroll
                        for (int j = 0; j < J; ++j) { fa[j] = (float)acc[j]; dd[j] = def[j]; }
                        const float *ip = p.in + c * p.stride_c + yout * p.stride_y + x;
                        float *o = reinterpret_cast<float *>(op);
#pragma unroll 1
                        while (rare) {
                            const int j0 = __ffs((int)rare) - 1;
                            rare &= rare - 1u;
                            const unsigned long long present = p.qall - dd[j0];
                            float r;
                            if (present == 0ull || pass) {
                                r = __ldg(ip + j0 * p.stride_y);
                                if (!mask_include_rt(p.mask, r, c, yout + j0, x)) r = p.fill;
                            } else if (present < p.fix_thresh) {
                                bool ok = false;
                                double pf = 1.0;
                                if (pipe_outer_row_populated<H>(p, rdf, b, wrow0, j0, m0, m1, m2, ccol))
                                    pf = pipe_present_f64<H>(p, rdf, b, wrow0, j0, m0, m1, m2, ccol, &ok);
                                if (ok) r = (float)((double)fa[j0] / pf);
                                else { r = __uint_as_float(PP_SENTINEL32); flagged = true; }   // redone exactly by sep_fixup_kernel
                            } else {
                                r = fa[j0] * __fdividef(qall_f, __ull2float_rn(present));
                            }
                            o[j0 * p.out_stride_y] = r;
                        }
                    }
                }
            }
            if (__ballot_sync(0xffffffffu, flagged) && lane == 0) pipe_list_tile(p, c, yout, nlive, x0 + 32 * cw);
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.rempty[s0]);
        }
    }
}

// ---- direct 2-D kernel: any odd x odd taps ------------------------------------------------------------
struct DirectParams {
    SpatialParams sp;
    const double *taps;      // device, normalised, (nty, ntx) row-major
    int nty, ntx;
};

template <int OUT64>
__global__ void __launch_bounds__(256)
direct2d_kernel(const __grid_constant__ DirectParams d) {
    const SpatialParams &p = d.sp;
    extern __shared__ double staps[];
    for (int i = threadIdx.x; i < d.nty * d.ntx; i += blockDim.x) staps[i] = d.taps[i];
    __syncthreads();
    const int64_t per_chan = p.ny * p.nx;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.nchan * per_chan) return;
    const int64_t c = g / per_chan;
    const int64_t rem = g - c * per_chan;
    const int64_t y = rem / p.nx, x = rem - y * p.nx;
    const int hy = d.nty >> 1, hx = d.ntx >> 1;
    double top = 0.0, bot = 0.0;
    float centre = 0.0f;
    for (int ky = 0; ky < d.nty; ++ky) {
        const int64_t yy = y + hy - ky;
        for (int kx = 0; kx < d.ntx; ++kx) {
            const int64_t xx = x + hx - kx;
            float v = 0.0f;
            if (xx >= 0 && xx < p.nx) {
                if (yy >= 0 && yy < p.ny) {
                    v = __ldg(p.in + c * p.stride_c + yy * p.stride_y + xx);
                    if (!mask_include_rt(p.mask, v, c, yy, xx)) v = p.fill;
                } else if (yy < 0 && p.halo_top && yy >= -p.halo_rows) {
                    v = __ldg(p.halo_top + (c * p.halo_rows + (p.halo_rows + yy)) * p.nx + xx);
                } else if (yy >= p.ny && p.halo_bot && yy < p.ny + p.halo_rows) {
                    v = __ldg(p.halo_bot + (c * p.halo_rows + (yy - p.ny)) * p.nx + xx);
                }
            }
            if (ky == hy && kx == hx) centre = v;
            const double k = staps[ky * d.ntx + kx];
            if (v == v) { top = fma(k, (double)v, top); bot += k; }
        }
    }
    double res = (bot == 0.0) ? (double)centre : top / bot;
    if (p.passthrough && p.passthrough[c]) res = (double)centre;
    if (OUT64) reinterpret_cast<double *>(p.out)[c * p.out_stride_c + y * p.out_stride_y + x] = res;
    else       reinterpret_cast<float *>(p.out)[c * p.out_stride_c + y * p.out_stride_y + x] = (float)res;
}

// ---- tiled direct 2-D kernel (the default for non-negative kernels; SC_DIRECT2D=0 forces direct2d_kernel) ----
// What `convolve_to` needs for rotated elliptical beams, where direct2d_kernel pays one L1 request per tap and
// output.  CTA = 256 threads = a 32-column x 64-row output tile of one channel; lane = column, warp = 8 rows.
// The filled input box (tile + kernel margin) is staged ONCE in shared memory, widened to float64, with a float
// validity plane (missing -> value 0, validity 0; zero padding outside the image is valid, as in astropy).
// Each thread keeps 8 float64 numerators: for every kernel column it walks the kernel rows with a 16-deep
// register window of its image column (1 LDS.64 + 1 LDS.32 + 1 broadcast tap per 8 DFMA + 8 FFMA); the
// denominator is summed in float32 per kernel column and carried in float64 across columns.
constexpr int DT_TX = 32, DT_RY = 8;
constexpr size_t DT_MAX_SMEM = 200 * 1024;        // taps + box have to fit (227 KB per CTA on sm_100)

// NW warps per CTA: the tile is 32 columns x 8 NW rows (8 warps for the usual kernels, 4 when the box of a wide
// kernel would not fit otherwise)
template <int OUT64, int NW>
__global__ void __launch_bounds__(32 * NW, NW == 4 ? 1 : 2)
direct2d_tiled_kernel(const __grid_constant__ DirectParams d, int tiles_x, int tiles_y) {
    constexpr int DT_TY = NW * DT_RY;
    constexpr int NTHREADS = 32 * NW;
    const SpatialParams &p = d.sp;
    extern __shared__ __align__(16) unsigned char dt_smem[];
    const int nty = d.nty, ntx = d.ntx, hy = nty >> 1, hx = ntx >> 1;
    const int bw = DT_TX + ntx - 1;                    // box width
    const int bh = DT_TY + nty - 1 + DT_RY;            // box height (+ one window of never-used rows)
    double *staps = reinterpret_cast<double *>(dt_smem);
    double *sval = staps + ((nty * ntx + 1) & ~1);
    float *sok = reinterpret_cast<float *>(sval + (size_t)bw * bh);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t b = blockIdx.x;
    const int64_t tx_i = b % tiles_x; b /= tiles_x;
    const int64_t ty_i = b % tiles_y;
    const int64_t c = b / tiles_y;
    const int64_t x0 = tx_i * DT_TX, y0 = ty_i * DT_TY;

    for (int i = tid; i < nty * ntx; i += NTHREADS) staps[i] = d.taps[i];
    for (int i = tid; i < bw * bh; i += NTHREADS) {
        const int q = i / bw, pcol = i - q * bw;
        const int64_t yy = y0 - hy + q, xx = x0 - hx + pcol;
        float v = 0.0f;                                // zero padding: a valid sample
        if (q < DT_TY + nty - 1 && xx >= 0 && xx < p.nx) {
            if (yy >= 0 && yy < p.ny) {
                v = __ldg(p.in + c * p.stride_c + yy * p.stride_y + xx);
                if (!mask_include_rt(p.mask, v, c, yy, xx)) v = p.fill;
            } else if (yy < 0 && p.halo_top && yy >= -p.halo_rows) {
                v = __ldg(p.halo_top + (c * p.halo_rows + (p.halo_rows + yy)) * p.nx + xx);
            } else if (yy >= p.ny && p.halo_bot && yy < p.ny + p.halo_rows) {
                v = __ldg(p.halo_bot + (c * p.halo_rows + (yy - p.ny)) * p.nx + xx);
            }
        }
        const bool good = v == v;
        sval[i] = good ? (double)v : 0.0;
        sok[i] = good ? 1.0f : 0.0f;
    }
    __syncthreads();

    const int r0 = warp * DT_RY;                       // first of this thread's 8 output rows (tile-local)
    double top[DT_RY], botd[DT_RY];
#pragma unroll
    for (int j = 0; j < DT_RY; ++j) { top[j] = 0.0; botd[j] = 0.0; }
    // out(r, x) = sum_{s, sx} K[nty-1-s][ntx-1-sx] * box[r + s][x + sx]
    for (int sx = 0; sx < ntx; ++sx) {
        const int kx = ntx - 1 - sx;
        const double *colv = sval + (size_t)r0 * bw + lane + sx;
        const float *colo = sok + (size_t)r0 * bw + lane + sx;
        double a[2 * DT_RY];
        float o[2 * DT_RY], botf[DT_RY];
#pragma unroll
        for (int j = 0; j < DT_RY; ++j) { a[j] = colv[(size_t)j * bw]; o[j] = colo[(size_t)j * bw]; botf[j] = 0.0f; }
        for (int sb = 0; sb < nty; sb += DT_RY) {
#pragma unroll
            for (int j = 0; j < DT_RY; ++j) {
                a[DT_RY + j] = colv[(size_t)(sb + DT_RY + j) * bw];
                o[DT_RY + j] = colo[(size_t)(sb + DT_RY + j) * bw];
            }
#pragma unroll
            for (int si = 0; si < DT_RY; ++si) {
                if (sb + si < nty) {                   // uniform across the CTA
                    const double k = staps[(nty - 1 - sb - si) * ntx + kx];
                    const float kf = (float)k;
#pragma unroll
                    for (int j = 0; j < DT_RY; ++j) {
                        top[j] = fma(k, a[j + si], top[j]);
                        botf[j] = fmaf(kf, o[j + si], botf[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < DT_RY; ++j) { a[j] = a[DT_RY + j]; o[j] = o[DT_RY + j]; }
        }
#pragma unroll
        for (int j = 0; j < DT_RY; ++j) botd[j] += (double)botf[j];
    }

    const int64_t x = x0 + lane;
    const bool pass = p.passthrough && p.passthrough[c];
#pragma unroll
    for (int j = 0; j < DT_RY; ++j) {
        const int64_t y = y0 + r0 + j;
        if (x < p.nx && y < p.ny) {
            const size_t ci = (size_t)(r0 + j + hy) * bw + lane + hx;
            const double centre = sok[ci] != 0.0f ? sval[ci] : nan64();
            double res = (botd[j] == 0.0) ? centre : top[j] / botd[j];
            if (pass) res = centre;
            if (OUT64) reinterpret_cast<double *>(p.out)[c * p.out_stride_c + y * p.out_stride_y + x] = res;
            else       reinterpret_cast<float *>(p.out)[c * p.out_stride_c + y * p.out_stride_y + x] = (float)res;
        }
    }
}

// flags[c] = 1 when no voxel of channel c is included by the mask: `_apply_spatial_function`
// (spectral_cube.py:161-172) copies such a plane through instead of convolving it.  One CTA per
// channel, early exit as soon as any thread sees an included voxel (the common case: first pass).
__global__ void __launch_bounds__(256)
plane_none_included_kernel(const __grid_constant__ SpatialParams p, uint8_t *flags) {
    __shared__ int found;
    const int64_t c = blockIdx.x;
    if (threadIdx.x == 0) found = 0;
    __syncthreads();
    const int64_t n = p.ny * p.nx;
    for (int64_t base = 0; base < n; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        bool inc = false;
        if (i < n) {
            const int64_t y = i / p.nx, x = i - y * p.nx;
            const float v = __ldg(p.in + c * p.stride_c + y * p.stride_y + x);
            inc = mask_include_rt(p.mask, v, c, y, x);
        }
        if (__syncthreads_or(inc)) { if (threadIdx.x == 0) found = 1; break; }
    }
    __syncthreads();
    if (threadIdx.x == 0) flags[c] = found ? 0 : 1;
}

// the tiled kernel when taps + box fit shared memory and no tap is negative (10x faster on a B200: 183 vs 1859 ms for 29x29 taps
// on 4096x4096x64, profiles/r02_convolve_to_n1.jsonl), else direct2d_kernel (the tiled
// kernel's "nothing valid" test is a float32 sum of non-negative terms)
static int launch_direct2d(const DirectParams &d, int out_dtype, bool nonneg, cudaStream_t s) {
    const SpatialParams &p = d.sp;
    const int nt = d.nty * d.ntx;
    const size_t taps_bytes = (size_t)((nt + 1) & ~1) * 8;
    const size_t smem8 = taps_bytes + (size_t)(DT_TX + d.ntx - 1) * (8 * DT_RY + d.nty - 1 + DT_RY) * 12;
    const size_t smem4 = taps_bytes + (size_t)(DT_TX + d.ntx - 1) * (4 * DT_RY + d.nty - 1 + DT_RY) * 12;
    if (env_int("SC_DIRECT2D", 1) == 1 && nonneg && smem4 <= DT_MAX_SMEM) {
        const int nw = smem8 <= DT_MAX_SMEM ? 8 : 4;
        const size_t smem = nw == 8 ? smem8 : smem4;
        const int64_t tiles_x = cdiv(p.nx, DT_TX), tiles_y = cdiv(p.ny, nw * DT_RY);
        const int64_t grid = tiles_x * tiles_y * p.nchan;
        SC_CHECK_ARG(grid < ((int64_t)1 << 31), "grid too large");
#define SC_LAUNCH_TILED(O64, NWARPS) do { \
        SC_CUDA(cudaFuncSetAttribute(direct2d_tiled_kernel<O64, NWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        direct2d_tiled_kernel<O64, NWARPS><<<(unsigned)grid, 32 * NWARPS, smem, s>>>(d, (int)tiles_x, (int)tiles_y); } while (0)
        if (out_dtype == SC_F64) { if (nw == 8) SC_LAUNCH_TILED(1, 8); else SC_LAUNCH_TILED(1, 4); }
        else                     { if (nw == 8) SC_LAUNCH_TILED(0, 8); else SC_LAUNCH_TILED(0, 4); }
#undef SC_LAUNCH_TILED
    } else {
        const int64_t total = p.nchan * p.ny * p.nx;
        if (out_dtype == SC_F64) direct2d_kernel<1><<<(unsigned)cdiv(total, 256), 256, (size_t)nt * 8, s>>>(d);
        else                     direct2d_kernel<0><<<(unsigned)cdiv(total, 256), 256, (size_t)nt * 8, s>>>(d);
    }
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}

static int maybe_passthrough_flags(SpatialParams &p, int plane_passthrough, void *workspace, size_t workspace_bytes,
                                   size_t offset, cudaStream_t s) {
    p.passthrough = nullptr;
    if (!plane_passthrough || p.mask.mode == MODE_NONE) return SC_OK;
    const size_t need = offset + (size_t)p.nchan + 256;
    if (!workspace || workspace_bytes < need) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return SC_ERR_WORKSPACE;
    }
    uint8_t *flags = (uint8_t *)workspace + offset;
    LaunchScope ls(0, s);
    plane_none_included_kernel<<<(unsigned)p.nchan, 256, 0, s>>>(p, flags);
    SC_CUDA(cudaGetLastError());
    p.passthrough = flags;
    return SC_OK;
}

// One warp per sampled block (8 rows x 128 columns, the pipe kernel's unit of work) of up to 16 planes: sel[0] counts the
// CROWDED blocks -- more than 1/64 of the samples missing, but not all of them -- and sel[1] the blocks seen.
__global__ void __launch_bounds__(256)
missing_sample_kernel(const __grid_constant__ SpatialParams p, unsigned int *sel) {
    const int64_t cstep = max((int64_t)1, p.nchan / 16);
    const int64_t ncs = (p.nchan + cstep - 1) / cstep, bands = (p.ny + 7) / 8, strips = (p.nx + 127) / 128;
    const int64_t total = ncs * bands * strips;
    const int lane = threadIdx.x & 31;
    unsigned int crowded = 0, seen = 0;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < total; i += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int64_t x0 = (i % strips) * 128, r = i / strips;
        const int64_t y0 = (r % bands) * 8, c = (r / bands) * cstep;
        unsigned int miss = 0, n = 0;
        float v[32];                                                      // all 32 loads in flight before the mask is evaluated
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int64_t y = y0 + (j >> 2), x = x0 + 32 * (j & 3) + lane;
            v[j] = (y < p.ny && x < p.nx) ? __ldg(p.in + c * p.stride_c + y * p.stride_y + x) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int64_t y = y0 + (j >> 2), x = x0 + 32 * (j & 3) + lane;
            if (y < p.ny && x < p.nx) {
                miss += (v[j] != v[j] || !mask_include_rt(p.mask, v[j], c, y, x)) ? 1u : 0u;
                n += 1u;
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) { miss += __shfl_xor_sync(0xffffffffu, miss, d); n += __shfl_xor_sync(0xffffffffu, n, d); }
        crowded += (miss * 64u > n && miss < n) ? 1u : 0u;
        seen += 1u;
    }
    if (lane == 0 && seen) { atomicAdd(&sel[0], crowded); atomicAdd(&sel[1], seen); }
}

template <int H, int OUT64>
static cudaError_t launch_sep_one(const SpatialParams &p, unsigned grid, cudaStream_t s) {
    constexpr int HB = (H + SP_R - 1) / SP_R;
    constexpr int NB = 2 * HB + 2;
    auto kern = sep_march_kernel<H, OUT64>;
    const size_t smem = sizeof(SpatialSmem<NB>);
    static unsigned long long configured = 0;        // per instantiation, one bit per device
    if (cudaError_t e = ensure_dyn_smem(kern, smem, &configured)) return e;
    kern<<<grid, SP_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

template <int H, int OUT64>
static cudaError_t launch_sparse_one(const SpatialParams &p, unsigned grid, cudaStream_t s) {
    constexpr int HB = (H + SP_R - 1) / SP_R;
    constexpr int NB = 2 * HB + 2;
    auto kern = sep_sparse_kernel<H, OUT64>;
    const size_t smem = sizeof(SparseSmem<NB>);
    static unsigned long long configured = 0;        // per instantiation, one bit per device
    if (cudaError_t e = ensure_dyn_smem(kern, smem, &configured)) return e;
    kern<<<grid, SQ_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

template <int H, int OUT64, int J>
static cudaError_t launch_pipe_j(SpatialParams p, unsigned grid, cudaStream_t s) {
    auto kern = sep_pipe_kernel<H, OUT64, J>;
    const size_t smem = sizeof(PipeSmem);
    static unsigned long long configured = 0;        // per instantiation, one bit per device
    if (cudaError_t e = ensure_dyn_smem(kern, smem, &configured)) return e;
    p.hpad = H;
    p.fix_bands = (p.ny + FX_MAXROWS - 1) / FX_MAXROWS;
    p.fix_segs = (p.nx + FX_TX - 1) / FX_TX;
    const size_t tiles = (size_t)(p.nchan * p.fix_bands * p.fix_segs);
    if (cudaError_t e = cudaMemsetAsync(p.fix_state, 0, 8, s)) return e;
    if (cudaError_t e = cudaMemsetAsync(p.fix_bitmap, 0, std::min(PP_FIX_BITMAP_BYTES, (tiles / 32 + 1) * 4), s)) return e;
    kern<<<grid, 2 * (PP_R * SP_TX / J) + 32, smem, s>>>(p);
    if (cudaError_t e = cudaGetLastError()) return e;
    sep_fixup_kernel<OUT64><<<148 * 12, FX_THREADS, 0, s>>>(p);            // returns at once when nothing was listed
    return cudaGetLastError();
}

template <int H, int OUT64>
static cudaError_t launch_pipe_one(const SpatialParams &p, unsigned grid, cudaStream_t s) {
    if (p.fix_state == nullptr) return cudaErrorInvalidValue;             // (the entry point sized the workspace for it)
    // outputs per thread: 16 (4 + 4 compute warps) or 8 (8 + 8); SC_SPATIAL_J overrides the default for experiments
    if (env_int("SC_SPATIAL_J", PP_DEFAULT_J) == 8) return launch_pipe_j<H, OUT64, 8>(p, grid, s);
    return launch_pipe_j<H, OUT64, 16>(p, grid, s);
}

template <int OUT64>
static cudaError_t launch_pipe_h(const SpatialParams &p, int h, unsigned grid, cudaStream_t s) {
    if (h <= 2)  return launch_pipe_one<2, OUT64>(p, grid, s);
    if (h <= 4)  return launch_pipe_one<4, OUT64>(p, grid, s);
    if (h <= 6)  return launch_pipe_one<6, OUT64>(p, grid, s);
    if (h <= 8)  return launch_pipe_one<8, OUT64>(p, grid, s);
    if (h <= 10) return launch_pipe_one<10, OUT64>(p, grid, s);
    if (h <= 12) return launch_pipe_one<12, OUT64>(p, grid, s);
    if (h <= 14) return launch_pipe_one<14, OUT64>(p, grid, s);
    return launch_pipe_one<16, OUT64>(p, grid, s);
}

template <int OUT64>
static cudaError_t launch_sparse_h(const SpatialParams &p, int h, unsigned grid, cudaStream_t s) {
    if (h <= 2)  return launch_sparse_one<2, OUT64>(p, grid, s);
    if (h <= 4)  return launch_sparse_one<4, OUT64>(p, grid, s);
    if (h <= 6)  return launch_sparse_one<6, OUT64>(p, grid, s);
    if (h <= 8)  return launch_sparse_one<8, OUT64>(p, grid, s);
    if (h <= 10) return launch_sparse_one<10, OUT64>(p, grid, s);
    if (h <= 12) return launch_sparse_one<12, OUT64>(p, grid, s);
    if (h <= 14) return launch_sparse_one<14, OUT64>(p, grid, s);
    return launch_sparse_one<16, OUT64>(p, grid, s);
}

template <int OUT64>
static cudaError_t launch_sep_h(const SpatialParams &p, int h, unsigned grid, cudaStream_t s) {
    if (h <= 2)  return launch_sep_one<2, OUT64>(p, grid, s);
    if (h <= 4)  return launch_sep_one<4, OUT64>(p, grid, s);
    if (h <= 6)  return launch_sep_one<6, OUT64>(p, grid, s);
    if (h <= 8)  return launch_sep_one<8, OUT64>(p, grid, s);
    if (h <= 10) return launch_sep_one<10, OUT64>(p, grid, s);
    if (h <= 12) return launch_sep_one<12, OUT64>(p, grid, s);
    if (h <= 14) return launch_sep_one<14, OUT64>(p, grid, s);
    return launch_sep_one<16, OUT64>(p, grid, s);
}

static int fill_common(SpatialParams &p, const float *in, void *out, int out_dtype,
                       int64_t nchan, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y,
                       int64_t out_stride_c, int64_t out_stride_y, const sc_mask_desc *mask, double fill,
                       const float *halo_top, const float *halo_bot, int halo_rows, int plane_passthrough) {
    int rc = check_cube_args(in, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(out != nullptr && (const void *)in != out, "out is NULL or aliases the input (spatial smoothing is out of place)");
    SC_CHECK_ARG(out_dtype == SC_F32 || out_dtype == SC_F64, "out_dtype must be SC_F32 or SC_F64");
    SC_CHECK_ARG(out_stride_y >= nx && out_stride_c >= out_stride_y, "bad output strides");
    SC_CHECK_ARG(halo_rows >= 0, "halo_rows must be >= 0");
    SC_CHECK_ARG(halo_rows > 0 || (!halo_top && !halo_bot), "halo buffers given but halo_rows == 0");
    p.in = in; p.out = out; p.nchan = nchan; p.ny = ny; p.nx = nx;
    p.stride_c = stride_c; p.stride_y = stride_y; p.out_stride_c = out_stride_c; p.out_stride_y = out_stride_y;
    p.halo_top = halo_top; p.halo_bot = halo_bot; p.halo_rows = halo_rows;
    p.fill = (float)fill;
    p.passthrough = nullptr;
    (void)plane_passthrough;
    return build_dev_mask(mask, in, stride_c, stride_y, &p.mask);
}

}  // namespace scb

using namespace scb;

extern "C" int sc_spatial_missing_sample(const float *in, int64_t nchan, int64_t ny, int64_t nx,
                                         int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask,
                                         unsigned int *counts, void *stream) {
    int rc = check_cube_args(in, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(counts != nullptr, "counts is NULL");
    SpatialParams p{};
    p.in = in; p.nchan = nchan; p.ny = ny; p.nx = nx; p.stride_c = stride_c; p.stride_y = stride_y;
    rc = build_dev_mask(mask, in, stride_c, stride_y, &p.mask);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    LaunchScope ls(0, s);
    missing_sample_kernel<<<148 * 2, 256, 0, s>>>(p, counts);
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}

extern "C" int sc_spatial_smooth_sep(const float *in, void *out, int out_dtype,
                                     int64_t nchan, int64_t ny, int64_t nx,
                                     int64_t stride_c, int64_t stride_y,
                                     int64_t out_stride_c, int64_t out_stride_y,
                                     const sc_mask_desc *mask, double fill,
                                     const double *taps_y, int ntaps_y, const double *taps_x, int ntaps_x,
                                     const float *halo_top, const float *halo_bot, int halo_rows,
                                     int plane_passthrough,
                                     void *workspace, size_t workspace_bytes, void *stream) {
    return sc_spatial_smooth_sep_ex(in, out, out_dtype, nchan, ny, nx, stride_c, stride_y, out_stride_c, out_stride_y,
                                    mask, fill, taps_y, ntaps_y, taps_x, ntaps_x, halo_top, halo_bot, halo_rows,
                                    plane_passthrough, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int sc_spatial_smooth_sep_ex(const float *in, void *out, int out_dtype,
                                        int64_t nchan, int64_t ny, int64_t nx,
                                        int64_t stride_c, int64_t stride_y,
                                        int64_t out_stride_c, int64_t out_stride_y,
                                        const sc_mask_desc *mask, double fill,
                                        const double *taps_y, int ntaps_y, const double *taps_x, int ntaps_x,
                                        const float *halo_top, const float *halo_bot, int halo_rows,
                                        int plane_passthrough, const unsigned int *strategy_counts,
                                        void *workspace, size_t workspace_bytes, void *stream) {
    SpatialParams p{};
    int rc = fill_common(p, in, out, out_dtype, nchan, ny, nx, stride_c, stride_y, out_stride_c, out_stride_y,
                         mask, fill, halo_top, halo_bot, halo_rows, plane_passthrough);
    if (rc) return rc;
    SC_CHECK_ARG(taps_y && taps_x, "taps are NULL");
    SC_CHECK_ARG(ntaps_y >= 1 && (ntaps_y & 1) && ntaps_x >= 1 && (ntaps_x & 1), "Kernel size must be odd in all axes.");
    const int hy = ntaps_y >> 1, hx = ntaps_x >> 1;
    const int h = hy > hx ? hy : hx;
    SC_CHECK_ARG(halo_rows == 0 || halo_rows >= hy, "halo_rows=%d is smaller than the kernel half-height %d", halo_rows, hy);
    double sy = 0.0, sx = 0.0;
    for (int k = 0; k < ntaps_y; ++k) sy += taps_y[k];
    for (int k = 0; k < ntaps_x; ++k) sx += taps_x[k];
    SC_CHECK_ARG(fabs(sy * sx) > 1e-8, "The kernel can't be normalized, because its sum is close to zero.");
    cudaStream_t s = (cudaStream_t)stream;
    rc = maybe_passthrough_flags(p, plane_passthrough, workspace, workspace_bytes,
                                 (size_t)ntaps_y * ntaps_x * 8 + 512, s);
    if (rc) return rc;
    const bool aligned = ((uintptr_t)in % 16 == 0) && stride_c % 4 == 0 && stride_y % 4 == 0 && nx % 4 == 0 &&
                         (!halo_top || (uintptr_t)halo_top % 16 == 0) && (!halo_bot || (uintptr_t)halo_bot % 16 == 0);
    // 0 auto (pipe kernel or, for heavily masked cubes, march), 1 direct, 3 march, 4 the round-1 sparse kernel, 5 pipe
    const int choice = env_int("SC_SPATIAL_KERNEL", 0);
    if (h <= 16 && aligned && choice != 1) {
        const int H = h <= 2 ? 2 : h <= 4 ? 4 : h <= 6 ? 6 : h <= 8 ? 8 : h <= 10 ? 10 : h <= 12 ? 12 : h <= 14 ? 14 : 16;
        double ksy = 0.0, ksx = 0.0;
        for (int k = 0; k < SP_MAX_TAPS; ++k) { p.ty[k] = p.tx[k] = 0.0; p.tyf[k] = p.txf[k] = 0.0f; }
        for (int k = 0; k < ntaps_y; ++k) { p.ty[k + H - hy] = taps_y[k] / sy; ksy += p.ty[k + H - hy]; p.tyf[k + H - hy] = (float)p.ty[k + H - hy]; }
        for (int k = 0; k < SP_MAX_TAPS; ++k) p.tx_scaled[k] = 0.0;
        for (int k = 0; k < ntaps_x; ++k) { p.tx[k + H - hx] = taps_x[k] / sx; ksx += p.tx[k + H - hx]; p.txf[k + H - hx] = (float)p.tx[k + H - hx]; p.tx_scaled[k + H - hx] = ldexp(p.tx[k + H - hx], 896); }
        p.lo_closed = nextafterf(p.mask.lo, INFINITY); p.hi_closed = nextafterf(p.mask.hi, -INFINITY);
        p.ksum = ksy * ksx;
        p.strips_per_row = (int)cdiv(nx, SP_TX);
        // row chunks: enough CTAs to fill the chip, each long enough to amortise the 2H-row run-in
        int64_t chunks = 1;
        const int64_t base = nchan * p.strips_per_row;
        while (base * chunks < 148 * 4 && ny / (chunks * 2) >= 8 * h + 16) chunks *= 2;
        const int forced = env_int("SC_SPATIAL_CHUNKS", 0);
        if (forced > 0) chunks = forced;
        p.rows_per_cta = (int)(cdiv(cdiv(ny, chunks), PP_R) * PP_R);     // (a multiple of both kernels' block heights)
        p.chunks = (int)cdiv(ny, p.rows_per_cta);
        const int64_t grid = base * p.chunks;
        SC_CHECK_ARG(grid < ((int64_t)1 << 31), "grid too large");
        // integer factors of the denominator (sparse kernel): needs factors in [0, 1]
        bool nonneg = true;
        unsigned long long qsy = 0, qsx = 0;
        for (int k = 0; k < SP_MAX_TAPS; ++k) {
            nonneg = nonneg && p.ty[k] >= 0.0 && p.tx[k] >= 0.0 && p.ty[k] <= 1.0 && p.tx[k] <= 1.0;
            p.qy[k] = (uint32_t)llrint(fmin(fmax(p.ty[k], 0.0), 1.0) * 2147483648.0);
            p.qx[k] = (uint32_t)llrint(fmin(fmax(p.tx[k], 0.0), 1.0) * 2147483648.0);
            qsy += p.qy[k]; qsx += p.qx[k];
        }
        nonneg = nonneg && qsy < (1ull << 32) && qsx < (1ull << 32) && qsy > 0 && qsx > 0;
        p.cqy[0] = p.cqx[0] = 0u;
        for (int k = 0; k < SP_MAX_TAPS; ++k) { p.cqy[k + 1] = p.cqy[k] + p.qy[k]; p.cqx[k + 1] = p.cqx[k] + p.qx[k]; }   // (wrap only when !nonneg: unused then)
        p.qall = qsy * qsx;
        p.qsx = (uint32_t)qsx;
        p.qscale31 = ldexp((double)p.qall, 896 - 31);
        // which denominator strategy: 2 = sparse, 3 = convolved, 0 = let a sample of the cube decide on the device
        p.sel = nullptr;
        const size_t sel_off = (size_t)ntaps_y * ntaps_x * 8 + 256;
        const bool can_sample = workspace && workspace_bytes >= sel_off + 8;
        if (nonneg && choice == 0 && strategy_counts) {
            p.sel = strategy_counts;                                 // the caller's (job-wide) sample
        } else if (nonneg && choice == 0 && can_sample) {
            unsigned int *sel = (unsigned int *)((uint8_t *)workspace + sel_off);
            SC_CUDA(cudaMemsetAsync(sel, 0, 8, s));
            LaunchScope ls0(0, s);
            missing_sample_kernel<<<148 * 2, 256, 0, s>>>(p, sel);
            SC_CUDA(cudaGetLastError());
            p.sel = sel;
        }
        // the pipe kernel's fix-up list lives behind the taps, the strategy counters and the pass-through flags
        const size_t fix_off = (((size_t)ntaps_y * ntaps_x * 8 + 512 + (size_t)nchan + 256) + 255) & ~(size_t)255;
        p.fix_state = nullptr; p.fix_list = nullptr; p.fix_bitmap = nullptr; p.fix_cap = 0;
        p.fix_thresh = p.qall >> 13;
        if (nonneg && choice != 3 && choice != 4) {
            const size_t need = fix_off + 256 + (size_t)PP_FIX_CAP * 16 + PP_FIX_BITMAP_BYTES;
            if (!workspace || workspace_bytes < need) {
                set_error("workspace too small: need %zu bytes, got %zu (sc_workspace_bytes(SC_OP_SPATIAL_SMOOTH, ...))", need, workspace_bytes);
                return SC_ERR_WORKSPACE;
            }
            p.fix_state = (unsigned int *)((uint8_t *)workspace + fix_off);
            p.fix_list = (uint32_t *)((uint8_t *)workspace + fix_off + 256);
            p.fix_bitmap = (uint32_t *)((uint8_t *)workspace + fix_off + 256 + (size_t)PP_FIX_CAP * 16);
            // (SC_SPATIAL_FIX_CAP shrinks the list so that a test can reach the list-full path: the fix-up then scans every tile)
            p.fix_cap = (uint32_t)std::min<long long>(PP_FIX_CAP, std::max(1, env_int("SC_SPATIAL_FIX_CAP", (int)PP_FIX_CAP)));
        }
        LaunchScope ls(SC_OP_SPATIAL_SMOOTH, s);
        cudaError_t e = cudaSuccess;
        if (nonneg && choice == 4)
            e = out_dtype == SC_F64 ? launch_sparse_h<1>(p, H, (unsigned)grid, s) : launch_sparse_h<0>(p, H, (unsigned)grid, s);
        else if (nonneg && choice != 3)
            e = out_dtype == SC_F64 ? launch_pipe_h<1>(p, H, (unsigned)grid, s) : launch_pipe_h<0>(p, H, (unsigned)grid, s);
        if (e == cudaSuccess && (!nonneg || choice == 3 || p.sel))
            e = out_dtype == SC_F64 ? launch_sep_h<1>(p, H, (unsigned)grid, s) : launch_sep_h<0>(p, H, (unsigned)grid, s);
        if (e != cudaSuccess) return cuda_fail(e, "separable spatial kernel launch");
        return SC_OK;
    }
    // fall back to the direct kernel on the outer product
    const int nt = ntaps_y * ntaps_x;
    const size_t need = (size_t)nt * 8 + 256;
    if (!workspace || workspace_bytes < need) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return SC_ERR_WORKSPACE;
    }
    SC_CHECK_ARG(nt <= 6000, "kernel too large for the direct path (%d taps)", nt);
    double *host = (double *)malloc((size_t)nt * 8);
    bool all_nonneg = true;
    for (int a = 0; a < ntaps_y; ++a) for (int b = 0; b < ntaps_x; ++b) {
        host[a * ntaps_x + b] = (taps_y[a] / sy) * (taps_x[b] / sx);
        all_nonneg = all_nonneg && host[a * ntaps_x + b] >= 0.0;
    }
    double *tdev = (double *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    cudaError_t ce = cudaMemcpyAsync(tdev, host, (size_t)nt * 8, cudaMemcpyHostToDevice, s);
    free(host);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaMemcpyAsync(taps)");
    DirectParams d{p, tdev, ntaps_y, ntaps_x};
    LaunchScope ls(SC_OP_SPATIAL_SMOOTH, s);
    return launch_direct2d(d, out_dtype, all_nonneg, s);
}

extern "C" int sc_spatial_smooth_2d(const float *in, void *out, int out_dtype,
                                    int64_t nchan, int64_t ny, int64_t nx,
                                    int64_t stride_c, int64_t stride_y,
                                    int64_t out_stride_c, int64_t out_stride_y,
                                    const sc_mask_desc *mask, double fill,
                                    const double *taps, int ntaps_y, int ntaps_x,
                                    const float *halo_top, const float *halo_bot, int halo_rows,
                                    int plane_passthrough,
                                    void *workspace, size_t workspace_bytes, void *stream) {
    SpatialParams p{};
    int rc = fill_common(p, in, out, out_dtype, nchan, ny, nx, stride_c, stride_y, out_stride_c, out_stride_y,
                         mask, fill, halo_top, halo_bot, halo_rows, plane_passthrough);
    if (rc) return rc;
    SC_CHECK_ARG(taps != nullptr, "taps is NULL");
    SC_CHECK_ARG(ntaps_y >= 1 && (ntaps_y & 1) && ntaps_x >= 1 && (ntaps_x & 1), "Kernel size must be odd in all axes.");
    SC_CHECK_ARG(halo_rows == 0 || halo_rows >= (ntaps_y >> 1), "halo_rows=%d is smaller than the kernel half-height %d", halo_rows, ntaps_y >> 1);
    const int nt = ntaps_y * ntaps_x;
    SC_CHECK_ARG(nt <= 6000, "kernel too large for the direct path (%d taps)", nt);
    double sum = 0.0;
    for (int i = 0; i < nt; ++i) sum += taps[i];
    SC_CHECK_ARG(fabs(sum) > 1e-8, "The kernel can't be normalized, because its sum is close to zero.");
    const size_t need = (size_t)nt * 8 + 256;
    if (!workspace || workspace_bytes < need) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return SC_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    rc = maybe_passthrough_flags(p, plane_passthrough, workspace, workspace_bytes, (size_t)nt * 8 + 512, s);
    if (rc) return rc;
    double *host = (double *)malloc((size_t)nt * 8);
    for (int i = 0; i < nt; ++i) host[i] = taps[i] / sum;
    double *tdev = (double *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    cudaError_t ce = cudaMemcpyAsync(tdev, host, (size_t)nt * 8, cudaMemcpyHostToDevice, s);
    free(host);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaMemcpyAsync(taps)");
    DirectParams d{p, tdev, ntaps_y, ntaps_x};
    LaunchScope ls(SC_OP_SPATIAL_SMOOTH, s);
    bool nonneg = true;
    for (int i = 0; i < nt; ++i) nonneg = nonneg && taps[i] / sum >= 0.0;
    return launch_direct2d(d, out_dtype, nonneg, s);
}
