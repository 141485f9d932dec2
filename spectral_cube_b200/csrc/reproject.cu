// reproject.cu -- spatial reprojection of every channel onto a new celestial WCS.
//
// Replaces `reproject.reproject_interp((data, header), wcs_out, shape_out=..., order=...)` as
// called at spectral_cube.py:2726-2732.  Two kernels:
//
//  * wcs_pixel_map_kernel: for every output pixel, output-WCS pixel -> world and input-WCS
//    world -> pixel in float64 (FITS WCS paper I linear part, paper II TAN/SIN zenithal
//    projections and the native<->celestial rotation).  The transform is carried through unit
//    vectors, so no inverse trigonometric function is evaluated.  The spectral axis is
//    uncorrelated with the celestial axes (enforced by the reference, spectral_cube.py:1513-1515),
//    so the two coordinate planes are computed once and shared by all channels.
//
//  * reproject_kernel: a thread owns one output pixel (CTA = 32 x 8 output tile, so the input
//    footprint of a CTA is compact and neighbouring tiles share lines in L1/L2), derives its
//    four neighbours and weights once, then walks the channels with four independent gathers
//    in flight per step.  Semantics follow reproject's `map_coordinates(order=1,
//    mode='constant', cval=nan)` on an edge-padded image: samples up to half a pixel outside
//    the outermost pixel centres are kept, everything further out is NaN, and a NaN neighbour
//    poisons the sample even at zero weight.  footprint = ~isnan(result).
#include "common.cuh"
#include "tma.cuh"
#include <limits.h>

namespace scb {

int check_cube_args(const float *cube, int64_t nchan, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y);
int env_int(const char *name, int dflt);

struct WcsCel {              // celestial part of a FITS WCS
    double crpix1, crpix2, crval1, crval2, m11, m12, m21, m22, lonpole;
    int sin_proj;            // 0 TAN, 1 SIN
    double i11, i12, i21, i22;   // inverse of m
};

__global__ void __launch_bounds__(256)
wcs_pixel_map_kernel(WcsCel wo, WcsCel wi, int64_t ny, int64_t nx, double *yin, double *xin) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ny * nx) return;
    const int64_t py = g / nx, px = g - py * nx;
    const double D2R = 0.017453292519943295, R2D = 57.29577951308232;
    // ---- output pixel -> native unit vector (paper I eq. 1; paper II eq. 54 / 59) ----
    const double dx = (double)px + 1.0 - wo.crpix1, dy = (double)py + 1.0 - wo.crpix2;
    const double xr = (wo.m11 * dx + wo.m12 * dy) * D2R, yr = (wo.m21 * dx + wo.m22 * dy) * D2R;
    const double r2 = xr * xr + yr * yr;
    double st, ctsp, ctcp;             // sin(theta), cos(theta) sin(phi), cos(theta) cos(phi)
    bool bad = false;
    if (!wo.sin_proj) { st = 1.0 / sqrt(1.0 + r2); ctsp = xr * st; ctcp = -yr * st; }
    else { bad = r2 > 1.0; st = sqrt(fmax(1.0 - r2, 0.0)); ctsp = xr; ctcp = -yr; }
    // ---- native -> celestial (paper II eq. 2), kept as a unit vector about alpha_p ----
    double sp, cp, sd, cd;
    sincos(wo.lonpole * D2R, &sp, &cp);
    sincos(wo.crval2 * D2R, &sd, &cd);
    const double ct_c = ctcp * cp + ctsp * sp;        // cos(theta) cos(phi - phi_p)
    const double ct_s = ctsp * cp - ctcp * sp;        // cos(theta) sin(phi - phi_p)
    const double cz = st * sd + ct_c * cd;            // sin(delta)
    const double cx = st * cd - ct_c * sd;            // cos(delta) cos(alpha - alpha_p)
    const double cy = -ct_s;                          // cos(delta) sin(alpha - alpha_p)
    // ---- shift the longitude origin to the input WCS's alpha_p ----
    double sa, ca;
    sincos((wo.crval1 - wi.crval1) * D2R, &sa, &ca);
    const double ex = cx * ca - cy * sa;              // cos(delta) cos(alpha - alpha_p')
    const double ey = cx * sa + cy * ca;              // cos(delta) sin(alpha - alpha_p')
    // ---- celestial -> native of the input WCS (paper II eq. 5) ----
    double sdi, cdi, spi, cpi;
    sincos(wi.crval2 * D2R, &sdi, &cdi);
    sincos(wi.lonpole * D2R, &spi, &cpi);
    const double st_i = cz * sdi + ex * cdi;          // sin(theta)
    const double ctc_i = cz * cdi - ex * sdi;         // cos(theta) cos(phi - phi_p)
    const double cts_i = -ey;                         // cos(theta) sin(phi - phi_p)
    const double ctsp_i = cts_i * cpi + ctc_i * spi;  // cos(theta) sin(phi)
    const double ctcp_i = ctc_i * cpi - cts_i * spi;  // cos(theta) cos(phi)
    double xi, yi;
    if (!wi.sin_proj) { bad |= st_i <= 0.0; xi = ctsp_i / st_i * R2D; yi = -ctcp_i / st_i * R2D; }
    else { bad |= st_i < 0.0; xi = ctsp_i * R2D; yi = -ctcp_i * R2D; }
    const double qx = wi.i11 * xi + wi.i12 * yi + wi.crpix1 - 1.0;
    const double qy = wi.i21 * xi + wi.i22 * yi + wi.crpix2 - 1.0;
    xin[g] = bad ? nan64() : qx;
    yin[g] = bad ? nan64() : qy;
}

struct ReprojParams {
    const float *in;
    void *out;
    float *out32;            // optional float32 copy of the result
    int *any_valid;          // optional device flag: some output voxel is not NaN
    uint8_t *footprint;
    int64_t nchan, ny_in, nx_in, stride_c, stride_y, ny_out, nx_out;
    const double *yin, *xin;
    float fill;
    int order;
    int chan_per_cta;
    DevMask mask;
};

template <int MODE>
__device__ __forceinline__ float load_filled(const ReprojParams &p, const float *plane, int64_t c, int64_t y, int64_t x) {
    const float v = __ldg(plane + y * p.stride_y + x);
    if (MODE == MODE_NONE) return v;
    return mask_include<MODE>(p.mask, v, c, y, x) ? v : p.fill;
}

template <int MODE, int OUT64>
__global__ void __launch_bounds__(256)
reproject_kernel(const __grid_constant__ ReprojParams p) {
    const int64_t xo = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
    const int64_t yo = (int64_t)blockIdx.y * 8 + (threadIdx.x >> 5);
    if (xo >= p.nx_out || yo >= p.ny_out) return;
    const int64_t o = yo * p.nx_out + xo;
    const double ys = p.yin[o], xs = p.xin[o];
    const bool outside = !(ys >= -0.5 && ys <= (double)p.ny_in - 0.5 && xs >= -0.5 && xs <= (double)p.nx_in - 0.5);
    int64_t r0 = 0, r1 = 0, c0 = 0, c1 = 0;
    double wy0 = 0, wy1 = 0, wx0 = 0, wx1 = 0;
    if (!outside) {
        if (p.order == 0) {
            // nearest neighbour: scipy order 0 rounds half up
            r0 = r1 = min(max((int64_t)floor(ys + 0.5), (int64_t)0), p.ny_in - 1);
            c0 = c1 = min(max((int64_t)floor(xs + 0.5), (int64_t)0), p.nx_in - 1);
            wy0 = 1.0; wx0 = 1.0;
        } else {
            const double fy = floor(ys), fx = floor(xs);
            wy1 = ys - fy; wy0 = 1.0 - wy1;
            wx1 = xs - fx; wx0 = 1.0 - wx1;
            // edge-replicated padding: neighbours clamp to the image
            r0 = min(max((int64_t)fy, (int64_t)0), p.ny_in - 1);
            r1 = min(max((int64_t)fy + 1, (int64_t)0), p.ny_in - 1);
            c0 = min(max((int64_t)fx, (int64_t)0), p.nx_in - 1);
            c1 = min(max((int64_t)fx + 1, (int64_t)0), p.nx_in - 1);
        }
    }
    const double w00 = wy0 * wx0, w01 = wy0 * wx1, w10 = wy1 * wx0, w11 = wy1 * wx1;
    const int64_t cbeg = (int64_t)blockIdx.z * p.chan_per_cta;
    const int64_t cend = min(p.nchan, cbeg + p.chan_per_cta);
    const int64_t plane_out = p.ny_out * p.nx_out;
    bool seen = false;
#pragma unroll 4
    for (int64_t c = cbeg; c < cend; ++c) {
        double res = nan64();
        if (!outside) {
            const float *plane = p.in + c * p.stride_c;
            if (p.order == 0) {
                res = (double)load_filled<MODE>(p, plane, c, r0, c0);
            } else {
                const double v00 = (double)load_filled<MODE>(p, plane, c, r0, c0);
                const double v01 = (double)load_filled<MODE>(p, plane, c, r0, c1);
                const double v10 = (double)load_filled<MODE>(p, plane, c, r1, c0);
                const double v11 = (double)load_filled<MODE>(p, plane, c, r1, c1);
                res = w00 * v00;
                res = fma(w01, v01, res);
                res = fma(w10, v10, res);
                res = fma(w11, v11, res);
            }
        }
        if (OUT64) reinterpret_cast<double *>(p.out)[c * plane_out + o] = res;
        else       reinterpret_cast<float *>(p.out)[c * plane_out + o] = (float)res;
        if (p.out32) p.out32[c * plane_out + o] = (float)res;
        if (p.footprint) p.footprint[c * plane_out + o] = (res == res) ? 1 : 0;
        seen |= res == res;
    }
    if (p.any_valid && seen) *p.any_valid = 1;
}

// ---- tiled kernel --------------------------------------------------------------------------------
// A CTA owns a 32 x 32 tile of output pixels.  The input neighbours of the tile lie in a compact
// (rotated) patch of the input image: its bounding box is found once from the coordinate planes, and when
// it fits RT_BOX x RT_BOX pixels (rotations up to 45 degrees at equal pixel scale) a producer warp streams
// that box for RT_CB channels per stage through a shared-memory ring with ONE tiled TMA request per stage
// (`cp.async.bulk.tensor.3d`; pixels outside the image arrive as zeros and are never addressed, the
// neighbours being clamped to the image first).  The 256 consumer threads keep the four weights and the
// shared-memory offsets of their four output pixels in registers and turn every stage into 4 x RT_CB
// coalesced output rows.  Global gathers (4 uncoalesced loads per output voxel, 16-32 sectors per warp
// load at a 30 degree rotation) become conflict-light LDS; HBM sees each input pixel once per tile
// overlap.  Tiles whose box does not fit (strong magnification) fall back to the direct gathers.
constexpr int RT = 32;                   // output tile edge
constexpr int RT_BOX = 48;               // input box edge
constexpr int RT_CB = 2;                 // channels per stage
constexpr int RT_STAGES = 4;
constexpr int RT_CONSUMERS = 256;
constexpr int RT_THREADS = RT_CONSUMERS + 32;
constexpr int RT_PX = RT * RT / RT_CONSUMERS;      // output pixels per thread

struct ReprojSmem {
    float data[RT_STAGES][RT_CB][RT_BOX][RT_BOX];
    uint64_t full[RT_STAGES];
    uint64_t empty[RT_STAGES];
    int bb[4];                           // xmin, ymin, xmax, ymax of the neighbours
};

template <int MODE, int OUT64>
__global__ void __launch_bounds__(RT_THREADS)
reproject_tiled_kernel(const __grid_constant__ ReprojParams p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ReprojSmem &sm = *reinterpret_cast<ReprojSmem *>(smem_raw);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < RT_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], RT_CONSUMERS / 32); }
        mbar_fence_init();
        sm.bb[0] = INT_MAX; sm.bb[1] = INT_MAX; sm.bb[2] = INT_MIN; sm.bb[3] = INT_MIN;
    }
    __syncthreads();

    // ---- per output pixel: neighbours and weights (consumers) ----
    const int64_t xo = (int64_t)blockIdx.x * RT + lane;
    int r0[RT_PX], c0[RT_PX];            // (image sides are far below 2^31)
    int dr[RT_PX], dc[RT_PX];            // r1 - r0, c1 - c0 (0 or 1)
    double w00[RT_PX], w01[RT_PX], w10[RT_PX], w11[RT_PX];
    bool outside[RT_PX];
    int xmin = INT_MAX, ymin = INT_MAX, xmax = INT_MIN, ymax = INT_MIN;
    if (warp < RT_CONSUMERS / 32) {
#pragma unroll
        for (int k = 0; k < RT_PX; ++k) {
            const int64_t yo = (int64_t)blockIdx.y * RT + warp + 8 * k;
            outside[k] = true; r0[k] = c0[k] = 0; dr[k] = dc[k] = 0;
            w00[k] = w01[k] = w10[k] = w11[k] = 0.0;
            if (xo < p.nx_out && yo < p.ny_out) {
                const int64_t o = yo * p.nx_out + xo;
                const double ys = p.yin[o], xs = p.xin[o];
                outside[k] = !(ys >= -0.5 && ys <= (double)p.ny_in - 0.5 && xs >= -0.5 && xs <= (double)p.nx_in - 0.5);
                if (!outside[k]) {
                    int64_t r1, c1;
                    if (p.order == 0) {
                        // nearest neighbour: scipy order 0 rounds half up
                        r1 = min(max((int64_t)floor(ys + 0.5), (int64_t)0), p.ny_in - 1);
                        c1 = min(max((int64_t)floor(xs + 0.5), (int64_t)0), p.nx_in - 1);
                        r0[k] = (int)r1; c0[k] = (int)c1;
                        w00[k] = 1.0;
                    } else {
                        const double fy = floor(ys), fx = floor(xs);
                        const double wy1 = ys - fy, wy0 = 1.0 - wy1, wx1 = xs - fx, wx0 = 1.0 - wx1;
                        // edge-replicated padding: neighbours clamp to the image
                        r0[k] = (int)min(max((int64_t)fy, (int64_t)0), p.ny_in - 1);
                        r1 = min(max((int64_t)fy + 1, (int64_t)0), p.ny_in - 1);
                        c0[k] = (int)min(max((int64_t)fx, (int64_t)0), p.nx_in - 1);
                        c1 = min(max((int64_t)fx + 1, (int64_t)0), p.nx_in - 1);
                        w00[k] = wy0 * wx0; w01[k] = wy0 * wx1; w10[k] = wy1 * wx0; w11[k] = wy1 * wx1;
                    }
                    dr[k] = (int)(r1 - r0[k]); dc[k] = (int)(c1 - c0[k]);
                    xmin = min(xmin, c0[k]); xmax = max(xmax, (int)c1);
                    ymin = min(ymin, r0[k]); ymax = max(ymax, (int)r1);
                }
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, d)); ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, d));
            xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, d)); ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, d));
        }
        if (lane == 0 && xmax >= xmin) {
            atomicMin(&sm.bb[0], xmin); atomicMin(&sm.bb[1], ymin); atomicMax(&sm.bb[2], xmax); atomicMax(&sm.bb[3], ymax);
        }
    }
    __syncthreads();
    // the box must start on a 16-byte boundary of the input row (the TMA unit traps otherwise)
    const int bx0 = sm.bb[0] & ~3, by0 = sm.bb[1];
    const bool any_inside = sm.bb[2] >= sm.bb[0];
    const bool fits = any_inside && sm.bb[2] - bx0 < RT_BOX && sm.bb[3] - by0 < RT_BOX;

    const int64_t cbeg = (int64_t)blockIdx.z * p.chan_per_cta;
    const int64_t cend = min(p.nchan, cbeg + p.chan_per_cta);
    const int nstage = (int)((cend - cbeg + RT_CB - 1) / RT_CB);
    const int64_t plane_out = p.ny_out * p.nx_out;

    if (warp == RT_CONSUMERS / 32) {
        // ---------------- producer warp ----------------
        if (!fits) return;
        const uint64_t pol = l2_evict_first_policy();
        for (int j = 0; j < nstage; ++j) {
            const int s = j % RT_STAGES;
            if (j >= RT_STAGES) mbar_wait(&sm.empty[s], ((j / RT_STAGES) - 1) & 1);
            if (lane == 0) {
                mbar_expect_tx(&sm.full[s], (uint32_t)(RT_CB * RT_BOX * RT_BOX * 4));
                tma_load_box3d(&sm.data[s][0][0][0], &tmap, bx0, by0, (int)(cbeg + (int64_t)j * RT_CB), &sm.full[s], pol);
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    bool seen = false;
    if (xo >= p.nx_out) return;                                      // (no block-wide barrier follows)
    if (fits) {
        int off[RT_PX];
#pragma unroll
        for (int k = 0; k < RT_PX; ++k) off[k] = outside[k] ? 0 : (r0[k] - by0) * RT_BOX + (c0[k] - bx0);
        for (int j = 0; j < nstage; ++j) {
            const int s = j % RT_STAGES;
            mbar_wait(&sm.full[s], (j / RT_STAGES) & 1);
#pragma unroll
            for (int cc = 0; cc < RT_CB; ++cc) {
                const int64_t c = cbeg + (int64_t)j * RT_CB + cc;
                if (c < cend) {
                    const float *plane = &sm.data[s][cc][0][0];
#pragma unroll
                    for (int k = 0; k < RT_PX; ++k) {
                        const int64_t yo = (int64_t)blockIdx.y * RT + warp + 8 * k;
                        if (yo < p.ny_out) {
                            double res = nan64();
                            if (!outside[k]) {
                                const float *q = plane + off[k];
                                float v00 = q[0], v01 = q[dc[k]], v10 = q[dr[k] * RT_BOX], v11 = q[dr[k] * RT_BOX + dc[k]];
                                if (MODE != MODE_NONE) {
                                    if (!mask_include<MODE>(p.mask, v00, c, r0[k], c0[k])) v00 = p.fill;
                                    if (!mask_include<MODE>(p.mask, v01, c, r0[k], c0[k] + dc[k])) v01 = p.fill;
                                    if (!mask_include<MODE>(p.mask, v10, c, r0[k] + dr[k], c0[k])) v10 = p.fill;
                                    if (!mask_include<MODE>(p.mask, v11, c, r0[k] + dr[k], c0[k] + dc[k])) v11 = p.fill;
                                }
                                if (p.order == 0) res = (double)v00;
                                else {
                                    res = w00[k] * (double)v00;
                                    res = fma(w01[k], (double)v01, res);
                                    res = fma(w10[k], (double)v10, res);
                                    res = fma(w11[k], (double)v11, res);
                                }
                            }
                            const int64_t o = c * plane_out + yo * p.nx_out + xo;
                            if (OUT64) reinterpret_cast<double *>(p.out)[o] = res;
                            else       reinterpret_cast<float *>(p.out)[o] = (float)res;
                            if (p.out32) p.out32[o] = (float)res;
                            if (p.footprint) p.footprint[o] = (res == res) ? 1 : 0;
                            seen |= res == res;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[s]);
        }
        if (p.any_valid && seen) *p.any_valid = 1;
        return;
    }
    // ---- the box does not fit (or nothing of the tile falls inside the image): direct gathers ----
    for (int64_t c = cbeg; c < cend; ++c) {
        const float *plane = p.in + c * p.stride_c;
#pragma unroll
        for (int k = 0; k < RT_PX; ++k) {
            const int64_t yo = (int64_t)blockIdx.y * RT + warp + 8 * k;
            if (yo >= p.ny_out) continue;
            double res = nan64();
            if (!outside[k]) {
                if (p.order == 0) {
                    res = (double)load_filled<MODE>(p, plane, c, r0[k], c0[k]);
                } else {
                    const double v00 = (double)load_filled<MODE>(p, plane, c, r0[k], c0[k]);
                    const double v01 = (double)load_filled<MODE>(p, plane, c, r0[k], c0[k] + dc[k]);
                    const double v10 = (double)load_filled<MODE>(p, plane, c, r0[k] + dr[k], c0[k]);
                    const double v11 = (double)load_filled<MODE>(p, plane, c, r0[k] + dr[k], c0[k] + dc[k]);
                    res = w00[k] * v00;
                    res = fma(w01[k], v01, res);
                    res = fma(w10[k], v10, res);
                    res = fma(w11[k], v11, res);
                }
            }
            const int64_t o = c * plane_out + yo * p.nx_out + xo;
            if (OUT64) reinterpret_cast<double *>(p.out)[o] = res;
            else       reinterpret_cast<float *>(p.out)[o] = (float)res;
            if (p.out32) p.out32[o] = (float)res;
            if (p.footprint) p.footprint[o] = (res == res) ? 1 : 0;
            seen |= res == res;
        }
    }
    if (p.any_valid && seen) *p.any_valid = 1;
}

template <int MODE, int OUT64>
static cudaError_t launch_reproject_tiled(const ReprojParams &p, const CUtensorMap &tmap, dim3 grid, cudaStream_t s) {
    auto kern = reproject_tiled_kernel<MODE, OUT64>;
    static unsigned long long configured = 0;        // per instantiation, one bit per device
    if (cudaError_t e = ensure_dyn_smem(kern, sizeof(ReprojSmem), &configured)) return e;
    kern<<<grid, RT_THREADS, sizeof(ReprojSmem), s>>>(p, tmap);
    return cudaGetLastError();
}

static bool unpack_wcs(const double *w, WcsCel *out) {
    out->crpix1 = w[0]; out->crpix2 = w[1]; out->crval1 = w[2]; out->crval2 = w[3];
    out->m11 = w[4]; out->m12 = w[5]; out->m21 = w[6]; out->m22 = w[7];
    out->lonpole = w[8]; out->sin_proj = w[9] != 0.0 ? 1 : 0;
    const double det = out->m11 * out->m22 - out->m12 * out->m21;
    if (det == 0.0 || det != det) return false;
    out->i11 = out->m22 / det; out->i12 = -out->m12 / det; out->i21 = -out->m21 / det; out->i22 = out->m11 / det;
    return true;
}

}  // namespace scb

using namespace scb;

extern "C" int sc_wcs_pixel_map(const double *wcs_out, const double *wcs_in,
                                int64_t ny_out, int64_t nx_out, double *yin, double *xin, void *stream) {
    SC_CHECK_ARG(wcs_out && wcs_in && yin && xin, "NULL argument");
    SC_CHECK_ARG(ny_out > 0 && nx_out > 0, "bad output shape");
    WcsCel wo, wi;
    SC_CHECK_ARG(unpack_wcs(wcs_out, &wo) && unpack_wcs(wcs_in, &wi), "singular pixel scale matrix");
    cudaStream_t s = (cudaStream_t)stream;
    LaunchScope ls(0, s);
    wcs_pixel_map_kernel<<<(unsigned)cdiv(ny_out * nx_out, 256), 256, 0, s>>>(wo, wi, ny_out, nx_out, yin, xin);
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}

extern "C" int sc_reproject(const float *in, void *out, int out_dtype, uint8_t *footprint,
                            int64_t nchan, int64_t ny_in, int64_t nx_in,
                            int64_t stride_c, int64_t stride_y,
                            int64_t ny_out, int64_t nx_out,
                            const sc_mask_desc *mask, double fill,
                            const double *yin, const double *xin, int order, void *stream) {
    return sc_reproject_ex(in, out, out_dtype, nullptr, footprint, nullptr, nchan, ny_in, nx_in, stride_c, stride_y,
                           ny_out, nx_out, mask, fill, yin, xin, order, stream);
}

extern "C" int sc_reproject_ex(const float *in, void *out, int out_dtype, float *out_f32, uint8_t *footprint, int *any_valid,
                               int64_t nchan, int64_t ny_in, int64_t nx_in,
                               int64_t stride_c, int64_t stride_y,
                               int64_t ny_out, int64_t nx_out,
                               const sc_mask_desc *mask, double fill,
                               const double *yin, const double *xin, int order, void *stream) {
    int rc = check_cube_args(in, nchan, ny_in, nx_in, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(out && yin && xin, "out / yin / xin must not be NULL");
    SC_CHECK_ARG(ny_out > 0 && nx_out > 0, "bad output shape");
    SC_CHECK_ARG(out_dtype == SC_F32 || out_dtype == SC_F64, "out_dtype must be SC_F32 or SC_F64");
    SC_CHECK_ARG(order == 0 || order == 1, "only nearest-neighbor (0) and bilinear (1) are implemented");
    ReprojParams p{};
    p.in = in; p.out = out; p.footprint = footprint; p.out32 = out_f32; p.any_valid = any_valid;
    if (any_valid) SC_CUDA(cudaMemsetAsync(any_valid, 0, sizeof(int), (cudaStream_t)stream));
    p.nchan = nchan; p.ny_in = ny_in; p.nx_in = nx_in; p.stride_c = stride_c; p.stride_y = stride_y;
    p.ny_out = ny_out; p.nx_out = nx_out; p.yin = yin; p.xin = xin; p.fill = (float)fill; p.order = order;
    rc = build_dev_mask(mask, in, stride_c, stride_y, &p.mask);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int m = p.mask.mode;
    const bool aligned = ((uintptr_t)in % 16 == 0) && stride_c % 4 == 0 && stride_y % 4 == 0;
    if (aligned && env_int("SC_REPROJECT_KERNEL", 0) != 1) {
        // tiled kernel: 32 x 32 output tiles, the input box through a TMA ring
        CUtensorMap tmap;
        rc = make_cube_tensor_map3(&tmap, in, nchan, ny_in, nx_in, stride_c, stride_y, RT_BOX, RT_BOX, RT_CB);
        if (rc) return rc;
        const int64_t tiles = cdiv(nx_out, RT) * cdiv(ny_out, RT);
        int64_t zchunks = 1;
        while (tiles * zchunks < 148 * 6 && zchunks * 2 <= nchan && nchan / (zchunks * 2) >= 8) zchunks *= 2;
        if (zchunks > 65535) zchunks = 65535;
        p.chan_per_cta = (int)(cdiv(cdiv(nchan, zchunks), RT_CB) * RT_CB);
        // channels per CTA also decide what the L2 can do for the overlap of neighbouring tiles' input boxes: a row of output
        // tiles streams tiles_x * box * chan_per_cta bytes before the next row (whose boxes overlap it) starts
        const int cap = env_int("SC_REPROJECT_CHAN", 0);
        if (cap > 0 && p.chan_per_cta > cap) p.chan_per_cta = (int)(cdiv(cap, RT_CB) * RT_CB);
        SC_CHECK_ARG(cdiv(nchan, p.chan_per_cta) <= 65535, "too many channel chunks");
        dim3 grid((unsigned)cdiv(nx_out, RT), (unsigned)cdiv(ny_out, RT), (unsigned)cdiv(nchan, p.chan_per_cta));
        SC_CHECK_ARG(grid.y <= 65535, "output image too tall for one launch");
        LaunchScope ls(SC_OP_REPROJECT, s);
        cudaError_t e;
        if (out_dtype == SC_F64) {
            if (m == MODE_NONE) e = launch_reproject_tiled<MODE_NONE, 1>(p, tmap, grid, s);
            else if (m == MODE_INTERVAL) e = launch_reproject_tiled<MODE_INTERVAL, 1>(p, tmap, grid, s);
            else e = launch_reproject_tiled<MODE_GENERIC, 1>(p, tmap, grid, s);
        } else {
            if (m == MODE_NONE) e = launch_reproject_tiled<MODE_NONE, 0>(p, tmap, grid, s);
            else if (m == MODE_INTERVAL) e = launch_reproject_tiled<MODE_INTERVAL, 0>(p, tmap, grid, s);
            else e = launch_reproject_tiled<MODE_GENERIC, 0>(p, tmap, grid, s);
        }
        if (e != cudaSuccess) return cuda_fail(e, "reproject_tiled_kernel launch");
        return SC_OK;
    }
    // channel chunks: enough CTAs for the chip, long enough runs to amortise the weight set-up
    const int64_t tiles = cdiv(nx_out, 32) * cdiv(ny_out, 8);
    int64_t zchunks = 1;
    while (tiles * zchunks < 148 * 16 && zchunks * 2 <= nchan && nchan / (zchunks * 2) >= 8) zchunks *= 2;
    if (zchunks > 65535) zchunks = 65535;
    p.chan_per_cta = (int)cdiv(nchan, zchunks);
    dim3 grid((unsigned)cdiv(nx_out, 32), (unsigned)cdiv(ny_out, 8), (unsigned)cdiv(nchan, p.chan_per_cta));
    SC_CHECK_ARG(grid.y <= 65535, "output image too tall for one launch");
    LaunchScope ls(SC_OP_REPROJECT, s);
    if (out_dtype == SC_F64) {
        if (m == MODE_NONE) reproject_kernel<MODE_NONE, 1><<<grid, 256, 0, s>>>(p);
        else if (m == MODE_INTERVAL) reproject_kernel<MODE_INTERVAL, 1><<<grid, 256, 0, s>>>(p);
        else reproject_kernel<MODE_GENERIC, 1><<<grid, 256, 0, s>>>(p);
    } else {
        if (m == MODE_NONE) reproject_kernel<MODE_NONE, 0><<<grid, 256, 0, s>>>(p);
        else if (m == MODE_INTERVAL) reproject_kernel<MODE_INTERVAL, 0><<<grid, 256, 0, s>>>(p);
        else reproject_kernel<MODE_GENERIC, 0><<<grid, 256, 0, s>>>(p);
    }
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}
