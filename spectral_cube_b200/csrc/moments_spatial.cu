// moments_spatial.cu -- moment0/1/2 along a SPATIAL axis (numpy axis 1 = y, axis 2 = x), and the
// cumulative angular pixel offsets they are taken against.
//
// Same reference functions as the spectral case (_moments.py:30-202) with
// pix_cen[axis] = cumulative great-circle offsets from the cube face (spectral_cube.py:1477-1492)
// and pix_size = _pix_size_slice(axis) (:1530-1533).  Goldens: tests/test_moments.py:22-48.
//  * axis 1: a thread owns one (channel, x) column and walks y -- a warp reads 128 contiguous
//    bytes per row.
//  * axis 2: a warp owns one (channel, y) row, lanes stride over x (coalesced) and the partial
//    sums are combined with warp shuffles in a fixed butterfly order.
// float64 raw sums about the offset of the central pixel; finalised like the spectral kernel.
#include "common.cuh"

namespace scb {

int check_cube_args(const float *cube, int64_t nchan, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y);

struct SpMomParams {
    const float *cube;
    int64_t nchan, ny, nx, stride_c, stride_y;
    const double *off;        // (ny, nx) cumulative offsets along the reduced axis
    double K, pix_size;
    double *m0, *m1, *m2;
    DevMask mask;
};

__device__ __forceinline__ void sp_finalize(const SpMomParams &p, int64_t o, double s0, double s1, double s2, int cnt) {
    const bool any = cnt > 0;
    const double mean = s1 / s0;
    if (p.m0) p.m0[o] = any ? s0 * p.pix_size : nan64();
    if (p.m1) p.m1[o] = any ? p.K + mean : nan64();
    if (p.m2) p.m2[o] = any ? ((cnt == 1 && s0 != 0.0) ? 0.0 : s2 / s0 - mean * mean) : nan64();
}

template <int MODE>
__global__ void __launch_bounds__(128)
moments_axis1_kernel(const __grid_constant__ SpMomParams p) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.nchan * p.nx) return;
    const int64_t c = g / p.nx, x = g - c * p.nx;
    const float *src = p.cube + c * p.stride_c + x;
    double s0 = 0, s1 = 0, s2 = 0;
    int cnt = 0;
    for (int64_t y = 0; y < p.ny; ++y) {
        const float f = ldg_stream1(src + y * p.stride_y);
        bool inc = mask_include<MODE>(p.mask, f, c, y, x);
        if (MODE != MODE_INTERVAL) inc = inc & (f == f);
        if (inc) {
            const double w = (double)f, d = __ldg(p.off + y * p.nx + x) - p.K;
            s0 += w; s1 = fma(w, d, s1); s2 = fma(w * d, d, s2); cnt += 1;
        }
    }
    sp_finalize(p, g, s0, s1, s2, cnt);
}

template <int MODE>
__global__ void __launch_bounds__(128)
moments_axis2_kernel(const __grid_constant__ SpMomParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= p.nchan * p.ny) return;
    const int64_t c = row / p.ny, y = row - c * p.ny;
    const float *src = p.cube + c * p.stride_c + y * p.stride_y;
    double s0 = 0, s1 = 0, s2 = 0;
    int cnt = 0;
    for (int64_t x = lane; x < p.nx; x += 32) {
        const float f = ldg_stream1(src + x);
        bool inc = mask_include<MODE>(p.mask, f, c, y, x);
        if (MODE != MODE_INTERVAL) inc = inc & (f == f);
        if (inc) {
            const double w = (double)f, d = __ldg(p.off + y * p.nx + x) - p.K;
            s0 += w; s1 = fma(w, d, s1); s2 = fma(w * d, d, s2); cnt += 1;
        }
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, sh);
        s1 += __shfl_xor_sync(0xffffffffu, s1, sh);
        s2 += __shfl_xor_sync(0xffffffffu, s2, sh);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, sh);
    }
    if (lane == 0) sp_finalize(p, row, s0, s1, s2, cnt);
}

// ---- pixel offsets (spectral_cube.py:1455-1492) -------------------------------------------------------
struct OffWcs { double crpix1, crpix2, crval1, crval2, m11, m12, m21, m22, lonpole; int sin_proj; };

__global__ void __launch_bounds__(256)
lonlat_kernel(OffWcs w, int64_t ny, int64_t nx, double *lon, double *lat) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ny * nx) return;
    const int64_t py = g / nx, px = g - py * nx;
    const double D2R = 0.017453292519943295;
    const double dx = (double)px + 1.0 - w.crpix1, dy = (double)py + 1.0 - w.crpix2;
    const double xr = (w.m11 * dx + w.m12 * dy) * D2R, yr = (w.m21 * dx + w.m22 * dy) * D2R;
    const double r2 = xr * xr + yr * yr;
    double st, ctsp, ctcp;
    if (!w.sin_proj) { st = 1.0 / sqrt(1.0 + r2); ctsp = xr * st; ctcp = -yr * st; }
    else { st = r2 > 1.0 ? nan64() : sqrt(fmax(1.0 - r2, 0.0)); ctsp = xr; ctcp = -yr; }
    double sp, cp, sd, cd;
    sincos(w.lonpole * D2R, &sp, &cp);
    sincos(w.crval2 * D2R, &sd, &cd);
    const double ct_c = ctcp * cp + ctsp * sp, ct_s = ctsp * cp - ctcp * sp;
    const double cz = st * sd + ct_c * cd, cx = st * cd - ct_c * sd, cy = -ct_s;
    lon[g] = w.crval1 * D2R + atan2(cy, cx);          // radians
    lat[g] = atan2(cz, hypot(cx, cy));
}

// astropy.coordinates.angular_separation (Vincenty), radians
__device__ __forceinline__ double angsep(double lon1, double lat1, double lon2, double lat2) {
    double sdlon, cdlon, slat1, clat1, slat2, clat2;
    sincos(lon2 - lon1, &sdlon, &cdlon);
    sincos(lat1, &slat1, &clat1);
    sincos(lat2, &slat2, &clat2);
    const double num1 = clat2 * sdlon, num2 = clat1 * slat2 - slat1 * clat2 * cdlon;
    const double den = slat1 * slat2 + clat1 * clat2 * cdlon;
    return atan2(hypot(num1, num2), den);
}

// axis 2 offsets: x[y,0] = 0, x[y,i] = x[y,i-1] + deg(sep((lon,lat)[y,i-1], (lon[y,i], lat[y,i-1])))
// axis 1 offsets: y[0,x] = 0, y[j,x] = y[j-1,x] + deg(sep((lon,lat)[j-1,x], (lon,lat)[j,x]))
__global__ void __launch_bounds__(128)
cum_offsets_kernel(const double *lon, const double *lat, int64_t ny, int64_t nx, int axis, double *off) {
    const double R2D = 57.29577951308232;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (axis == 2) {
        if (t >= ny) return;
        double acc = 0.0;
        off[t * nx] = 0.0;
        for (int64_t i = 1; i < nx; ++i) {
            const int64_t a = t * nx + i - 1, b = t * nx + i;
            acc += angsep(lon[a], lat[a], lon[b], lat[a]) * R2D;
            off[b] = acc;
        }
    } else {
        if (t >= nx) return;
        double acc = 0.0;
        off[t] = 0.0;
        for (int64_t j = 1; j < ny; ++j) {
            const int64_t a = (j - 1) * nx + t, b = j * nx + t;
            acc += angsep(lon[a], lat[a], lon[b], lat[b]) * R2D;
            off[b] = acc;
        }
    }
}

}  // namespace scb

using namespace scb;

extern "C" int sc_pixel_offsets(const double *wcs, int64_t ny, int64_t nx, int axis, double *offsets,
                                void *workspace, size_t workspace_bytes, void *stream) {
    SC_CHECK_ARG(wcs && offsets, "NULL argument");
    SC_CHECK_ARG(ny > 0 && nx > 0, "bad shape");
    SC_CHECK_ARG(axis == 1 || axis == 2, "axis must be 1 or 2");
    const size_t need = (size_t)ny * nx * 16 + 256;
    if (!workspace || workspace_bytes < need) {
        set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
        return SC_ERR_WORKSPACE;
    }
    OffWcs w{wcs[0], wcs[1], wcs[2], wcs[3], wcs[4], wcs[5], wcs[6], wcs[7], wcs[8], wcs[9] != 0.0 ? 1 : 0};
    double *lon = (double *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    double *lat = lon + ny * nx;
    cudaStream_t s = (cudaStream_t)stream;
    LaunchScope ls(0, s);
    lonlat_kernel<<<(unsigned)cdiv(ny * nx, 256), 256, 0, s>>>(w, ny, nx, lon, lat);
    SC_CUDA(cudaGetLastError());
    const int64_t n = axis == 2 ? ny : nx;
    cum_offsets_kernel<<<(unsigned)cdiv(n, 128), 128, 0, s>>>(lon, lat, ny, nx, axis, offsets);
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}

extern "C" int sc_moments_spatial(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                                  int64_t stride_c, int64_t stride_y, int axis,
                                  const sc_mask_desc *mask, const double *offsets, double pix_size,
                                  int want_bits, double *out_m0, double *out_m1, double *out_m2, void *stream) {
    int rc = check_cube_args(cube, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(axis == 1 || axis == 2, "axis must be 1 or 2");
    SC_CHECK_ARG(offsets != nullptr, "offsets is NULL");
    SC_CHECK_ARG(want_bits > 0 && want_bits <= 7, "want_bits=%d must be a combination of 1|2|4", want_bits);
    SC_CHECK_ARG(!(want_bits & SC_WANT_M0) || out_m0, "out_m0 is NULL but moment 0 was requested");
    SC_CHECK_ARG(!(want_bits & SC_WANT_M1) || out_m1, "out_m1 is NULL but moment 1 was requested");
    SC_CHECK_ARG(!(want_bits & SC_WANT_M2) || out_m2, "out_m2 is NULL but moment 2 was requested");
    SpMomParams p{};
    p.cube = cube; p.nchan = nchan; p.ny = ny; p.nx = nx; p.stride_c = stride_c; p.stride_y = stride_y;
    p.off = offsets; p.pix_size = pix_size;
    p.m0 = (want_bits & SC_WANT_M0) ? out_m0 : nullptr;
    p.m1 = (want_bits & SC_WANT_M1) ? out_m1 : nullptr;
    p.m2 = (want_bits & SC_WANT_M2) ? out_m2 : nullptr;
    rc = build_dev_mask(mask, cube, stride_c, stride_y, &p.mask);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    SC_CUDA(cudaMemcpyAsync(&p.K, offsets + (ny / 2) * nx + nx / 2, sizeof(double), cudaMemcpyDeviceToHost, s));
    SC_CUDA(cudaStreamSynchronize(s));
    LaunchScope ls(SC_OP_MOMENTS, s);
    const int m = p.mask.mode;
    if (axis == 1) {
        const unsigned grid = (unsigned)cdiv(nchan * nx, 128);
        if (m == MODE_NONE) moments_axis1_kernel<MODE_NONE><<<grid, 128, 0, s>>>(p);
        else if (m == MODE_INTERVAL) moments_axis1_kernel<MODE_INTERVAL><<<grid, 128, 0, s>>>(p);
        else moments_axis1_kernel<MODE_GENERIC><<<grid, 128, 0, s>>>(p);
    } else {
        const unsigned grid = (unsigned)cdiv(nchan * ny, 4);
        if (m == MODE_NONE) moments_axis2_kernel<MODE_NONE><<<grid, 128, 0, s>>>(p);
        else if (m == MODE_INTERVAL) moments_axis2_kernel<MODE_INTERVAL><<<grid, 128, 0, s>>>(p);
        else moments_axis2_kernel<MODE_GENERIC><<<grid, 128, 0, s>>>(p);
    }
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}
