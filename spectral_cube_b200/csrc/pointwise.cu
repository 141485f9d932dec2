// pointwise.cu -- the in-place epilogue of `convolve_to` (spectral_cube.py:3369-3378, :4207-4233):
//   * values in Jy/beam are multiplied by target.sr / beam.sr after each channel image has been convolved;
//   * with the numpy class's default `convolve_fft`, output pixels whose kernel window holds no valid input
//     come out as 0.0 (astropy zeroes results whose interpolation weight is below 10 eps) where the direct
//     `convolve` -- and the device kernels -- keep the NaN: `nan_to_zero` applies that rule;
//   * channel planes that `_apply_spatial_function` copies through (nothing included by the mask, :161-172)
//     are left alone: `skip_planes[c] != 0`.
// One pass, 4 B read + 4 B written per voxel (8 + 8 for float64), HBM-bound; 16-byte vector accesses when
// pointer, pitch and row length allow it, a scalar grid-stride loop otherwise.
#include "common.cuh"

namespace scb {

template <typename T>
__device__ __forceinline__ T scale_one(T v, double factor, int nan_to_zero) {
    if (v != v) return nan_to_zero ? (T)0 : v;
    return (T)((double)v * factor);
}

template <typename T>
__global__ void __launch_bounds__(256)
scale_rows_kernel(T *__restrict__ data, int64_t rows, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y,
                  double factor, int nan_to_zero, const uint8_t *__restrict__ skip_planes) {
    const int64_t total = rows * nx;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        const int64_t r = i / nx, x = i - r * nx;
        const int64_t c = r / ny, y = r - c * ny;
        if (skip_planes && skip_planes[c]) continue;
        T *p = data + c * stride_c + y * stride_y + x;
        *p = scale_one<T>(*p, factor, nan_to_zero);
    }
}

// float32 rows whose length, pitch and base address are multiples of 16 bytes: nx4 = nx / 4 float4 per row
__global__ void __launch_bounds__(256)
scale_rows_f32x4_kernel(float *__restrict__ data, int64_t rows, int64_t ny, int64_t nx4, int64_t stride_c, int64_t stride_y,
                        double factor, int nan_to_zero, const uint8_t *__restrict__ skip_planes) {
    const int64_t total = rows * nx4;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        const int64_t r = i / nx4, x4 = i - r * nx4;
        const int64_t c = r / ny, y = r - c * ny;
        if (skip_planes && skip_planes[c]) continue;
        float4 *p = reinterpret_cast<float4 *>(data + c * stride_c + y * stride_y) + x4;
        float4 v = *p;
        v.x = scale_one<float>(v.x, factor, nan_to_zero); v.y = scale_one<float>(v.y, factor, nan_to_zero);
        v.z = scale_one<float>(v.z, factor, nan_to_zero); v.w = scale_one<float>(v.w, factor, nan_to_zero);
        *p = v;
    }
}

// ---- mosaic_cubes (cube_utils.py:791-856): the two elementwise steps around the reprojections ----------------
//   acc[c] += nan_to_num(reprojected[c]) for every channel (:838-841), weight += footprint of channel 0 (:834-836);
//   at the end acc[c] /= weight (:847-849; x / 0 and 0 / 0 follow IEEE like numpy under errstate(divide='ignore')).
// float64 accumulators (the reference's `np.zeros(shape_opt)`), 16 + 8 B/voxel of traffic, HBM-bound.
__device__ __forceinline__ double nan_to_num64(double v) {
    if (v != v) return 0.0;
    if (v == INFINITY) return DBL_MAX;                 // numpy.nan_to_num: +-inf -> the largest finite float64
    if (v == -INFINITY) return -DBL_MAX;
    return v;
}

template <typename T>
__global__ void __launch_bounds__(256)
mosaic_accumulate_kernel(double *__restrict__ acc, double *__restrict__ weight, const T *__restrict__ src,
                         const uint8_t *__restrict__ footprint0, int64_t plane, int64_t total) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
        acc[i] += nan_to_num64((double)src[i]);
        if (i < plane && weight) weight[i] += footprint0[i] ? 1.0 : 0.0;
    }
}

__global__ void __launch_bounds__(256)
mosaic_normalize_kernel(double *__restrict__ acc, const double *__restrict__ weight, int64_t plane, int64_t total) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) acc[i] = acc[i] / weight[i % plane];
}

}  // namespace scb

using namespace scb;

extern "C" int sc_mosaic_accumulate(double *acc, double *weight, const void *reprojected, int dtype, const uint8_t *footprint0,
                                    int64_t nchan, int64_t ny, int64_t nx, void *stream) {
    SC_CHECK_ARG(acc && reprojected, "NULL buffer");
    SC_CHECK_ARG(dtype == SC_F32 || dtype == SC_F64, "reprojected data must be float32 or float64 (dtype %d)", dtype);
    SC_CHECK_ARG(nchan > 0 && ny > 0 && nx > 0, "bad shape");
    SC_CHECK_ARG(!weight || footprint0, "a weight plane needs the footprint of channel 0");
    const int64_t plane = ny * nx, total = nchan * plane;
    cudaStream_t s = (cudaStream_t)stream;
    LaunchScope ls(SC_OP_POINTWISE, s);
    const int64_t cap = 148 * 16, want = cdiv(total, 256);
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (dtype == SC_F64) mosaic_accumulate_kernel<double><<<grid, 256, 0, s>>>(acc, weight, (const double *)reprojected, footprint0, plane, total);
    else                 mosaic_accumulate_kernel<float><<<grid, 256, 0, s>>>(acc, weight, (const float *)reprojected, footprint0, plane, total);
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}

extern "C" int sc_mosaic_normalize(double *acc, const double *weight, int64_t nchan, int64_t ny, int64_t nx, void *stream) {
    SC_CHECK_ARG(acc && weight, "NULL buffer");
    SC_CHECK_ARG(nchan > 0 && ny > 0 && nx > 0, "bad shape");
    const int64_t plane = ny * nx, total = nchan * plane;
    cudaStream_t s = (cudaStream_t)stream;
    LaunchScope ls(SC_OP_POINTWISE, s);
    const int64_t cap = 148 * 16, want = cdiv(total, 256);
    mosaic_normalize_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(acc, weight, plane, total);
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}

extern "C" int sc_scale(void *data, int dtype, int64_t nchan, int64_t ny, int64_t nx,
                        int64_t stride_c, int64_t stride_y, double factor, int nan_to_zero,
                        const uint8_t *skip_planes, void *stream) {
    SC_CHECK_ARG(data != nullptr, "NULL buffer");
    SC_CHECK_ARG(dtype == SC_F32 || dtype == SC_F64, "sc_scale works on float32 or float64 data (dtype %d)", dtype);
    SC_CHECK_ARG(nchan >= 0 && ny >= 0 && nx >= 0, "negative shape");
    SC_CHECK_ARG(stride_y >= nx && (nchan <= 1 || stride_c >= ny * stride_y), "strides smaller than the shape");
    const int64_t rows = nchan * ny;
    if (rows == 0 || nx == 0 || (factor == 1.0 && !nan_to_zero)) return SC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    LaunchScope ls(SC_OP_POINTWISE, s);
    const int64_t cap = 148 * 16;                  // grid-stride: 16 CTAs of 256 threads per SM
    if (dtype == SC_F32 && nx % 4 == 0 && stride_c % 4 == 0 && stride_y % 4 == 0 && (uintptr_t)data % 16 == 0) {
        const int64_t want = cdiv(rows * (nx / 4), 256);
        scale_rows_f32x4_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(
            (float *)data, rows, ny, nx / 4, stride_c, stride_y, factor, nan_to_zero, skip_planes);
    } else {
        const int64_t want = cdiv(rows * nx, 256);
        const unsigned grid = (unsigned)(want < cap ? want : cap);
        if (dtype == SC_F32) scale_rows_kernel<float><<<grid, 256, 0, s>>>((float *)data, rows, ny, nx, stride_c, stride_y, factor, nan_to_zero, skip_planes);
        else                 scale_rows_kernel<double><<<grid, 256, 0, s>>>((double *)data, rows, ny, nx, stride_c, stride_y, factor, nan_to_zero, skip_planes);
    }
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}
