// maskops.cu -- materialise the mask predicate / the filled data (masks.py:143-237,
// base_class.py:389-417) and pack filled edge rows for the halo exchange.
#include "common.cuh"

namespace scb {

int check_cube_args(const float *cube, int64_t nchan, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y);

struct MaskOpParams {
    const float *cube;
    int64_t nchan, ny, nx, stride_c, stride_y;
    int64_t row0, nrows;        // row window (pack); full image otherwise
    float fill;
    uint8_t *out_u8;
    float *out_f32;
    DevMask mask;
};

// OP 0: include -> uint8 ; OP 1: filled float32
template <int MODE, int OP>
__global__ void __launch_bounds__(256)
maskop_kernel(const __grid_constant__ MaskOpParams p) {
    const int64_t per_chan = p.nrows * p.nx;
    const int64_t total = p.nchan * per_chan;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = i / per_chan;
        const int64_t r = i - c * per_chan;
        const int64_t yl = r / p.nx;
        const int64_t x = r - yl * p.nx;
        const int64_t y = p.row0 + yl;
        const float v = __ldg(p.cube + c * p.stride_c + y * p.stride_y + x);
        const bool inc = mask_include<MODE>(p.mask, v, c, y, x);
        if (OP == 0) p.out_u8[i] = inc ? 1 : 0;
        else         p.out_f32[i] = inc ? v : p.fill;
    }
}

template <int OP>
static int launch_maskop(MaskOpParams &p, const sc_mask_desc *mask, cudaStream_t s) {
    int rc = build_dev_mask(mask, p.cube, p.stride_c, p.stride_y, &p.mask);
    if (rc) return rc;
    const int64_t total = p.nchan * p.nrows * p.nx;
    int64_t nb = cdiv(total, 256); if (nb > 148 * 32) nb = 148 * 32; unsigned grid = (unsigned)nb;
    LaunchScope ls(0, s);
    if (p.mask.mode == MODE_NONE)          maskop_kernel<MODE_NONE, OP><<<grid, 256, 0, s>>>(p);
    else if (p.mask.mode == MODE_INTERVAL) maskop_kernel<MODE_INTERVAL, OP><<<grid, 256, 0, s>>>(p);
    else                                   maskop_kernel<MODE_GENERIC, OP><<<grid, 256, 0, s>>>(p);
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}

}  // namespace scb

using namespace scb;

extern "C" int sc_mask_include(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                               int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask,
                               uint8_t *out, void *stream) {
    int rc = check_cube_args(cube, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(out != nullptr, "out is NULL");
    MaskOpParams p{};
    p.cube = cube; p.nchan = nchan; p.ny = ny; p.nx = nx; p.stride_c = stride_c; p.stride_y = stride_y;
    p.row0 = 0; p.nrows = ny; p.out_u8 = out;
    return launch_maskop<0>(p, mask, (cudaStream_t)stream);
}

extern "C" int sc_fill_masked(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                              int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask, double fill,
                              float *out, void *stream) {
    int rc = check_cube_args(cube, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(out != nullptr, "out is NULL");
    MaskOpParams p{};
    p.cube = cube; p.nchan = nchan; p.ny = ny; p.nx = nx; p.stride_c = stride_c; p.stride_y = stride_y;
    p.row0 = 0; p.nrows = ny; p.fill = (float)fill; p.out_f32 = out;
    return launch_maskop<1>(p, mask, (cudaStream_t)stream);
}

extern "C" int sc_pack_filled_rows(const float *in, int64_t nchan, int64_t ny, int64_t nx,
                                   int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask, double fill,
                                   int64_t row0, int64_t nrows, float *out, void *stream) {
    int rc = check_cube_args(in, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(out != nullptr, "out is NULL");
    SC_CHECK_ARG(row0 >= 0 && nrows > 0 && row0 + nrows <= ny, "rows [%lld, %lld) outside the image", (long long)row0, (long long)(row0 + nrows));
    MaskOpParams p{};
    p.cube = in; p.nchan = nchan; p.ny = ny; p.nx = nx; p.stride_c = stride_c; p.stride_y = stride_y;
    p.row0 = row0; p.nrows = nrows; p.fill = (float)fill; p.out_f32 = out;
    return launch_maskop<1>(p, mask, (cudaStream_t)stream);
}
