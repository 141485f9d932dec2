// host_pipeline.cu -- moment maps of a cube that lives in HOST memory.
//
// The reference reads its cube from a numpy array / memmap (spectral_cube.py:1614 ->
// _moments.py); this entry point is what a drop-in binding calls with that host buffer.
// Row blocks (all channels x R rows) are streamed through two device staging buffers on two
// streams, so the PCIe copy of block i+1 overlaps the reduction of block i; the float64 maps
// are assembled on the device and copied back once.
#include "common.cuh"
#include <vector>

namespace scb {
int check_cube_args(const float *cube, int64_t nchan, int64_t ny, int64_t nx, int64_t stride_c, int64_t stride_y);
int moments_axis0_device(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                         int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask,
                         const double *chan_offset_dev, double K,
                         double pix_size, double m1_offset, int want_bits,
                         double *out_m0, double *out_m1, double *out_m2,
                         double2 *tab_dev, cudaStream_t s);

struct HostPipe {
    float *stage[2] = {nullptr, nullptr};
    double *maps = nullptr;
    char *ws = nullptr;
    cudaStream_t st[2] = {nullptr, nullptr};
    ~HostPipe() {
        for (int i = 0; i < 2; ++i) { if (st[i]) cudaStreamDestroy(st[i]); if (stage[i]) cudaFree(stage[i]); }
        if (maps) cudaFree(maps);
        if (ws) cudaFree(ws);
    }
};
}  // namespace scb

using namespace scb;

extern "C" int sc_moments_axis0_host(const float *cube_host, int64_t nchan, int64_t ny, int64_t nx,
                                     int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask,
                                     const double *chan_offset, double pix_size, double m1_offset,
                                     int want_bits, double *out_m0_host, double *out_m1_host,
                                     double *out_m2_host, size_t staging_bytes, int device) {
    int rc = check_cube_args(cube_host, nchan, ny, nx, stride_c, stride_y);
    if (rc) return rc;
    SC_CHECK_ARG(want_bits > 0 && want_bits <= 7, "want_bits=%d must be a combination of 1|2|4", want_bits);
    SC_CHECK_ARG(!(want_bits & SC_WANT_M0) || out_m0_host, "out_m0 is NULL but moment 0 was requested");
    SC_CHECK_ARG(!(want_bits & SC_WANT_M1) || out_m1_host, "out_m1 is NULL but moment 1 was requested");
    SC_CHECK_ARG(!(want_bits & SC_WANT_M2) || out_m2_host, "out_m2 is NULL but moment 2 was requested");
    SC_CHECK_ARG(mask_is_self_only(mask), "the host pipeline accepts only masks that refer to the cube itself");
    SC_CHECK_ARG(stride_y == nx || stride_c % stride_y == 0, "host cube: stride_c must be a multiple of stride_y for strided rows");
    if (want_bits & (SC_WANT_M1 | SC_WANT_M2)) SC_CHECK_ARG(chan_offset != nullptr, "chan_offset is NULL");
    SC_CUDA(cudaSetDevice(device));

    if (staging_bytes == 0) staging_bytes = (size_t)512 << 20;
    // rows per block: multiple of 1 row, nx padded to a multiple of 4 floats so the TMA kernel applies
    const int64_t nxp = (nx + 3) / 4 * 4;
    const size_t row_bytes = (size_t)nchan * nxp * sizeof(float);
    int64_t R = (int64_t)((staging_bytes / 2) / row_bytes);
    if (R < 1) R = 1;
    if (R > ny) R = ny;
    const bool padded = nxp != nx;

    HostPipe hp;
    for (int i = 0; i < 2; ++i) {
        SC_CUDA(cudaStreamCreateWithFlags(&hp.st[i], cudaStreamNonBlocking));
        SC_CUDA(cudaMalloc(&hp.stage[i], (size_t)R * row_bytes));
    }
    const size_t plane = (size_t)ny * nx;
    SC_CUDA(cudaMalloc(&hp.maps, 3 * plane * sizeof(double)));
    const size_t ws_each = (size_t)nchan * 24 + 512;
    SC_CUDA(cudaMalloc(&hp.ws, 2 * ws_each + 512));
    double *m0 = hp.maps, *m1 = hp.maps + plane, *m2 = hp.maps + 2 * plane;

    double K = 0.0;
    double2 *tab[2] = {nullptr, nullptr};
    double *xdev[2] = {nullptr, nullptr};
    if (want_bits & (SC_WANT_M1 | SC_WANT_M2)) {
        K = chan_offset[nchan / 2];
        for (int i = 0; i < 2; ++i) {
            tab[i] = (double2 *)(((uintptr_t)hp.ws + (size_t)i * ws_each + 255) & ~(uintptr_t)255);
            xdev[i] = (double *)(tab[i] + nchan);
            SC_CUDA(cudaMemcpyAsync(xdev[i], chan_offset, (size_t)nchan * 8, cudaMemcpyHostToDevice, hp.st[i]));
        }
    }

    int blk = 0;
    for (int64_t y0 = 0; y0 < ny; y0 += R, ++blk) {
        const int b = blk & 1;
        const int64_t r = (y0 + R <= ny) ? R : ny - y0;
        cudaStream_t s = hp.st[b];
        const float *src = cube_host + y0 * stride_y;
        if (padded) SC_CUDA(cudaMemsetAsync(hp.stage[b], 0xFF, (size_t)r * row_bytes, s));   // NaN padding
        if (stride_y == nx && !padded) {
            // a row block of one channel is contiguous: one 2-D copy (height = channels)
            SC_CUDA(cudaMemcpy2DAsync(hp.stage[b], (size_t)r * nx * 4, src, (size_t)stride_c * 4,
                                      (size_t)r * nx * 4, (size_t)nchan, cudaMemcpyHostToDevice, s));
        } else {
            cudaMemcpy3DParms cp = {};
            cp.srcPtr = make_cudaPitchedPtr((void *)src, (size_t)stride_y * 4, (size_t)nx, (size_t)(stride_c / stride_y));
            cp.dstPtr = make_cudaPitchedPtr((void *)hp.stage[b], (size_t)nxp * 4, (size_t)nx, (size_t)r);
            cp.extent = make_cudaExtent((size_t)nx * 4, (size_t)r, (size_t)nchan);
            cp.kind = cudaMemcpyHostToDevice;
            SC_CUDA(cudaMemcpy3DAsync(&cp, s));
        }
        if (!padded) {
            rc = moments_axis0_device(hp.stage[b], nchan, r, nx, r * nx, nx, mask, xdev[b], K, pix_size, m1_offset,
                                      want_bits, m0 + y0 * nx, m1 + y0 * nx, m2 + y0 * nx, tab[b], s);
        } else {
            // padded rows: outputs keep the true nx, so run on the staged block row by row layout
            rc = moments_axis0_device(hp.stage[b], nchan, r, nx, r * nxp, nxp, mask, xdev[b], K, pix_size, m1_offset,
                                      want_bits, m0 + y0 * nx, m1 + y0 * nx, m2 + y0 * nx, tab[b], s);
        }
        if (rc) return rc;
    }
    for (int i = 0; i < 2; ++i) SC_CUDA(cudaStreamSynchronize(hp.st[i]));
    if (want_bits & SC_WANT_M0) SC_CUDA(cudaMemcpy(out_m0_host, m0, plane * 8, cudaMemcpyDeviceToHost));
    if (want_bits & SC_WANT_M1) SC_CUDA(cudaMemcpy(out_m1_host, m1, plane * 8, cudaMemcpyDeviceToHost));
    if (want_bits & SC_WANT_M2) SC_CUDA(cudaMemcpy(out_m2_host, m2, plane * 8, cudaMemcpyDeviceToHost));
    return SC_OK;
}
