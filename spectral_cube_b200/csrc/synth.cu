// synth.cu -- reproducible synthetic Gaussian-line cubes generated on the device.
//
// Mirrors the reference's test generator (spectral_cube/tests/utilities.py:53-112: one
// Gaussian line per spaxel, sigma = 8 channels, plus unit-variance noise) but draws every
// number from a counter-based integer hash of (seed, global voxel index) and uses only
// integer arithmetic, one table lookup and two correctly-rounded float32 operations, so
// oracle/synth.py regenerates any sub-block bit-identically with numpy.
#include "common.cuh"

namespace scb {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t k) {
    uint64_t z = k * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t sum16x4(uint64_t h) {
    return (uint32_t)(h & 0xFFFF) + (uint32_t)((h >> 16) & 0xFFFF) + (uint32_t)((h >> 32) & 0xFFFF) + (uint32_t)(h >> 48);
}

// stream offsets that separate the per-spaxel and per-voxel draws
#define SYN_STREAM_AMP   0x1000000000000000ULL
#define SYN_STREAM_CEN   0x2000000000000000ULL
#define SYN_STREAM_NAN   0x3000000000000000ULL
#define SYN_STREAM_N1    0x4000000000000000ULL
#define SYN_STREAM_N2    0x5000000000000000ULL

__global__ void __launch_bounds__(256)
synth_cube_kernel(float *cube, int64_t nchan, int64_t ny, int64_t nx,
                  int64_t y0, int64_t x0, int64_t ny_total, int64_t nx_total,
                  uint64_t seed, const float *__restrict__ profile, int nan_permille, int border) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ny * nx) return;
    const int64_t yl = g / nx, xl = g - yl * nx;
    const int64_t Y = y0 + yl, X = x0 + xl;
    const uint64_t sp = (uint64_t)(Y * nx_total + X);
    const bool in_border = Y < border || X < border || Y >= ny_total - border || X >= nx_total - border;

    const uint64_t ha = mix64(seed + SYN_STREAM_AMP + sp);
    const uint64_t hc = mix64(seed + SYN_STREAM_CEN + sp);
    const float amp = __fmul_rn((float)(uint32_t)(ha >> 40), 10.0f / 16777216.0f);
    const int64_t span = 8 * nchan;                                // 16 * nchan / 2 lattice points
    const int64_t c0_16 = 4 * nchan + (int64_t)(hc % (uint64_t)span);
    const float nscale = 1.0f / 53510.0f;                          // Irwin-Hall(8) of 16-bit ints -> unit variance

    float *out = cube + yl * nx + xl;
    const uint64_t plane = (uint64_t)(ny_total * nx_total);
    for (int64_t c = 0; c < nchan; ++c) {
        const uint64_t vi = (uint64_t)c * plane + sp;
        float val;
        bool isnan_v = in_border;
        if (!isnan_v && nan_permille > 0) {
            const uint64_t hn = mix64(seed + SYN_STREAM_NAN + vi);
            isnan_v = (((hn >> 32) * 1000ULL) >> 32) < (uint64_t)nan_permille;
        }
        if (isnan_v) {
            val = nan32();
        } else {
            const uint64_t h1 = mix64(seed + SYN_STREAM_N1 + vi);
            const uint64_t h2 = mix64(seed + SYN_STREAM_N2 + vi);
            const int32_t s = (int32_t)(sum16x4(h1) + sum16x4(h2)) - 262140;
            const float noise = __fmul_rn((float)s, nscale);
            int64_t k = 16 * c - c0_16;
            if (k < 0) k = -k;
            val = __fadd_rn(__fmul_rn(amp, __ldg(profile + k)), noise);
        }
        out[c * ny * nx] = val;
    }
}

}  // namespace scb

using namespace scb;

extern "C" int sc_synth_cube(float *cube, int64_t nchan, int64_t ny, int64_t nx,
                             int64_t y0, int64_t x0, int64_t ny_total, int64_t nx_total,
                             uint64_t seed, const float *profile, int nan_permille, int border,
                             void *stream) {
    SC_CHECK_ARG(cube && profile, "cube/profile must not be NULL");
    SC_CHECK_ARG(nchan > 0 && ny > 0 && nx > 0, "bad shape");
    SC_CHECK_ARG(y0 >= 0 && x0 >= 0 && y0 + ny <= ny_total && x0 + nx <= nx_total, "block does not fit the full plane");
    SC_CHECK_ARG(nan_permille >= 0 && nan_permille <= 1000 && border >= 0, "bad nan_permille/border");
    cudaStream_t s = (cudaStream_t)stream;
    LaunchScope ls(0, s);
    synth_cube_kernel<<<(unsigned)cdiv(ny * nx, 256), 256, 0, s>>>(cube, nchan, ny, nx, y0, x0, ny_total, nx_total,
                                                                    seed, profile, nan_permille, border);
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}
