// ingest.cu -- FITS pixel decoding on the device: big-endian BITPIX -32 / -64 / 16 / 32 samples to
// native float32 with BSCALE / BZERO and BLANK -> NaN applied.
//
// Replaces what `astropy.io.fits` does on the CPU between `load_fits_cube` (io/fits.py:171-260) and the
// first kernel: the byte swap of the big-endian data block (plus scaling for integer BITPIX), today the
// dominant wall-clock cost for real files.  The host streams the raw file bytes
// through pinned staging buffers (spectral_cube_b200/io_fits.py); this kernel turns each staged block
// into float32 voxels at HBM speed: 4 (or 2 / 8) B in, 4 B out per voxel, 16-byte vector accesses.
#include "common.cuh"

namespace scb {

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// BITPIX -32: four samples per thread
__global__ void __launch_bounds__(256)
decode_be_f32_kernel(const uint4 *__restrict__ in, float4 *__restrict__ out, int64_t n4, double bscale, double bzero, int scaled) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const uint4 r = in[i];
    float4 v = make_float4(__uint_as_float(bswap32(r.x)), __uint_as_float(bswap32(r.y)),
                           __uint_as_float(bswap32(r.z)), __uint_as_float(bswap32(r.w)));
    if (scaled) {
        v.x = (float)fma((double)v.x, bscale, bzero); v.y = (float)fma((double)v.y, bscale, bzero);
        v.z = (float)fma((double)v.z, bscale, bzero); v.w = (float)fma((double)v.w, bscale, bzero);
    }
    out[i] = v;
}

// every other case, one sample per thread (also the unaligned tail of BITPIX -32)
__global__ void __launch_bounds__(256)
decode_be_generic_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, int64_t n, int bitpix,
                         double bscale, double bzero, int has_blank, long long blank) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v;
    bool isblank = false;
    if (bitpix == -32) {
        const uint8_t *b = in + 4 * i;
        v = (double)__uint_as_float(((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]);
    } else if (bitpix == -64) {
        const uint8_t *b = in + 8 * i;
        unsigned long long u = 0;
        for (int k = 0; k < 8; ++k) u = (u << 8) | b[k];
        v = __longlong_as_double((long long)u);
    } else if (bitpix == 16) {
        const uint8_t *b = in + 2 * i;
        const short s = (short)(((uint16_t)b[0] << 8) | b[1]);
        isblank = has_blank && (long long)s == blank;
        v = (double)s;
    } else if (bitpix == 32) {
        const uint8_t *b = in + 4 * i;
        const int s = (int)(((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]);
        isblank = has_blank && (long long)s == blank;
        v = (double)s;
    } else {                                                     // BITPIX 8: unsigned bytes
        const unsigned char s = in[i];
        isblank = has_blank && (long long)s == blank;
        v = (double)s;
    }
    out[i] = isblank ? nan32() : (float)fma(v, bscale, bzero);
}

}  // namespace scb

using namespace scb;

extern "C" int sc_fits_decode(const void *raw_be, float *out, int64_t n, int bitpix,
                              double bscale, double bzero, int has_blank, long long blank, void *stream) {
    SC_CHECK_ARG(raw_be != nullptr && out != nullptr, "NULL buffer");
    SC_CHECK_ARG(n >= 0, "negative sample count");
    SC_CHECK_ARG(bitpix == -32 || bitpix == -64 || bitpix == 16 || bitpix == 32 || bitpix == 8, "unsupported BITPIX %d", bitpix);
    if (n == 0) return SC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    LaunchScope ls(SC_OP_INGEST, s);
    const bool scaled = !(bscale == 1.0 && bzero == 0.0);
    if (bitpix == -32 && (uintptr_t)raw_be % 16 == 0 && (uintptr_t)out % 16 == 0) {
        const int64_t n4 = n / 4;
        if (n4 > 0) decode_be_f32_kernel<<<(unsigned)cdiv(n4, 256), 256, 0, s>>>((const uint4 *)raw_be, (float4 *)out, n4, bscale, bzero, scaled ? 1 : 0);
        const int64_t rest = n - 4 * n4;
        if (rest > 0) decode_be_generic_kernel<<<1, 256, 0, s>>>((const uint8_t *)raw_be + 16 * n4, out + 4 * n4, rest, bitpix, bscale, bzero, 0, 0);
    } else {
        decode_be_generic_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>((const uint8_t *)raw_be, out, n, bitpix, bscale, bzero, has_blank, blank);
    }
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}
