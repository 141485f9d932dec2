// Which 3-D box shapes does cp.async.bulk.tensor.3d accept on this part?  (scratch)
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, int bytes, int x, int y, int c, float *out) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(smem_u32(raw)), "l"(&tm), "r"(x), "r"(y), "r"(c), "r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" :: "r"(smem_u32(&bar)) : "memory");
        out[0] = reinterpret_cast<float *>(raw)[0] + reinterpret_cast<float *>(raw)[bytes / 4 - 1];
    }
}
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const int64_t nx = 4096, ny = 4096, nc = 8;
    float *src, *out; cudaMalloc(&src, nx * ny * nc * 4); cudaMalloc(&out, 4); cudaMemset(src, 0, nx * ny * nc * 4);
    EncodeFn encode; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &q);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    int shapes[][3] = {{256, 1, 16}, {64, 2, 16}, {32, 32, 2}, {48, 48, 2}, {48, 48, 1}, {64, 48, 2}, {48, 32, 2}, {64, 64, 2}, {48, 2, 2}, {40, 40, 2}};
    for (auto &sh : shapes) {
        cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nc};
        cuuint64_t strides[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * ny * 4};
        cuuint32_t box[3] = {(cuuint32_t)sh[0], (cuuint32_t)sh[1], (cuuint32_t)sh[2]};
        cuuint32_t es[3] = {1, 1, 1};
        CUtensorMap tm;
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        int bytes = sh[0] * sh[1] * sh[2] * 4;
        for (int x : {0, 5}) {
            k<<<1, 32, bytes>>>(tm, bytes, x, 3, 1, out);
            cudaError_t e = cudaDeviceSynchronize();
            printf("box {%d,%d,%d} x=%d encode=%d run=%s\n", sh[0], sh[1], sh[2], x, (int)r, cudaGetErrorString(e));
            if (e != cudaSuccess) { printf("(context lost; stopping)\n"); return 0; }
        }
    }
    return 0;
}
