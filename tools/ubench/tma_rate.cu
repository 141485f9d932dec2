// Micro-benchmark: global->shared streaming rate of a TMA ring on sm_100a as a function of the
// request shape -- R separate 1-D bulk copies of `rowbytes` per stage (cp.async.bulk) against ONE
// 3-D tensor-map box of the same bytes (cp.async.bulk.tensor.3d).  Rows of a stage are a plane
// stride apart, like the channels of a cube.  Scratch tool for DESIGN.md.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t par) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" :: "r"(smem_u32(b)), "r"(par) : "memory");
}

constexpr int STAGE_BYTES = 16384;
constexpr int STAGES = 6;
constexpr int CONSUMERS = 128;

struct Args {
    const float *src; int64_t plane;      // floats per plane
    int rowfloats, rows;                  // per stage
    int tiles_per_plane_row; int64_t nx, ny; int nblk; float *out;
};

template <bool TENSOR>
__global__ void __launch_bounds__(CONSUMERS + 32) ring_kernel(const __grid_constant__ Args a, const __grid_constant__ CUtensorMap tm) {
    extern __shared__ __align__(128) unsigned char raw[];
    float *data = reinterpret_cast<float *>(raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(raw + STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONSUMERS / 32); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int64_t tile = blockIdx.x;
    const int64_t y = tile / a.tiles_per_plane_row, x0 = (tile % a.tiles_per_plane_row) * a.rowfloats;
    if (warp == CONSUMERS / 32) {
        for (int j = 0; j < a.nblk; ++j) {
            const int s = j % STAGES;
            if (j >= STAGES) mbar_wait(&empty[s], ((j / STAGES) - 1) & 1);
            if (TENSOR) {
                if (lane == 0) {
                    mbar_expect_tx(&full[s], STAGE_BYTES);
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 :: "r"(smem_u32(data + s * (STAGE_BYTES / 4))), "l"(&tm), "r"((int)x0), "r"((int)y), "r"(j * a.rows), "r"(smem_u32(&full[s])) : "memory");
                }
            } else {
                if (lane == 0) mbar_expect_tx(&full[s], STAGE_BYTES);
                __syncwarp();
                for (int r = lane; r < a.rows; r += 32) {
                    const float *g = a.src + ((int64_t)j * a.rows + r) * a.plane + y * a.nx + x0;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(smem_u32(data + s * (STAGE_BYTES / 4) + r * a.rowfloats)), "l"(g), "r"(a.rowfloats * 4), "r"(smem_u32(&full[s])) : "memory");
                }
            }
        }
    } else {
        float acc = 0.f;
        for (int j = 0; j < a.nblk; ++j) {
            const int s = j % STAGES;
            mbar_wait(&full[s], (j / STAGES) & 1);
            acc += data[s * (STAGE_BYTES / 4) + tid];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 1.2345f) a.out[0] = acc;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int64_t nx = 2048, ny = 2048; const int nchan = 512;
    const int64_t plane = nx * ny;
    float *src, *out;
    cudaMalloc(&src, plane * nchan * 4); cudaMalloc(&out, 4);
    cudaMemset(src, 0, plane * nchan * 4);
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres);
    const size_t smem = STAGES * STAGE_BYTES + 2 * STAGES * 8;
    cudaFuncSetAttribute(ring_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(ring_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int tensor = 0; tensor < 2; ++tensor) {
        for (int rowbytes = 256; rowbytes <= (tensor ? 1024 : 8192); rowbytes *= 2) {
            Args a; a.src = src; a.plane = plane; a.rowfloats = rowbytes / 4; a.rows = STAGE_BYTES / rowbytes;
            a.tiles_per_plane_row = (int)(nx / a.rowfloats); a.nx = nx; a.ny = ny; a.nblk = nchan / a.rows; a.out = out;
            if (a.rows > nchan) continue;
            CUtensorMap tm; memset(&tm, 0, sizeof(tm));
            if (tensor) {
                cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nchan};
                cuuint64_t strides[2] = {(cuuint64_t)nx * 4, (cuuint64_t)plane * 4};
                cuuint32_t box[3] = {(cuuint32_t)a.rowfloats, 1, (cuuint32_t)a.rows};
                cuuint32_t es[3] = {1, 1, 1};
                if (a.rows > 256) continue;
                CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            }
            const int grid = (int)(a.tiles_per_plane_row * ny);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (tensor) ring_kernel<true><<<grid, CONSUMERS + 32, smem>>>(a, tm); else ring_kernel<false><<<grid, CONSUMERS + 32, smem>>>(a, tm);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            cudaError_t err = cudaGetLastError();
            printf("%s row %5d B x %3d rows/stage  grid %6d : %7.3f ms  %7.1f GB/s  (%s)\n", tensor ? "tensor3d" : "bulk1d  ", rowbytes, a.rows, grid, ms,
                   (double)plane * nchan * 4 / ms / 1e6, cudaGetErrorString(err));
        }
    }
    return 0;
}
