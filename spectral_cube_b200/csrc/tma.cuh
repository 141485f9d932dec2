// tma.cuh -- sm_100a bulk-async-copy (TMA) + mbarrier primitives used to stage spectral slabs
// of the cube in shared memory.  One elected producer lane issues `cp.async.bulk` (SASS:
// UBLKCP) row copies that complete on an mbarrier; consumer warps wait on the barrier's
// phase parity and hand the stage back through a second ("empty") barrier.
#pragma once
#include <stdint.h>
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

namespace scb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// L2 policy for read-once streams
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// global -> shared bulk copy of `bytes` (multiple of 16; both addresses 16-byte aligned),
// completion counted in bytes on `bar`.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                            uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_load_1d_nohint(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                                   uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// global -> shared TILED copy of one box {box_x, 1, box_c} of a (c, y, x) float32 cube described by a
// tensor map (SASS: UTMALDG): ONE instruction per ring slot whatever the number of rows; elements
// outside the cube arrive as zeros and still count towards the barrier's byte total.
__device__ __forceinline__ void tma_load_box3d(void *smem_dst, const CUtensorMap *tmap, int x, int y, int c,
                                               uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3, %4}], [%5], %6;"
        :: "r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(c), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// Host: describe a float32 cube view (element strides, x contiguous) and the box one request moves.
// Needs base % 16 == 0, stride_y % 4 == 0, stride_c % 4 == 0, box_x % 4 == 0, box_x, box_c <= 256.
int make_cube_tensor_map(CUtensorMap *out, const float *base, int64_t nchan, int64_t ny, int64_t nx,
                         int64_t stride_c, int64_t stride_y, int box_x, int box_c);
// the same with a box {box_x, box_y, box_c}
int make_cube_tensor_map3(CUtensorMap *out, const float *base, int64_t nchan, int64_t ny, int64_t nx,
                          int64_t stride_c, int64_t stride_y, int box_x, int box_y, int box_c);

// shared -> global bulk store (bulk_group completion)
__device__ __forceinline__ void tma_store_1d(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory");
}
// make generic-proxy smem writes visible to the async proxy before a bulk store
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace scb
