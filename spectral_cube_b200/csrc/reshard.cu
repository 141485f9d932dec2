// reshard.cu -- rows -> channels re-shard of a row-sharded cube over NVLink peer memory (the one exchange
// `reproject` needs on a cube that is partitioned over the spatial plane).
//
// Every rank holds rows [y0, y0 + rows) of all channels; afterwards rank d holds ALL rows of channels [cb[d], cb[d+1]).
// The NCCL formulation (all_to_all into per-source blocks + a concatenation pass) moves every byte three times on the
// receiving side.  Here ONE kernel per rank reads its local block once (16-byte streaming loads) and stores every
// channel straight into its final place in the destination rank's (chans_d, ny_total, nx) buffer through peer-mapped
// pointers (st.global over NVLink 5 / NVSwitch; 512 contiguous bytes per warp and request): no staging, no second pass.
// The buffers are symmetric allocations the host layer exchanges once (`torch.distributed._symmetric_memory`); barriers
// before and after the kernel are the caller's (signal pads on the same stream).
#include "common.cuh"

namespace scb {

constexpr int RS_MAX_RANKS = 16;

struct ReshardParams {
    const float *local;
    int64_t nchan, rows, nx4, stride_c, stride_y;      // nx4 = nx / 4
    int64_t ny_total, y0;
    int world, rank;
    int64_t chans_max;                                  // largest channel block of the partition
    float *peer[RS_MAX_RANKS];                          // rank d's (chans_d, ny_total, nx) buffer, mapped here
    int64_t cb[RS_MAX_RANKS + 1];                       // channel bounds of the destination partition
};

__global__ void __launch_bounds__(256)
reshard_scatter_kernel(const __grid_constant__ ReshardParams p) {
    // one (channel, row) line per loop trip of a CTA: the destination and both base pointers are uniform, threads move
    // the line's float4 elements.  Consecutive lines go to DIFFERENT destinations, starting with this rank's right-hand
    // neighbour: at any moment every rank stores to all peers and every peer receives from all ranks at an equal share
    // (channel-major order made all ranks store into rank 0 first, then rank 1 ...: 238 GB/s per direction, incast-bound).
    const int64_t items = (int64_t)p.world * p.chans_max * p.rows;
    for (int64_t w = blockIdx.x; w < items; w += gridDim.x) {
        const int d = (int)((w + p.rank + 1) % p.world);
        const int64_t k = w / p.world;
        const int64_t cl = k / p.rows, r = k - cl * p.rows;
        const int64_t c = p.cb[d] + cl;
        if (c >= p.cb[d + 1]) continue;                 // (unequal partition: this destination's block is shorter)
        const float4 *src = reinterpret_cast<const float4 *>(p.local + c * p.stride_c + r * p.stride_y);
        float4 *dst = reinterpret_cast<float4 *>(p.peer[d] + ((c - p.cb[d]) * p.ny_total + p.y0 + r) * (p.nx4 * 4));
        // four independent 16-byte loads in flight per thread before the first peer store
        for (int64_t i0 = threadIdx.x; i0 < p.nx4; i0 += 4 * (int64_t)blockDim.x) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + u * (int64_t)blockDim.x;
                if (i < p.nx4) v[u] = ldg_stream4(reinterpret_cast<const float *>(src + i));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + u * (int64_t)blockDim.x;
                if (i < p.nx4) dst[i] = v[u];
            }
        }
    }
}

}  // namespace scb

using namespace scb;

extern "C" int sc_reshard_scatter(const float *local, int64_t nchan, int64_t rows, int64_t nx,
                                  int64_t stride_c, int64_t stride_y,
                                  const uint64_t *peer_ptrs, int world, int rank, const int64_t *chan_bounds,
                                  int64_t ny_total, int64_t y0, void *stream) {
    SC_CHECK_ARG(local != nullptr && peer_ptrs != nullptr && chan_bounds != nullptr, "NULL argument");
    SC_CHECK_ARG(nchan > 0 && rows > 0 && nx > 0, "bad shape");
    SC_CHECK_ARG(world >= 1 && world <= RS_MAX_RANKS, "world size %d (1 .. %d ranks)", world, RS_MAX_RANKS);
    SC_CHECK_ARG(rank >= 0 && rank < world, "rank %d of %d", rank, world);
    SC_CHECK_ARG(nx % 4 == 0 && stride_c % 4 == 0 && stride_y % 4 == 0 && (uintptr_t)local % 16 == 0,
                 "the peer re-shard moves 16-byte vectors: nx, the strides and the base address must be multiples of 4 floats");
    SC_CHECK_ARG(y0 >= 0 && y0 + rows <= ny_total, "rows [%lld, %lld) outside the %lld rows of the image",
                 (long long)y0, (long long)(y0 + rows), (long long)ny_total);
    SC_CHECK_ARG(chan_bounds[0] == 0 && chan_bounds[world] == nchan, "channel bounds must run from 0 to nchan");
    ReshardParams p{};
    p.local = local; p.nchan = nchan; p.rows = rows; p.nx4 = nx / 4; p.stride_c = stride_c; p.stride_y = stride_y;
    p.ny_total = ny_total; p.y0 = y0; p.world = world; p.rank = rank; p.chans_max = 0;
    for (int d = 0; d < world; ++d) {
        SC_CHECK_ARG(chan_bounds[d + 1] >= chan_bounds[d], "channel bounds must ascend");
        SC_CHECK_ARG(peer_ptrs[d] != 0 && peer_ptrs[d] % 16 == 0, "peer buffer %d is NULL or not 16-byte aligned", d);
        p.peer[d] = reinterpret_cast<float *>(peer_ptrs[d]);
        p.cb[d] = chan_bounds[d];
        if (chan_bounds[d + 1] - chan_bounds[d] > p.chans_max) p.chans_max = chan_bounds[d + 1] - chan_bounds[d];
    }
    p.cb[world] = chan_bounds[world];
    cudaStream_t s = (cudaStream_t)stream;
    LaunchScope ls(SC_OP_POINTWISE, s);
    const int64_t items = (int64_t)world * p.chans_max * rows;
    const int64_t cap = 148 * 8;
    reshard_scatter_kernel<<<(unsigned)(items < cap ? items : cap), 256, 0, s>>>(p);
    SC_CUDA(cudaGetLastError());
    return SC_OK;
}
