// Micro-benchmark: DFMA issue rate on sm_100a as a function of warps/SM and independent chains
// per thread (register vs constant-bank multiplier).  Scratch tool for DESIGN.md's FP64 floor.
#include <cstdio>
#include <cuda_runtime.h>

struct P { double k[8]; };

template <int ILP, bool CONSTOP>
__global__ void dfma_kernel(const __grid_constant__ P p, double *out, int iters, double seed) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = seed + i + threadIdx.x;
    double r0 = seed * 1.0000001, r1 = seed * 0.9999999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (CONSTOP) a[i] = fma(a[i], p.k[u], p.k[(u + 1) & 7]);
                else         a[i] = fma(a[i], (u & 1) ? r0 : r1, r1);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP, bool CONSTOP>
void run(int threads, int ctas_per_sm, double *out) {
    P p; for (int i = 0; i < 8; ++i) p.k[i] = 1.0 + 1e-9 * i;
    int iters = 4096;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int grid = 148 * ctas_per_sm;
    dfma_kernel<ILP, CONSTOP><<<grid, threads>>>(p, out, 16, 1.0);
    cudaEventRecord(a);
    dfma_kernel<ILP, CONSTOP><<<grid, threads>>>(p, out, iters, 1.0);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double n = (double)grid * threads * iters * 8.0 * ILP;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("ILP %d const %d threads %4d ctas/SM %d warps/SM %2d : %.3f ms  %.2f TDFMA/s  %.1f lanes/clk/SM (at %d MHz)\n", ILP, (int)CONSTOP,
           threads, ctas_per_sm, threads * ctas_per_sm / 32, ms, n / ms / 1e9, n / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
}

int main() {
    double *out; cudaMalloc(&out, 8);
    run<1, false>(256, 2, out); run<4, false>(256, 2, out); run<8, false>(256, 2, out);
    run<4, true>(256, 2, out);  run<8, true>(256, 2, out);
    run<4, true>(128, 1, out);  run<4, true>(256, 1, out); run<4, true>(512, 2, out); run<4, true>(1024, 2, out);
    run<16, true>(256, 2, out); run<16, false>(256, 2, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
