// common.cuh -- shared pieces of libsc_b200: error slot, launch accounting, the device-side
// mask predicate (restating spectral_cube/masks.py), streaming load/store helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>
#include <float.h>
#include "../../include/sc_b200.h"

namespace scb {

// ---- error slot (thread local) ---------------------------------------------------------
void set_error(const char *fmt, ...);
int  cuda_fail(cudaError_t e, const char *what);
void count_launch(int op, cudaStream_t s, bool begin);   // timing hooks + launch counter

#define SC_CHECK_ARG(cond, ...) do { if (!(cond)) { scb::set_error(__VA_ARGS__); return SC_ERR_ARG; } } while (0)
#define SC_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return scb::cuda_fail(e_, #call); } while (0)

// RAII-less helper: bracket a launch with optional timing events and bump the counter.
struct LaunchScope {
    int op; cudaStream_t s;
    LaunchScope(int op_, cudaStream_t s_) : op(op_), s(s_) { count_launch(op, s, true); }
    ~LaunchScope() { count_launch(op, s, false); }
};

// cudaFuncSetAttribute applies to the CURRENT device only: `done` keeps one bit per device ordinal, so a process
// that drives several GPUs configures every kernel once on each of them (a race sets the attribute twice: benign).
template <typename K>
static inline cudaError_t ensure_dyn_smem(K kern, size_t smem, unsigned long long *done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (*done & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) *done |= bit;
    return e;
}

// ---- device mask program ------------------------------------------------------------------
// MODE_NONE:     mask is None                      -> include = !isnan(v)   (nansum semantics)
// MODE_INTERVAL: isfinite(self) [& self OP scalar] -> include = lo < v && v < hi  (float32
//                bounds chosen on the host so the test equals (double)v OP thr exactly)
// MODE_GENERIC:  anything else: interpret the node list.
// MODE_INTERVAL_OTHER: the same interval test on ANOTHER cube's voxel (the mask a smoothed cube keeps:
//                built on the data it was smoothed from, spectral_cube.py:3043-3045).  Kernels without a
//                dedicated path treat it as MODE_GENERIC (the node program is always carried).
enum { MODE_NONE = 0, MODE_INTERVAL = 1, MODE_GENERIC = 2, MODE_INTERVAL_OTHER = 3 };

struct DevMask {
    int   mode;
    float lo, hi;
    const float *other;    // MODE_INTERVAL_OTHER: the cube the interval is tested on, and its strides
    int64_t other_sc, other_sy;
    sc_mask_desc prog;     // only read in MODE_GENERIC
};

// Host: classify a descriptor.  Returns SC_OK or an error.
// `allow_other`: the caller has a MODE_INTERVAL_OTHER path (only the moment kernels do); everyone else
// gets MODE_GENERIC for masks that live on another cube.
int build_dev_mask(const sc_mask_desc *m, const float *cube, int64_t stride_c, int64_t stride_y,
                   DevMask *out, bool allow_other = false);
// Host: true if any node references memory other than the cube itself.
bool mask_is_self_only(const sc_mask_desc *m);

__device__ __forceinline__ bool cmp_f64(int op, double x, double t) {
    switch (op) {
        case SC_GT: return x >  t;
        case SC_GE: return x >= t;
        case SC_LT: return x <  t;
        case SC_LE: return x <= t;
        case SC_EQ: return x == t;
        default:    return x != t;
    }
}

// Evaluate the generic mask program for voxel (c, y, x); `self_v` is the cube value there.
__device__ __forceinline__ bool eval_mask_generic(const sc_mask_desc &m, float self_v,
                                                  int64_t c, int64_t y, int64_t x) {
    uint32_t bits = 0;
    const int n = m.n_nodes;
    for (int i = 0; i < n; ++i) {
        const sc_mask_node &nd = m.nodes[i];
        bool r = false;
        switch (nd.kind) {
            case SC_MASK_FINITE: {
                float v = nd.data ? __ldg(nd.data + c * nd.ds_c + y * nd.ds_y + x) : self_v;
                r = fabsf(v) <= FLT_MAX;
                break;
            }
            case SC_MASK_CMP_SCALAR: {
                float v = nd.data ? __ldg(nd.data + c * nd.ds_c + y * nd.ds_y + x) : self_v;
                r = cmp_f64(nd.op, (double)v, nd.value);
                break;
            }
            case SC_MASK_CMP_ARRAY: {
                float v = nd.data ? __ldg(nd.data + c * nd.ds_c + y * nd.ds_y + x) : self_v;
                int64_t off = c * nd.as_c + y * nd.as_y + x * nd.as_x;
                double t = nd.array_dtype == SC_F64 ? __ldg((const double *)nd.array + off)
                                                    : (double)__ldg((const float *)nd.array + off);
                r = cmp_f64(nd.op, (double)v, t);
                break;
            }
            case SC_MASK_BOOL: {
                int64_t off = c * nd.as_c + y * nd.as_y + x * nd.as_x;
                r = __ldg((const uint8_t *)nd.array + off) != 0;
                break;
            }
            case SC_MASK_AND: r = ((bits >> nd.a) & 1u) & ((bits >> nd.b) & 1u); break;
            case SC_MASK_OR:  r = ((bits >> nd.a) & 1u) | ((bits >> nd.b) & 1u); break;
            case SC_MASK_XOR: r = ((bits >> nd.a) & 1u) ^ ((bits >> nd.b) & 1u); break;
            case SC_MASK_NOT: r = !((bits >> nd.a) & 1u); break;
            default: break;
        }
        bits |= (r ? 1u : 0u) << i;
    }
    return n == 0 ? true : ((bits >> (n - 1)) & 1u) != 0;
}

// include predicate of the MASK alone (no NaN test), for the three modes
template <int MODE>
__device__ __forceinline__ bool mask_include(const DevMask &m, float v, int64_t c, int64_t y, int64_t x) {
    if (MODE == MODE_NONE) return true;
    if (MODE == MODE_INTERVAL) return (v > m.lo) & (v < m.hi);
    return eval_mask_generic(m.prog, v, c, y, x);
}

// ---- streaming loads: read-once data bypasses L1 allocation ---------------------------------
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ldg_stream2(const float *p) {
    float2 r;
    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];"
                 : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float *p) {
    float r;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream4(float *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream1(float *p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

template <int VEC> struct VecLoad;
template <> struct VecLoad<4> {
    __device__ __forceinline__ static void load(const float *p, float (&v)[4]) {
        float4 t = ldg_stream4(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};
template <> struct VecLoad<2> {
    __device__ __forceinline__ static void load(const float *p, float (&v)[2]) {
        float2 t = ldg_stream2(p); v[0] = t.x; v[1] = t.y;
    }
};
template <> struct VecLoad<1> {
    __device__ __forceinline__ static void load(const float *p, float (&v)[1]) { v[0] = ldg_stream1(p); }
};

__device__ __forceinline__ double nan64() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ float  nan32() { return __int_as_float(0x7fc00000); }

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace scb
