// core.cu -- library plumbing: error slot, launch accounting / per-op timing, mask classification.
#include "common.cuh"
#include "tma.cuh"
#include <atomic>
#include <string.h>
#include <mutex>

namespace scb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static std::atomic<int> g_timing{0};

struct OpTimer { cudaEvent_t a = nullptr, b = nullptr; bool valid = false; };
static OpTimer g_timers[16];
static std::mutex g_timer_mu;

void set_error(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return SC_ERR_CUDA;
}

void count_launch(int op, cudaStream_t s, bool begin) {
    if (begin) g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_timing.load(std::memory_order_relaxed) || op <= 0 || op >= 16) return;
    std::lock_guard<std::mutex> lk(g_timer_mu);
    OpTimer &t = g_timers[op];
    if (!t.a) { cudaEventCreate(&t.a); cudaEventCreate(&t.b); }
    if (begin) { cudaEventRecord(t.a, s); t.valid = false; }
    else       { cudaEventRecord(t.b, s); t.valid = true; }
}

// ---- tensor maps (TMA descriptors) -------------------------------------------------------------
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_cube_tensor_map(CUtensorMap *out, const float *base, int64_t nchan, int64_t ny, int64_t nx,
                         int64_t stride_c, int64_t stride_y, int box_x, int box_c) {
    return make_cube_tensor_map3(out, base, nchan, ny, nx, stride_c, stride_y, box_x, 1, box_c);
}

int make_cube_tensor_map3(CUtensorMap *out, const float *base, int64_t nchan, int64_t ny, int64_t nx,
                          int64_t stride_c, int64_t stride_y, int box_x, int box_y, int box_c) {
    static std::atomic<TensorMapEncodeFn> cached{nullptr};
    TensorMapEncodeFn fn = cached.load(std::memory_order_acquire);
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
        if (e != cudaSuccess || !sym || q != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled is not available from the driver (cudaGetDriverEntryPoint: %d)", (int)e);
            return SC_ERR_CUDA;
        }
        fn = (TensorMapEncodeFn)sym;
        cached.store(fn, std::memory_order_release);
    }
    // a one-row cube still needs a non-zero plane stride for the descriptor
    const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nchan};
    const cuuint64_t strides[2] = {(cuuint64_t)stride_y * 4u, (cuuint64_t)stride_c * 4u};
    const cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, (cuuint32_t)box_c};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for a %lld x %lld x %lld cube, strides %lld / %lld, box %d x %d",
                  (int)r, (long long)nchan, (long long)ny, (long long)nx, (long long)stride_c, (long long)stride_y, box_x, box_c);
        return SC_ERR_CUDA;
    }
    return SC_OK;
}

// ---- float32 thresholds equivalent to a float64 comparison -----------------------------------
// (double)v > thr  <=>  v > f32_floor(thr);   (double)v >= thr <=> v > prev(f32_ceil(thr))
static float f32_floor(double t) {           // largest float <= t
    float f = (float)t;                       // round to nearest
    if ((double)f > t) f = nextafterf(f, -INFINITY);
    return f;
}
static float f32_ceil(double t) {            // smallest float >= t
    float f = (float)t;
    if ((double)f < t) f = nextafterf(f, INFINITY);
    return f;
}

static bool node_is_self(const sc_mask_node &nd, const float *cube, int64_t sc, int64_t sy) {
    return nd.data == nullptr || (nd.data == cube && nd.ds_c == sc && nd.ds_y == sy);
}

// Fold FINITE / CMP_SCALAR(GT,GE,LT,LE) on the cube itself joined by AND into one open
// float32 interval (lo, hi).  Returns false if the subtree is anything else.
static bool fold_interval(const sc_mask_desc *m, int i, const float *cube, int64_t sc, int64_t sy,
                          float *lo, float *hi, bool other = false) {
    // other = true: `cube` is a cube that is NOT the one being processed; leaves must reference exactly it
    auto node_is_self = [other](const sc_mask_node &n, const float *cb, int64_t c, int64_t y) {
        if (other) return n.data == cb && n.ds_c == c && n.ds_y == y;
        return n.data == nullptr || (n.data == cb && n.ds_c == c && n.ds_y == y);
    };
    const sc_mask_node &nd = m->nodes[i];
    switch (nd.kind) {
        case SC_MASK_FINITE:
            if (!node_is_self(nd, cube, sc, sy)) return false;
            // v > -inf && v < +inf  <=> finite
            return true;
        case SC_MASK_CMP_SCALAR: {
            if (!node_is_self(nd, cube, sc, sy)) return false;
            double t = nd.value;
            if (t != t) return false;                         // NaN threshold: leave to generic
            if (nd.op == SC_GT)      { float b = f32_floor(t); if (b > *lo) *lo = b; }
            else if (nd.op == SC_GE) { float b = nextafterf(f32_ceil(t), -INFINITY); if (f32_ceil(t) == -INFINITY) b = -INFINITY; if (b > *lo) *lo = b; }
            else if (nd.op == SC_LT) { float b = f32_ceil(t); if (b < *hi) *hi = b; }
            else if (nd.op == SC_LE) { float b = nextafterf(f32_floor(t), INFINITY); if (f32_floor(t) == INFINITY) b = INFINITY; if (b < *hi) *hi = b; }
            else return false;
            return true;
        }
        case SC_MASK_AND:
            return fold_interval(m, nd.a, cube, sc, sy, lo, hi, other) && fold_interval(m, nd.b, cube, sc, sy, lo, hi, other);
        default:
            return false;
    }
}

static bool subtree_has_finite(const sc_mask_desc *m, int i) {
    const sc_mask_node &nd = m->nodes[i];
    if (nd.kind == SC_MASK_FINITE) return true;
    if (nd.kind == SC_MASK_AND) return subtree_has_finite(m, nd.a) || subtree_has_finite(m, nd.b);
    return false;
}

int build_dev_mask(const sc_mask_desc *m, const float *cube, int64_t stride_c, int64_t stride_y,
                   DevMask *out, bool allow_other) {
    memset(out, 0, sizeof(*out));
    if (m == nullptr || m->n_nodes == 0) { out->mode = MODE_NONE; return SC_OK; }
    SC_CHECK_ARG(m->n_nodes > 0 && m->n_nodes <= SC_MASK_MAX_NODES, "mask: n_nodes=%d out of range", m->n_nodes);
    for (int i = 0; i < m->n_nodes; ++i) {
        const sc_mask_node &nd = m->nodes[i];
        SC_CHECK_ARG(nd.kind >= SC_MASK_FINITE && nd.kind <= SC_MASK_NOT, "mask: node %d has unknown kind %d", i, nd.kind);
        if (nd.kind >= SC_MASK_AND) {
            SC_CHECK_ARG(nd.a >= 0 && nd.a < i, "mask: node %d child a=%d must precede it", i, nd.a);
            if (nd.kind != SC_MASK_NOT)
                SC_CHECK_ARG(nd.b >= 0 && nd.b < i, "mask: node %d child b=%d must precede it", i, nd.b);
        }
        if (nd.kind == SC_MASK_CMP_SCALAR || nd.kind == SC_MASK_CMP_ARRAY)
            SC_CHECK_ARG(nd.op >= SC_GT && nd.op <= SC_NE, "mask: node %d has bad comparison op %d", i, nd.op);
        if (nd.kind == SC_MASK_CMP_ARRAY)
            SC_CHECK_ARG(nd.array && (nd.array_dtype == SC_F32 || nd.array_dtype == SC_F64), "mask: node %d needs a float32/float64 array", i);
        if (nd.kind == SC_MASK_BOOL)
            SC_CHECK_ARG(nd.array && nd.array_dtype == SC_U8, "mask: node %d needs a uint8 array", i);
    }
    // the node program is always carried (self references normalised to NULL) so a kernel may fall
    // back to interpreting it even when the interval form exists
    out->prog = *m;
    for (int i = 0; i < m->n_nodes; ++i) {
        sc_mask_node &nd = out->prog.nodes[i];
        if (nd.kind <= SC_MASK_CMP_ARRAY && node_is_self(nd, cube, stride_c, stride_y)) nd.data = nullptr;
    }
    float lo = -INFINITY, hi = INFINITY;
    int root = m->n_nodes - 1;
    // The interval form excludes +-inf at open ends, so it is only valid when isfinite is part
    // of the conjunction (it always is for cubes read from FITS, io/fits.py:214).
    if (subtree_has_finite(m, root) && fold_interval(m, root, cube, stride_c, stride_y, &lo, &hi)) {
        out->mode = MODE_INTERVAL; out->lo = lo; out->hi = hi;
        return SC_OK;
    }
    // the same conjunction on another cube (a smoothed cube under the mask of its source)
    for (int i = 0; allow_other && i < m->n_nodes; ++i) {
        const sc_mask_node &nd = m->nodes[i];
        if (nd.kind > SC_MASK_CMP_SCALAR || nd.data == nullptr || node_is_self(nd, cube, stride_c, stride_y)) continue;
        lo = -INFINITY; hi = INFINITY;
        if (subtree_has_finite(m, root) && fold_interval(m, root, nd.data, nd.ds_c, nd.ds_y, &lo, &hi, true)) {
            out->mode = MODE_INTERVAL_OTHER; out->lo = lo; out->hi = hi;
            out->other = nd.data; out->other_sc = nd.ds_c; out->other_sy = nd.ds_y;
            return SC_OK;
        }
        break;
    }
    out->mode = MODE_GENERIC;
    out->prog = *m;
    // normalise "self" references to NULL so the kernel reuses the value it already loaded
    for (int i = 0; i < m->n_nodes; ++i) {
        sc_mask_node &nd = out->prog.nodes[i];
        if (nd.kind <= SC_MASK_CMP_ARRAY && node_is_self(nd, cube, stride_c, stride_y)) nd.data = nullptr;
    }
    return SC_OK;
}

bool mask_is_self_only(const sc_mask_desc *m) {
    if (!m) return true;
    for (int i = 0; i < m->n_nodes; ++i) {
        const sc_mask_node &nd = m->nodes[i];
        if (nd.kind <= SC_MASK_CMP_ARRAY && nd.data != nullptr) return false;
        if (nd.kind == SC_MASK_CMP_ARRAY || nd.kind == SC_MASK_BOOL) return false;
    }
    return true;
}

}  // namespace scb

extern "C" {

const char *sc_last_error(void) { return scb::g_err; }
int sc_version(void) { return 1; }
int64_t sc_launch_count(void) { return scb::g_launches.load(); }
void sc_enable_kernel_timing(int enable) { scb::g_timing.store(enable ? 1 : 0); }

float sc_last_kernel_ms(int op) {
    if (op <= 0 || op >= 16) return 0.f;
    std::lock_guard<std::mutex> lk(scb::g_timer_mu);
    scb::OpTimer &t = scb::g_timers[op];
    if (!t.a || !t.valid) return 0.f;
    if (cudaEventSynchronize(t.b) != cudaSuccess) return 0.f;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t.a, t.b) != cudaSuccess) return 0.f;
    return ms;
}

size_t sc_workspace_bytes(int op, int64_t nchan, int64_t ny, int64_t nx, int64_t aux) {
    (void)ny; (void)nx;
    switch (op) {
        case SC_OP_MOMENTS:         return (size_t)nchan * 24 + 512;          // {d, d^2} table + offsets
        case SC_OP_SMOOTH_MOMENTS:  return (size_t)nchan * 24 + (size_t)aux * 8 + 1024;
        case SC_OP_SPECTRAL_SMOOTH: return (size_t)aux * 8 + 256;             // normalised taps
        case SC_OP_SPATIAL_SMOOTH:  return (size_t)aux * 8 + (size_t)nchan + 2048 + ((size_t)12 << 20) + 4096;   // taps + pass-through flags + the separable kernel's fix-up list (2^19 tiles of 16 bytes) and tile bitmap (4 MiB)
        case SC_OP_SPECTRAL_INTERP: return (size_t)aux * 32 + 256;            // per-output-channel LUT
        default:                    return 256;
    }
}

}  // extern "C"
