// stubs.cu -- entry points declared in include/sc_b200.h whose kernels are not written yet.
// They fail loudly (SC_ERR_UNSUPPORTED); nothing falls back to the CPU.
#include "common.cuh"

#define SC_STUB(name) \
    scb::set_error(name " is not implemented yet in this build of libsc_b200"); return SC_ERR_UNSUPPORTED;

extern "C" {

int sc_moments_spatial(const float *, int64_t, int64_t, int64_t, int64_t, int64_t, int,
                       const sc_mask_desc *, const double *, double, int, double *, double *, double *, void *) { SC_STUB("sc_moments_spatial") }

}
