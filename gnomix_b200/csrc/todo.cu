// todo.cu -- entry points declared in include/gnx.h whose kernels are not written
// yet.  They fail loudly (no CPU fallback).  Each moves to its own file when built.
#include "common.cuh"

extern "C" {

int gnx_svc_model_create(gnx_svc_t** out, int, int64_t, int64_t, int64_t, const int8_t*, const int32_t*, const double*, const double*,
                         const double*, const double*, const int32_t*, int) {
    if (out) *out = nullptr;
    gnx::set_error("gnx_svc_model_create: CovRSK kernel (K2/K3) not built yet");
    return 9;
}
void gnx_svc_model_destroy(gnx_svc_t*) {}
int gnx_svc_predict(const gnx_svc_t*, const int8_t*, int64_t, int64_t, double*, void*) {
    gnx::set_error("gnx_svc_predict: CovRSK kernel (K2/K3) not built yet");
    return 9;
}
int gnx_svc_kernel_window(const gnx_svc_t*, int, const int8_t*, int64_t, int64_t, int32_t*, void*) {
    gnx::set_error("gnx_svc_kernel_window: CovRSK kernel (K2) not built yet");
    return 9;
}
}
