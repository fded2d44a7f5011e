// common.cu -- version, error string, device checks
#include <stdarg.h>

#include "common.cuh"

namespace gnx {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int require_blackwell() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device: %s (libgnx has no CPU fallback)", cudaGetErrorString(e));
        return 1;
    }
    int major = 0;
    GNX_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        set_error("device %d has compute capability %d.x; libgnx is built for sm_100a only", dev, major);
        return 1;
    }
    return 0;
}

int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return n;
}

}  // namespace gnx

extern "C" {

int gnx_version(void) { return GNX_VERSION; }

const char* gnx_last_error(void) { return gnx::g_err; }

int gnx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int d = 0; d < n; d++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
    }
    return ok;
}

}  // extern "C"
