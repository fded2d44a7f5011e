// common.cuh -- error plumbing shared by every translation unit of libgnx.so
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/gnx.h"
#include "../../include/gnx_math.h"

namespace gnx {

void set_error(const char* fmt, ...);

#define GNX_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) {                                                            \
            gnx::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

#define GNX_REQUIRE(cond, ...)                \
    do {                                      \
        if (!(cond)) {                        \
            gnx::set_error(__VA_ARGS__);      \
            return 2;                         \
        }                                     \
    } while (0)

// Fails (non-zero + message) unless the current device is sm_100-class.
int require_blackwell();
int sm_count();

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace gnx
