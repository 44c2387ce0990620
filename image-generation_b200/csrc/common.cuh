// Shared helpers for libb200grbm.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200grbm.h"
#include "../../include/b200grbm_spec.h"

namespace b200grbm {

// thread-local error text, returned by b200grbm_last_error()
char *error_buffer();
int32_t fail(int32_t code, const char *fmt, ...);
int32_t check_cuda(cudaError_t e, const char *what);
// verifies that the current device is sm_100 (there is no fallback path)
int32_t require_device();
int32_t sm_count();

#define B200_TRY(expr)                     \
    do {                                   \
        int32_t _rc = (expr);              \
        if (_rc != 0) return _rc;          \
    } while (0)

#define B200_CUDA(expr) B200_TRY(::b200grbm::check_cuda((expr), #expr))

__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }

}  // namespace b200grbm
