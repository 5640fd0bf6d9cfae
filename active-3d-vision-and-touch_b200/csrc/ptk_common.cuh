// Shared helpers for the libptk_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptk.h"

namespace ptk {

void set_error(const char *fmt, ...);

#define PTK_CHECK_CUDA(expr)                                                          \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            ::ptk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                             __FILE__, __LINE__);                                     \
            return PTK_ERR_CUDA;                                                      \
        }                                                                             \
    } while (0)

#define PTK_REQUIRE(cond, code, ...)         \
    do {                                     \
        if (!(cond)) {                       \
            ::ptk::set_error(__VA_ARGS__);   \
            return (code);                   \
        }                                    \
    } while (0)

void count_launch();  // ptk_launch_count(): one tick per kernel launch of this library

// after EVERY <<<>>> of this library, exactly once
#define PTK_CHECK_LAUNCH()                   \
    do {                                     \
        ::ptk::count_launch();               \
        PTK_CHECK_CUDA(cudaGetLastError());  \
    } while (0)

inline cudaStream_t as_stream(ptk_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // cached cudaDevAttrMultiProcessorCount of the current device

// Development override: integer value of environment variable `name`, read once per call site
// (0 / unset = keep the built-in choice).  Used by the tuning tools only.
#define PTK_TUNING_ENV(name)                                       \
    ([]() -> int {                                                 \
        static const int v = []() {                                \
            const char *e = getenv(name);                          \
            return e ? atoi(e) : 0;                                \
        }();                                                       \
        return v;                                                  \
    }())

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace ptk
