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

// NVTX range around an ABI entry point (host side, header-only NVTX v3: a no-op unless a profiler is attached).
// Timelines of Nsight Systems / `ncu --nvtx` then show ptk_chamfer_fwd, ptk_gcn_stack_bwd, ... around their kernels.
struct NvtxRange {
    explicit NvtxRange(const char *name);
    ~NvtxRange();
};
#define PTK_NVTX(name) ::ptk::NvtxRange _ptk_nvtx_range(name)

// ---- programmatic dependent launch (PDL).  The kernels of one GCN pass are 5-135 us each and follow one another
// on a single stream; a plain launch starts a kernel ~2 us after its predecessor drained (measured: 1.0 ms of idle
// gaps over the ~550 kernels of a reconstruction step).  Launched through launch_pdl(), a kernel may become resident
// before its predecessor has drained; it must call pdl_wait() before touching global memory (that blocks until the
// whole predecessor grid has completed and flushed).  Every kernel launched this way waits, so ordering stays
// transitive along the stream.
// PTK_NO_PDL=1 launches the same kernels without the attribute (A/B measurements).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Early trigger: measured HARMFUL here (reconstruction step 20.10 ms with it, 18.97 ms without, 19.62 ms without PDL):
// successor CTAs that become resident during the tail of a multi-wave kernel take SM slots and skew the successor's
// own CTA placement.  Without it the successor is still launched early (its launch latency overlaps the predecessor)
// and released when the predecessor completes.  -DPTK_PDL_EARLY_TRIGGER re-enables it for experiments.
__device__ __forceinline__ void pdl_launch_dependents() {
#ifdef PTK_PDL_EARLY_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
#endif

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
