// libptk_b200: version / error / device queries of the C ABI (include/ptk.h).
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include <nvtx3/nvToolsExt.h>

#include "ptk_common.cuh"

namespace ptk {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool pdl_enabled() {
    static const bool on = []() {
        const char *e = getenv("PTK_NO_PDL");
        return !(e && atoi(e) != 0);
    }();
    return on;
}

NvtxRange::NvtxRange(const char *name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

int sm_count() {
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return cached;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
            cached = n;
            cached_dev = dev;
        }
    }
    return cached;
}

}  // namespace ptk

extern "C" int ptk_version(void) { return PTK_ABI_VERSION; }

extern "C" const char *ptk_last_error(void) { return ptk::g_err; }

extern "C" uint64_t ptk_launch_count(void) { return ptk::g_launches.load(std::memory_order_relaxed); }

extern "C" int ptk_device_info(int device, int *sm_count, int *clock_khz, int *l2_bytes,
                               int *smem_optin, int *cc_major, int *cc_minor) {
    int v = 0;
    if (sm_count) { PTK_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)); *sm_count = v; }
    if (clock_khz) { PTK_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device)); *clock_khz = v; }
    if (l2_bytes) { PTK_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device)); *l2_bytes = v; }
    if (smem_optin) { PTK_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)); *smem_optin = v; }
    if (cc_major) { PTK_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device)); *cc_major = v; }
    if (cc_minor) { PTK_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device)); *cc_minor = v; }
    return PTK_OK;
}
