// libqrag.so: error state, device query, version.
#include "common.cuh"

#include <atomic>
#include <mutex>

namespace qrag {

static thread_local char g_err[512] = "";

char* err_buf() { return g_err; }

static std::atomic<int> g_overlap{QRAG_OVERLAP_SAFE};
int overlap_mode() { return g_overlap.load(std::memory_order_relaxed); }
static std::atomic<int> g_fmap_kernel{QRAG_FMAP_AUTO};
int fmap_kernel_mode() { return g_fmap_kernel.load(std::memory_order_relaxed); }

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

const DeviceProps& device_props() {
    static DeviceProps props[64];
    static bool init[64];
    static std::mutex mu;
    static DeviceProps none{0, 0, 0, 0, false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        cudaGetLastError();
        return none;
    }
    std::lock_guard<std::mutex> lock(mu);
    if (!init[dev]) {
        DeviceProps d{};
        int v = 0;
        bool ok = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess;
        d.sm_count = v;
        ok = ok && cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess;
        ok = ok && cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev) == cudaSuccess;
        ok = ok && cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess;
        d.ok = ok && d.sm_count > 0;
        props[dev] = d;
        init[dev] = true;
    }
    return props[dev];
}

}  // namespace qrag

extern "C" const char* qrag_last_error(void) { return qrag::err_buf(); }

extern "C" int qrag_version(void) { return 100; }   // 0.1.0

extern "C" int qrag_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    const qrag::DeviceProps& d = qrag::device_props();
    QRAG_REQUIRE(d.ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    if (sm_count) *sm_count = d.sm_count;
    if (cc_major) *cc_major = d.cc_major;
    if (cc_minor) *cc_minor = d.cc_minor;
    return QRAG_OK;
}

extern "C" int qrag_set_overlap(int mode) {
    QRAG_REQUIRE(mode >= QRAG_OVERLAP_NONE && mode <= QRAG_OVERLAP_INTERLEAVED, QRAG_ERR_INVALID,
                 "overlap mode %d (expected QRAG_OVERLAP_NONE / _SAFE / _INPUTS_STABLE / _INTERLEAVED)", mode);
    qrag::g_overlap.store(mode, std::memory_order_relaxed);
    return QRAG_OK;
}

extern "C" int qrag_get_overlap(void) { return qrag::overlap_mode(); }

extern "C" int qrag_set_fmap_kernel(int mode) {
    QRAG_REQUIRE(mode == QRAG_FMAP_AUTO || mode == QRAG_FMAP_GENERIC, QRAG_ERR_INVALID,
                 "feature-map kernel mode %d (expected QRAG_FMAP_AUTO / QRAG_FMAP_GENERIC)", mode);
    qrag::g_fmap_kernel.store(mode, std::memory_order_relaxed);
    return QRAG_OK;
}
