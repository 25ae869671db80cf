// Shared helpers for libqrag.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/qrag.h"

namespace qrag {

// thread-local last-error text (qrag_last_error)
char* err_buf();
int set_error(int code, const char* fmt, ...);

#define QRAG_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return ::qrag::set_error(QRAG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,        \
                                     cudaGetErrorString(_e), __FILE__, __LINE__);          \
    } while (0)

#define QRAG_REQUIRE(cond, code, ...)                                                      \
    do {                                                                                   \
        if (!(cond)) return ::qrag::set_error(code, __VA_ARGS__);                          \
    } while (0)

// Launch-error check that does not synchronise.
#define QRAG_LAUNCH_CHECK(name)                                                            \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess)                                                             \
            return ::qrag::set_error(QRAG_ERR_CUDA, "launch of %s failed: %s", name,       \
                                     cudaGetErrorString(_e));                              \
    } while (0)

struct DeviceProps {
    int sm_count;
    int cc_major, cc_minor;
    int max_smem_optin;
    bool ok;
};
// cached per process for the current device; ok == false if no CUDA device
const DeviceProps& device_props();

// process-wide stream-overlap policy (qrag_set_overlap)
int overlap_mode();
// process-wide choice of the feature-map kernel (qrag_set_fmap_kernel)
int fmap_kernel_mode();

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int next_pow2(int64_t v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// 128-bit streaming load: read-only path, do not allocate in L1.
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// Canonical ordering of (score, position): descending score, ascending position.
// `a` sorts before `b`?
__device__ __forceinline__ bool before_desc(double sa, int ia, double sb, int ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}

}  // namespace qrag
