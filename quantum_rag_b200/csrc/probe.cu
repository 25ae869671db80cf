// Measured device peaks the benchmarks divide by where MEASURED_PEAKS.json has no entry: the FP64 FMA rate
// (the statevector kernels K1 / K1b are bound by the FP64 pipe, not by HBM or the tensor cores).
#include "common.cuh"

namespace qrag {

// 8 independent dependency chains per thread, 4 warps per scheduler: saturates the FP64 pipe
// (tools/fp64_bench.cu: one warp with >= 8 chains already issues a DFMA every 2 clocks per scheduler).
__global__ void __launch_bounds__(512) fp64_rate_kernel(int iters, double t, double* sink) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(t, a[(i + 1) & 7], a[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;                     // never true: keeps the chains alive
}

}  // namespace qrag

using namespace qrag;

extern "C" int qrag_probe_fp64_fma_rate(double* fma_per_s, double* scratch, void* stream) {
    QRAG_REQUIRE(fma_per_s && scratch, QRAG_ERR_INVALID, "null pointer argument");
    const DeviceProps& dp = device_props();
    QRAG_REQUIRE(dp.ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    cudaStream_t st = (cudaStream_t)stream;
    const int iters = 6000, grid = dp.sm_count * 2;
    cudaEvent_t e0, e1;
    QRAG_CUDA_CHECK(cudaEventCreate(&e0));
    QRAG_CUDA_CHECK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {                // the first repetition warms the clocks up
        QRAG_CUDA_CHECK(cudaEventRecord(e0, st));
        fp64_rate_kernel<<<grid, 512, 0, st>>>(iters, 1e-9, scratch);
        QRAG_CUDA_CHECK(cudaEventRecord(e1, st));
        QRAG_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        QRAG_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = (double)grid * 512 * iters * 64 / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    QRAG_LAUNCH_CHECK("fp64_rate_kernel");
    *fma_per_s = best;
    return QRAG_OK;
}
