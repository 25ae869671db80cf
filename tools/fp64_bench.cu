// Micro-benchmark: FP64 FMA issue rate per scheduler as a function of resident warps and of the number of
// independent dependency chains per thread.  Answers why fmap_warp_kernel (2 warps per scheduler) keeps the FP64
// pipe 61 % busy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_bench tools/fp64_bench.cu && /tmp/fp64_bench
#include <cuda_runtime.h>
#include <stdio.h>

template <int ILP>
__global__ void bench(int iters, double t, double* sink, long long* cycles) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
    __syncthreads();
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(t, a[(i + 1) % ILP], a[i]);   // butterfly-like: reads a neighbour
    }
    const long long c1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = c1 - c0;
}

// Same arithmetic, but the body is R x 16 DFMAs of straight-line code (no inner loop): does instruction fetch keep up
// when a warp streams through tens of KB of unique instructions, as the fully unrolled fmap_warp_kernel does?
template <int R>
__global__ void bench_straight(int iters, double t, double* sink, long long* cycles, int skew) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    __syncthreads();
    if (skew) {                                        // desynchronise the warps: each streams a different part of the body
        const long long until = clock64() + (long long)(threadIdx.x >> 5) * (R * 16 * 4 / (blockDim.x >> 5));
        while (clock64() < until) {}
    }
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma(t, a[(i + 1 + r % 7) % 16], a[i]);
    }
    const long long c1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456) sink[0] = s;
    if (threadIdx.x == blockDim.x - 32 && blockIdx.x == 0) cycles[0] = c1 - c0;
}

template <int R>
void run_straight(int warps_per_sched, double* sink, long long* cyc, int skew) {
    const int iters = 16384 / R;
    bench_straight<R><<<148, warps_per_sched * 128>>>(iters, 1e-9, sink, cyc, skew);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double per_sched = (double)iters * R * 16 * warps_per_sched;
    printf("warps/scheduler %d  straight-line body %5d DFMAs (%4d KB of code)%s: %.2f clk per warp-DFMA per scheduler\n",
           warps_per_sched, R * 16, R * 16 * 16 / 1024, skew ? ", warps skewed" : "", c / per_sched);
}

template <int ILP>
void run(int warps_per_sched, double* sink, long long* cyc) {
    const int iters = 2000;
    bench<ILP><<<148, warps_per_sched * 128>>>(iters, 1e-9, sink, cyc);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double per_sched = (double)iters * 8 * ILP * warps_per_sched;      // warp-level DFMAs per scheduler
    printf("warps/scheduler %d  chains/thread %2d : %.2f clk per warp-DFMA per scheduler (pipe floor 2.00) -> %.0f %% of peak\n",
           warps_per_sched, ILP, c / per_sched, 200.0 * per_sched / c);
}

int main() {
    double* sink; long long* cyc;
    cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
    for (int w = 1; w <= 4; ++w) { run<2>(w, sink, cyc); run<4>(w, sink, cyc); run<8>(w, sink, cyc); run<16>(w, sink, cyc); run<32>(w, sink, cyc); if (w <= 2) { run<64>(w, sink, cyc); run<100>(w, sink, cyc); } }
    for (int sk = 0; sk <= 1; ++sk)
        for (int w = 1; w <= 2; ++w) { run_straight<8>(w, sink, cyc, sk); run_straight<64>(w, sink, cyc, sk); run_straight<96>(w, sink, cyc, sk); run_straight<128>(w, sink, cyc, sk); run_straight<192>(w, sink, cyc, sk); run_straight<256>(w, sink, cyc, sk); run_straight<512>(w, sink, cyc, sk); run_straight<1024>(w, sink, cyc, sk); }
    return 0;
}
