// Canonical exact scoring of fp32 rows against one fp32 query with fp64 accumulation.
//
// Used by the CUDA-core brute-force search (search_exact.cu) and by the exact rescoring stage of
// the tensor-core search (search_tc.cu).  Both must produce bit-identical scores for the same
// (query, row) pair, so the lane -> element map, the accumulation order and the reduction tree
// live here and nowhere else.  A row's score does not depend on which of the 4 batch slots it
// occupies nor on the other rows of the batch.
#pragma once

#include "common.cuh"

namespace qrag {

constexpr int XS_THREADS = 256;           // block size the query staging is defined for
constexpr int XS_WARPS = XS_THREADS / 32;
constexpr int XS_ROWS = 4;                // rows per warp batch

// Transposed butterfly over 4 (sum, |d|^2) pairs: lane L with (L & 7) == 0 ends with the sum of
// row L >> 3; its |d|^2 sits 4 lanes up.
__device__ __forceinline__ double xs_reduce4pairs(const double (&v)[8], int lane) {
    double w4[4], w2[2], w1;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double send = b4 ? v[i] : v[i + 4];
        const double keep = b4 ? v[i + 4] : v[i];
        w4[i] = keep + __shfl_xor_sync(FULL_MASK, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double send = b3 ? w4[i] : w4[i + 2];
        const double keep = b3 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(FULL_MASK, send, 8);
    }
    const double send = b2 ? w2[0] : w2[1];
    const double keep = b2 ? w2[1] : w2[0];
    w1 = keep + __shfl_xor_sync(FULL_MASK, send, 4);
    w1 += __shfl_xor_sync(FULL_MASK, w1, 2);
    w1 += __shfl_xor_sync(FULL_MASK, w1, 1);
    return w1;
}

// Stage the query as fp64 in shared memory (qs[D]) and return |q|^2.  All XS_THREADS threads of
// the block call it; `red` is XS_WARPS doubles of scratch.  Ends with a __syncthreads().
__device__ __forceinline__ double xs_stage_query(const float* __restrict__ qrow, int D, double* qs, double* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double part = 0.0;
    for (int i = tid; i < D; i += XS_THREADS) {
        const double v = (double)qrow[i];
        qs[i] = v;
        part = fma(v, v, part);
    }
    part = warp_sum(part);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    double nq2 = 0.0;
#pragma unroll
    for (int w = 0; w < XS_WARPS; ++w) nq2 += red[w];
    return nq2;
}

// 16 bytes of a row: from global memory through the streaming read-only path, or (STREAM = false) from wherever
// the caller staged the row -- shared memory filled by a bulk copy in tc_rescore_bulk_kernel.  The values, hence
// the scores, are the same.
template <bool STREAM>
__device__ __forceinline__ float4 xs_load4(const float4* p) {
    if (STREAM) return ldg_stream(p);
    return *p;
}

// One warp scores XS_ROWS rows.  l2: acc = sum (q - d)^2; otherwise acc = q.d and |d|^2.
// Returns the transposed-reduced value (see xs_reduce4pairs); nd2 receives |d|^2 of the same row.
template <bool VEC, bool STREAM = true>
__device__ __forceinline__ double xs_score4(const float* const (&rp)[XS_ROWS], const double* qs, int D, bool l2,
                                            int lane, double& nd2) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0;
    if (VEC) {
        const int D4 = D >> 2;
#pragma unroll 2
        for (int j = lane; j < D4; j += 32) {
            float4 v[XS_ROWS];
#pragma unroll
            for (int i = 0; i < XS_ROWS; ++i) v[i] = xs_load4<STREAM>(reinterpret_cast<const float4*>(rp[i]) + j);
            const double2 qa = *reinterpret_cast<const double2*>(qs + 4 * j);
            const double2 qb = *reinterpret_cast<const double2*>(qs + 4 * j + 2);
            const double qv[4] = {qa.x, qa.y, qb.x, qb.y};
#pragma unroll
            for (int i = 0; i < XS_ROWS; ++i) {
                const double d[4] = {(double)v[i].x, (double)v[i].y, (double)v[i].z, (double)v[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (l2) {
                        const double t = qv[e] - d[e];
                        acc[2 * i] = fma(t, t, acc[2 * i]);
                    } else {
                        acc[2 * i] = fma(qv[e], d[e], acc[2 * i]);
                        acc[2 * i + 1] = fma(d[e], d[e], acc[2 * i + 1]);
                    }
                }
            }
        }
    } else {
        for (int j = lane; j < D; j += 32) {
            const double qv = qs[j];
#pragma unroll
            for (int i = 0; i < XS_ROWS; ++i) {
                const double d = (double)__ldg(rp[i] + j);
                if (l2) {
                    const double t = qv - d;
                    acc[2 * i] = fma(t, t, acc[2 * i]);
                } else {
                    acc[2 * i] = fma(qv, d, acc[2 * i]);
                    acc[2 * i + 1] = fma(d, d, acc[2 * i + 1]);
                }
            }
        }
    }
    const double tot = xs_reduce4pairs(acc, lane);
    nd2 = __shfl_down_sync(FULL_MASK, tot, 4);
    return tot;
}

// L2 scoring that also returns q.d and |d|^2 of the same rows (the packed search + rerank form needs the
// amplitude fidelity of an L2 candidate from the one read of its row).  The L2 sum goes through the same
// instruction sequence as in xs_score4, q.d and |d|^2 through the sequence of the non-L2 form, so all three
// agree bit for bit with what the separate kernels compute.
template <bool VEC, bool STREAM = true>
__device__ __forceinline__ double xs_score4_l2dot(const float* const (&rp)[XS_ROWS], const double* qs, int D, int lane,
                                                  double& nd2, double& dot) {
    double acc[8], acd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = 0.0; acd[i] = 0.0; }
    if (VEC) {
        const int D4 = D >> 2;
#pragma unroll 2
        for (int j = lane; j < D4; j += 32) {
            float4 v[XS_ROWS];
#pragma unroll
            for (int i = 0; i < XS_ROWS; ++i) v[i] = xs_load4<STREAM>(reinterpret_cast<const float4*>(rp[i]) + j);
            const double2 qa = *reinterpret_cast<const double2*>(qs + 4 * j);
            const double2 qb = *reinterpret_cast<const double2*>(qs + 4 * j + 2);
            const double qv[4] = {qa.x, qa.y, qb.x, qb.y};
#pragma unroll
            for (int i = 0; i < XS_ROWS; ++i) {
                const double d[4] = {(double)v[i].x, (double)v[i].y, (double)v[i].z, (double)v[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double t = qv[e] - d[e];
                    acc[2 * i] = fma(t, t, acc[2 * i]);
                    acc[2 * i + 1] = fma(d[e], d[e], acc[2 * i + 1]);
                    acd[2 * i] = fma(qv[e], d[e], acd[2 * i]);
                }
            }
        }
    } else {
        for (int j = lane; j < D; j += 32) {
            const double qv = qs[j];
#pragma unroll
            for (int i = 0; i < XS_ROWS; ++i) {
                const double d = (double)__ldg(rp[i] + j);
                const double t = qv - d;
                acc[2 * i] = fma(t, t, acc[2 * i]);
                acc[2 * i + 1] = fma(d, d, acc[2 * i + 1]);
                acd[2 * i] = fma(qv, d, acd[2 * i]);
            }
        }
    }
    const double tot = xs_reduce4pairs(acc, lane);
    nd2 = __shfl_down_sync(FULL_MASK, tot, 4);
    dot = xs_reduce4pairs(acd, lane);
    return tot;
}

// Amplitude-encoded fidelity (q.d)^2 / (|q|^2 |d|^2) from the reduced sums: the expression of amp_fidelity.cu.
__device__ __forceinline__ double xs_fidelity(double dot, double nd2, double nq2) {
    const double den = nq2 * nd2;
    return den > 0.0 ? (dot * dot) / den : 0.0;
}

// Sort key (ascending = better) of a row from its reduced sums.
__device__ __forceinline__ double xs_key(int metric, double tot, double nd2, double nq2) {
    if (metric == QRAG_METRIC_L2) return tot;
    if (metric == QRAG_METRIC_IP) return -tot;
    const double den = nq2 * nd2;
    return den > 0.0 ? -(tot / sqrt(den)) : -0.0;
}

}  // namespace qrag
