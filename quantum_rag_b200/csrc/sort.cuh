// Block-level ordering primitives shared by the rerank, search and merge kernels.
//
// All lists are ordered ascending by the pair (key, tag): callers that want
// "score descending, position ascending" (quantum.py:70-72) store key = -score,
// which is an exact, order-reversing map on finite doubles and on +-inf.
#pragma once

#include "common.cuh"
#include "tma.cuh"

namespace qrag {

template <typename Tag>
__device__ __forceinline__ bool pair_less(double ka, Tag ta, double kb, Tag tb) {
    return (ka < kb) || (ka == kb && ta < tb);
}

// In-place bitonic sort of P (power of two) pairs held in shared memory.
// Called by `nt` threads (ids 0..nt-1) that synchronise on barrier `bar_id` (0 = the whole
// block via __syncthreads, otherwise a named barrier over exactly those nt threads);
// ends with a barrier.
template <typename Tag>
__device__ void block_bitonic_sort(double* __restrict__ key, Tag* __restrict__ tag, int P, int tid, int nt,
                                   int bar_id) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += nt) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // bit j clear
                const int hi = lo | j;
                const double kl = key[lo], kh = key[hi];
                const Tag tl = tag[lo], th = tag[hi];
                const bool ascending = (lo & k) == 0;
                const bool hi_first = pair_less<Tag>(kh, th, kl, tl);
                if (hi_first == ascending) {
                    key[lo] = kh; key[hi] = kl;
                    tag[lo] = th; tag[hi] = tl;
                }
            }
            if (bar_id == 0) __syncthreads();
            else named_bar_sync(bar_id, nt);
        }
    }
}

template <typename Tag>
__device__ __forceinline__ void block_bitonic_sort(double* __restrict__ key, Tag* __restrict__ tag, int P) {
    block_bitonic_sort<Tag>(key, tag, P, threadIdx.x, blockDim.x, 0);
}

template <typename Tag> struct TagPad;
template <> struct TagPad<int> { static __device__ __forceinline__ int value() { return 0x7fffffff; } };
template <> struct TagPad<long long> {
    static __device__ __forceinline__ long long value() { return 0x7fffffffffffffffLL; }
};

__device__ __forceinline__ double pos_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

}  // namespace qrag
