// Segmented stable sort: the `sorted(scored, key=score, reverse=True)[:top_k]` of the reference
// (quantum.py:70-76, classical.py:302-308) for nq lists of C fp64 scores.
//
// Order: (score descending | ascending, input position ascending) -- a total order, so any correct sort is "stable".
//   C <= 4096   one CTA per list, bitonic sort of (key, position) in shared memory.
//   C >  4096   the reference sorts lists of any length, so the drop-in must too (no library sort on the path):
//               blocks of 4096 are sorted in shared memory into a workspace, then merged level by level in global
//               memory; an element's place in the merged run is its own index plus the number of elements of the
//               partner run that sort before it (binary search), so every level is one fully parallel kernel.
#include "common.cuh"
#include "sort.cuh"

namespace qrag {

constexpr int SS_BLOCK = QRAG_MAX_SORT_LEN;        // elements sorted in shared memory at once

__global__ void __launch_bounds__(256) sort_scores_kernel(const double* __restrict__ scores, int nq, int64_t C,
                                                          int top_k, int descending, int32_t* __restrict__ out_perm,
                                                          double* __restrict__ out_sorted) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int P = 1;
    while (P < C) P <<= 1;
    double* key = reinterpret_cast<double*>(smem_raw);
    int* tag = reinterpret_cast<int*>(key + P);
    for (int q = blockIdx.x; q < nq; q += gridDim.x) {
        const double* s = scores + (size_t)q * C;
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            const bool real = i < C;
            const double v = real ? s[i] : 0.0;
            key[i] = real ? (descending ? -v : v) : pos_inf();
            tag[i] = real ? i : TagPad<int>::value();
        }
        __syncthreads();
        block_bitonic_sort<int>(key, tag, P);
        for (int i = threadIdx.x; i < top_k; i += blockDim.x) {
            out_perm[(size_t)q * top_k + i] = tag[i];
            if (out_sorted) out_sorted[(size_t)q * top_k + i] = descending ? -key[i] : key[i];
        }
        __syncthreads();
    }
}

// level 0 of the long sort: CTA (block b, list q) sorts elements [b * 4096, (b + 1) * 4096) of list q into the workspace
// (keys negated for descending order; slots past C hold (+inf, INT_MAX) and sort last)
__global__ void __launch_bounds__(1024) sort_blocks_kernel(const double* __restrict__ scores, int64_t C, int64_t Cpad,
                                                           int descending, double* __restrict__ wkey, int* __restrict__ wtag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* key = reinterpret_cast<double*>(smem_raw);
    int* tag = reinterpret_cast<int*>(key + SS_BLOCK);
    const int64_t q = blockIdx.y, base = (int64_t)blockIdx.x * SS_BLOCK;
    const double* s = scores + q * C;
    for (int i = threadIdx.x; i < SS_BLOCK; i += blockDim.x) {
        const int64_t e = base + i;
        const bool real = e < C;
        const double v = real ? s[e] : 0.0;
        key[i] = real ? (descending ? -v : v) : pos_inf();
        tag[i] = real ? (int)e : TagPad<int>::value();
    }
    __syncthreads();
    block_bitonic_sort<int>(key, tag, SS_BLOCK);
    for (int i = threadIdx.x; i < SS_BLOCK; i += blockDim.x) {
        wkey[q * Cpad + base + i] = key[i];
        wtag[q * Cpad + base + i] = tag[i];
    }
}

// one merge level: runs of length L (sorted) -> runs of length 2L.  One thread per element; a run without a partner
// (odd count at this level) is copied.
__global__ void __launch_bounds__(256) merge_level_kernel(const double* __restrict__ ikey, const int* __restrict__ itag,
                                                          double* __restrict__ okey, int* __restrict__ otag, int64_t Cpad,
                                                          int64_t L) {
    const int64_t q = blockIdx.y;
    const double* k = ikey + q * Cpad;
    const int* t = itag + q * Cpad;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < Cpad; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t run = e / L, a = e - run * L;
        const int64_t partner = run ^ 1, pbase = partner * L, obase = (run >> 1) * 2 * L;
        const double ke = k[e];
        const int te = t[e];
        int64_t cnt = 0;
        if (pbase < Cpad) {                                       // elements of the partner run that sort before (ke, te)
            int64_t lo = 0, hi = (Cpad - pbase) < L ? (Cpad - pbase) : L;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                const double km = k[pbase + mid];
                if (km < ke || (km == ke && t[pbase + mid] < te)) lo = mid + 1;
                else hi = mid;
            }
            cnt = lo;
        }
        okey[q * Cpad + obase + a + cnt] = ke;
        otag[q * Cpad + obase + a + cnt] = te;
    }
}

__global__ void __launch_bounds__(256) sort_emit_kernel(const double* __restrict__ wkey, const int* __restrict__ wtag,
                                                        int64_t Cpad, int top_k, int descending,
                                                        int32_t* __restrict__ out_perm, double* __restrict__ out_sorted) {
    const int64_t q = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < top_k; i += gridDim.x * blockDim.x) {
        out_perm[q * top_k + i] = wtag[q * Cpad + i];
        if (out_sorted) {
            const double kv = wkey[q * Cpad + i];
            out_sorted[q * top_k + i] = descending ? -kv : kv;
        }
    }
}

static size_t ss_align(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace qrag

using namespace qrag;

extern "C" int qrag_sort_scores_workspace(int nq, int64_t C, size_t* bytes) {
    QRAG_REQUIRE(bytes != nullptr && nq >= 0 && C >= 0, QRAG_ERR_INVALID, "bad argument");
    *bytes = 0;
    if (C <= SS_BLOCK) return QRAG_OK;
    QRAG_REQUIRE(C < ((int64_t)1 << 31) - SS_BLOCK, QRAG_ERR_UNSUPPORTED, "C=%lld: positions are int32", (long long)C);
    const int64_t Cpad = ceil_div(C, SS_BLOCK) * SS_BLOCK;
    *bytes = 2 * (ss_align((size_t)nq * Cpad * 8) + ss_align((size_t)nq * Cpad * 4)) + 256;
    return QRAG_OK;
}

extern "C" int qrag_sort_scores_stable(const double* scores, int nq, int64_t C, int top_k, int descending,
                                       int32_t* out_perm, double* out_sorted, void* workspace, size_t workspace_bytes,
                                       void* stream) {
    QRAG_REQUIRE(scores && out_perm, QRAG_ERR_INVALID, "null pointer argument");
    QRAG_REQUIRE(nq >= 0 && C >= 0, QRAG_ERR_INVALID, "bad sizes nq=%d C=%lld", nq, (long long)C);
    QRAG_REQUIRE(top_k >= 0 && top_k <= C, QRAG_ERR_INVALID, "top_k=%d outside [0, C=%lld]", top_k, (long long)C);
    if (nq == 0 || C == 0 || top_k == 0) return QRAG_OK;
    const DeviceProps& dp = device_props();
    QRAG_REQUIRE(dp.ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    cudaStream_t st = (cudaStream_t)stream;
    if (C <= SS_BLOCK) {
        const size_t smem = (size_t)next_pow2(C) * (sizeof(double) + sizeof(int));
        if (smem > 48 * 1024)
            QRAG_CUDA_CHECK(cudaFuncSetAttribute(sort_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = nq < dp.sm_count * 32 ? nq : dp.sm_count * 32;
        sort_scores_kernel<<<grid, 256, smem, st>>>(scores, nq, C, top_k, descending, out_perm, out_sorted);
        QRAG_LAUNCH_CHECK("sort_scores_kernel");
        return QRAG_OK;
    }
    size_t need = 0;
    int rc = qrag_sort_scores_workspace(nq, C, &need);
    if (rc) return rc;
    QRAG_REQUIRE(workspace != nullptr && workspace_bytes >= need, QRAG_ERR_WORKSPACE,
                 "lists of %lld scores need a workspace of %zu bytes (qrag_sort_scores_workspace), got %zu", (long long)C, need,
                 workspace_bytes);
    QRAG_REQUIRE(nq <= 65535, QRAG_ERR_UNSUPPORTED, "long sort handles up to 65535 lists per call (got %d)", nq);
    const int64_t nblocks = ceil_div(C, SS_BLOCK), Cpad = nblocks * SS_BLOCK;
    unsigned char* base = reinterpret_cast<unsigned char*>(((size_t)workspace + 255) / 256 * 256);
    double* key[2];
    int* tag[2];
    key[0] = reinterpret_cast<double*>(base);
    tag[0] = reinterpret_cast<int*>(base + ss_align((size_t)nq * Cpad * 8));
    unsigned char* half = base + ss_align((size_t)nq * Cpad * 8) + ss_align((size_t)nq * Cpad * 4);
    key[1] = reinterpret_cast<double*>(half);
    tag[1] = reinterpret_cast<int*>(half + ss_align((size_t)nq * Cpad * 8));
    const size_t smem = (size_t)SS_BLOCK * (sizeof(double) + sizeof(int));
    QRAG_CUDA_CHECK(cudaFuncSetAttribute(sort_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sort_blocks_kernel<<<dim3((unsigned)nblocks, (unsigned)nq), 1024, smem, st>>>(scores, C, Cpad, descending, key[0], tag[0]);
    QRAG_LAUNCH_CHECK("sort_blocks_kernel");
    int cur = 0;
    const int64_t gx = ceil_div(Cpad, 256) < 4096 ? ceil_div(Cpad, 256) : 4096;
    for (int64_t L = SS_BLOCK; L < Cpad; L <<= 1) {
        merge_level_kernel<<<dim3((unsigned)gx, (unsigned)nq), 256, 0, st>>>(key[cur], tag[cur], key[cur ^ 1], tag[cur ^ 1], Cpad, L);
        QRAG_LAUNCH_CHECK("merge_level_kernel");
        cur ^= 1;
    }
    sort_emit_kernel<<<dim3((unsigned)(ceil_div(top_k, 256) < 1024 ? ceil_div(top_k, 256) : 1024), (unsigned)nq), 256, 0, st>>>(
        key[cur], tag[cur], Cpad, top_k, descending, out_perm, out_sorted);
    QRAG_LAUNCH_CHECK("sort_emit_kernel");
    return QRAG_OK;
}
