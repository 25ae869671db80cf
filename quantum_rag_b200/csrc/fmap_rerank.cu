// Feature-map rerank (BASELINE config 5 as a RERANK: amplitude state + L reference layers at n = 10, top_k of C
// candidates per query) by filter-then-certify, the design of the tensor-core search applied to the statevector path:
//
//   filter   every candidate's state evolved in complex64 (fmap_warp_kernel<float>: normalisation, gate parameters and
//            the overlap in fp64, the 1024 amplitudes in fp32 -- half the registers, 12 warps per SM instead of 8, and
//            the FP32 pipe at twice the FP64 pipe's rate): approximate fidelities F32 with |F32 - F| <= delta;
//   select   per query s_k = k-th largest F32; U = {c : F32[c] >= s_k - 2 delta}, in candidate order.  The exact k-th best
//            e_k >= s_k - delta, and every candidate with F >= e_k has F32 >= e_k - delta >= s_k - 2 delta, so U holds the
//            exact top-k INCLUDING every candidate tied with its last member;
//   certify  the members of U (typically k + 1 or 2 of 1000) re-evolved in complex128 by the exact kernel
//            (fmap_warp_kernel<double>, gathered rows): the scores that are returned;
//   final    U sorted by (exact score desc, candidate position asc) -- QuantumReranker.rerank's stable sort
//            (quantum.py:70-76) -- and cut to top_k.
// Rankings and returned scores are those of the all-complex128 path, bit for bit; a query whose U exceeds its list
// capacity is flagged in `status` (never silent) and must be rerun through qrag_amp_fidelity + qrag_sort_scores_stable.
//
// delta.  Every step of the evolution is a norm-preserving (up to a common scale) linear map applied in floating point;
// gate parameters are evaluated in fp64 and rounded to fp32 once (relative error u = 2^-24 each):
//   RY butterfly, scaled form  a_j' = fma(-t, a_k, a_j): one rounding per output and one rounded coefficient, 2u;
//              direct form  a_j' = fma(-s, a_k, c a_j): two roundings and two rounded coefficients, 4u;
//   RZ diagonal  a_j' = a_j * (tab[j] * cst): each of the two tables is a product of 5 rounded unit phases (5 x (u + 3u)
//              for the fp32 complex products) times, in the scaled form, 5 rounded cosines (5u + 4u + 2u): 31u per
//              table; the two complex products that apply them 3u each: 68u.
// Per layer 10 x 4u + 68u = 108u, plus 2u for the rounded initial amplitudes: the computed state satisfies
// |psi32 - psi| <= eta = (108 L + 2) u (first order; the maps have condition number 1, so perturbations add).  With both
// states perturbed, |<d|q>| <= 1:  |F32 - F| <= 2 (2 eta + eta^2) + (2 eta + eta^2)^2 < 4.5 eta.  delta = 2 x 4.5 eta
// (a factor 2 of safety on a worst-case bound; measured: tests/test_gpu_amplitude.py, max |F32 - F| ~ 0.01 delta).
#include "common.cuh"
#include "sort.cuh"

namespace qrag {

int fmap_warp_try(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx, int64_t C, int D,
                  int n_qubits, int layers, double* out64, float* out32, cudaStream_t st, bool* handled);       // fmap_warp.cu
int fmap_warp_filter_try(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx, int64_t C,
                         int D, int n_qubits, int layers, double* out64, cudaStream_t st, bool* handled);        // fmap_warp.cu

static double fmap_filter_error_bound(int layers) {
    const double u = 5.9604644775390625e-08;            // 2^-24
    const double eta = (108.0 * layers + 2.0) * u;
    return 2.0 * 4.5 * eta;
}

static int fr_list_cap(int64_t C, int top_k) {
    int64_t cap = 2 * (int64_t)top_k + 32;
    if (cap < 64) cap = 64;
    if (cap > C) cap = C;
    return (int)cap;
}

struct FrParams {
    const double* approx;      // [nq, C] filter scores (-inf = padding candidate)
    const int64_t* idx;        // [nq, C] or null (dense candidates)
    int64_t C; int cap; int top_k; double delta;
    int* list;                 // [nq, cap] candidate positions of U, ascending, -1 padded
    int64_t* gather;           // [nq, cap] row of the (flattened) candidate array for the certify pass, -1 padded
    double* exact;             // [nq, cap] certified scores
    int32_t* status;           // [nq]
    double* out_scores; int32_t* out_pos; int64_t* out_ids;
};

// one CTA per query: s_k by a bitonic sort of a copy of the scores, then U in candidate order (ordered compaction)
__global__ void __launch_bounds__(256) fr_select_kernel(const FrParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* key = reinterpret_cast<double*>(smem_raw);                 // [P] -score, ascending
    __shared__ int s_warp[8];
    __shared__ double s_thr;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = (int)p.C;
    int P = 1;
    while (P < C) P <<= 1;
    const double* a = p.approx + (size_t)q * C;
    for (int i = tid; i < P; i += blockDim.x) key[i] = i < C ? -a[i] : pos_inf();
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += blockDim.x) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
                const double kl = key[lo], kh = key[hi];
                if ((kh < kl) == ((lo & k) == 0)) { key[lo] = kh; key[hi] = kl; }
            }
            __syncthreads();
        }
    if (tid == 0) s_thr = -key[p.top_k - 1] - 2.0 * p.delta;           // s_k - 2 delta (-inf stays -inf)
    __syncthreads();
    const double thr = s_thr;
    // ordered compaction: thread t owns positions [t * per, (t + 1) * per)
    const int per = (C + blockDim.x - 1) / blockDim.x;
    const int c0 = tid * per, c1 = min(C, c0 + per);
    int mine = 0;
    for (int c = c0; c < c1; ++c) mine += (a[c] >= thr) && (a[c] > -pos_inf());
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int base = incl - mine;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    int total = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += s_warp[w];
    int* list = p.list + (size_t)q * p.cap;
    int64_t* gather = p.gather + (size_t)q * p.cap;
    for (int c = c0; c < c1; ++c) {
        if ((a[c] >= thr) && (a[c] > -pos_inf())) {
            if (base < p.cap) {
                list[base] = c;
                gather[base] = p.idx ? p.idx[(size_t)q * C + c] : (int64_t)q * C + c;
            }
            ++base;
        }
    }
    for (int i = total + tid; i < p.cap; i += blockDim.x) { list[i] = -1; gather[i] = -1; }
    if (tid == 0) p.status[q] = total > p.cap ? 1 : 0;
}

// one CTA per query: U by (exact score desc, list index asc = candidate position asc), top_k out
__global__ void __launch_bounds__(256) fr_final_kernel(const FrParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int P = 1;
    while (P < p.cap) P <<= 1;
    double* key = reinterpret_cast<double*>(smem_raw);
    int* tag = reinterpret_cast<int*>(key + P);
    const int q = blockIdx.x, tid = threadIdx.x;
    const int* list = p.list + (size_t)q * p.cap;
    const double* ex = p.exact + (size_t)q * p.cap;
    for (int i = tid; i < P; i += blockDim.x) {
        const bool real = i < p.cap && list[i] >= 0;
        key[i] = real ? -ex[i] : pos_inf();
        tag[i] = real ? i : TagPad<int>::value();
    }
    __syncthreads();
    block_bitonic_sort<int>(key, tag, P);
    for (int i = tid; i < p.top_k; i += blockDim.x) {
        const int j = tag[i];
        const bool real = j != TagPad<int>::value();
        const int pos = real ? list[j] : -1;
        p.out_scores[(size_t)q * p.top_k + i] = real ? -key[i] : -pos_inf();
        p.out_pos[(size_t)q * p.top_k + i] = pos;
        if (p.out_ids) p.out_ids[(size_t)q * p.top_k + i] = (real && p.idx) ? p.idx[(size_t)q * p.C + pos] : -1;
    }
}

static size_t fr_align(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace qrag

using namespace qrag;

extern "C" int qrag_fmap_rerank_workspace(int nq, int64_t C, int top_k, size_t* bytes) {
    QRAG_REQUIRE(bytes != nullptr && nq >= 0 && C >= 0 && top_k >= 1, QRAG_ERR_INVALID, "bad argument");
    const int cap = fr_list_cap(C, top_k);
    *bytes = fr_align((size_t)nq * C * 8) + fr_align((size_t)nq * cap * 4) + fr_align((size_t)nq * cap * 8) +
             fr_align((size_t)nq * cap * 8) + 256;
    return QRAG_OK;
}

extern "C" int qrag_fmap_filter_error_bound(int layers, double* delta) {
    QRAG_REQUIRE(delta != nullptr && layers >= 1, QRAG_ERR_INVALID, "bad argument");
    *delta = fmap_filter_error_bound(layers);
    return QRAG_OK;
}

// diagnostic: the filter pass alone (complex64 evolution), for measuring its error against the bound
extern "C" int qrag_fmap_filter_scores(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx,
                                       int64_t C, int D, int n_qubits, int layers, double* out64, void* stream) {
    QRAG_REQUIRE(Q && out64 && ((cand != nullptr) != (X != nullptr && idx != nullptr)), QRAG_ERR_INVALID, "bad argument");
    QRAG_REQUIRE(cand != nullptr || N >= 1, QRAG_ERR_INVALID, "X needs its row count N >= 1");
    if (nq == 0 || C == 0) return QRAG_OK;
    bool handled = false;
    int rc = fmap_warp_filter_try(Q, nq, cand, X, N, idx, C, D, n_qubits, layers, out64, (cudaStream_t)stream, &handled);
    if (rc) return rc;
    QRAG_REQUIRE(handled, QRAG_ERR_UNSUPPORTED, "the complex64 filter serves n_qubits == 10, 1 <= D <= 1024, layers >= 1");
    return QRAG_OK;
}

extern "C" int qrag_fmap_rerank(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx,
                                int64_t C, int D, int n_qubits, int layers, int top_k, double* out_scores, int32_t* out_pos,
                                int64_t* out_ids, int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
    QRAG_REQUIRE(Q && out_scores && out_pos && status, QRAG_ERR_INVALID, "null pointer argument");
    QRAG_REQUIRE((cand != nullptr) != (X != nullptr && idx != nullptr), QRAG_ERR_INVALID,
                 "pass either cand, or X together with idx");
    QRAG_REQUIRE(cand != nullptr || N >= 1, QRAG_ERR_INVALID, "X needs its row count N >= 1 (got %lld)", (long long)N);
    QRAG_REQUIRE(nq >= 0 && D >= 1, QRAG_ERR_INVALID, "bad sizes nq=%d D=%d", nq, D);
    QRAG_REQUIRE(C >= 1 && C <= QRAG_MAX_SORT_LEN, QRAG_ERR_UNSUPPORTED, "feature-map rerank needs 1 <= C <= %d (got %lld)",
                 QRAG_MAX_SORT_LEN, (long long)C);
    QRAG_REQUIRE(top_k >= 1 && top_k <= C, QRAG_ERR_INVALID, "top_k=%d outside [1, C=%lld]", top_k, (long long)C);
    QRAG_REQUIRE(out_ids == nullptr || idx != nullptr, QRAG_ERR_INVALID, "out_ids needs idx");
    QRAG_REQUIRE(n_qubits == 10 && layers >= 1 && D <= 1024, QRAG_ERR_UNSUPPORTED,
                 "feature-map rerank serves n_qubits == 10, D <= 1024, layers >= 1 (got n=%d D=%d layers=%d)", n_qubits, D, layers);
    QRAG_REQUIRE((int64_t)nq * C < ((int64_t)1 << 40), QRAG_ERR_UNSUPPORTED, "batch too large");
    if (nq == 0) return QRAG_OK;
    size_t need = 0;
    int rc = qrag_fmap_rerank_workspace(nq, C, top_k, &need);
    if (rc) return rc;
    QRAG_REQUIRE(workspace != nullptr && workspace_bytes >= need, QRAG_ERR_WORKSPACE,
                 "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    const DeviceProps& dp = device_props();
    QRAG_REQUIRE(dp.ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    cudaStream_t st = (cudaStream_t)stream;
    const int cap = fr_list_cap(C, top_k);
    unsigned char* base = reinterpret_cast<unsigned char*>(((size_t)workspace + 255) / 256 * 256);
    FrParams p{};
    double* approx = reinterpret_cast<double*>(base);
    base += fr_align((size_t)nq * C * 8);
    p.list = reinterpret_cast<int*>(base);
    base += fr_align((size_t)nq * cap * 4);
    p.gather = reinterpret_cast<int64_t*>(base);
    base += fr_align((size_t)nq * cap * 8);
    p.exact = reinterpret_cast<double*>(base);
    p.approx = approx; p.idx = idx; p.C = C; p.cap = cap; p.top_k = top_k; p.delta = fmap_filter_error_bound(layers);
    p.status = status; p.out_scores = out_scores; p.out_pos = out_pos; p.out_ids = out_ids;

    bool handled = false;
    rc = fmap_warp_filter_try(Q, nq, cand, X, N, idx, C, D, n_qubits, layers, approx, st, &handled);
    if (rc) return rc;
    QRAG_REQUIRE(handled, QRAG_ERR_UNSUPPORTED, "layers=%d too deep for the register kernel", layers);
    const size_t smem_sel = (size_t)next_pow2(C) * 8;
    fr_select_kernel<<<nq, 256, smem_sel, st>>>(p);
    QRAG_LAUNCH_CHECK("fr_select_kernel");
    // certify: the listed candidates as gathered rows of the flattened candidate array (dense) or of the corpus
    const float* rows = cand ? cand : X;
    const int64_t nrows = cand ? (int64_t)nq * C : N;
    handled = false;
    rc = fmap_warp_try(Q, nq, nullptr, rows, nrows, p.gather, cap, D, n_qubits, layers, p.exact, nullptr, st, &handled);
    if (rc) return rc;
    QRAG_REQUIRE(handled, QRAG_ERR_UNSUPPORTED, "layers=%d too deep for the register kernel", layers);
    const size_t smem_fin = (size_t)next_pow2(cap) * 12;
    fr_final_kernel<<<nq, 256, smem_fin, st>>>(p);
    QRAG_LAUNCH_CHECK("fr_final_kernel");
    return QRAG_OK;
}
