// K2: amplitude-encoded state fidelity, batched, with an optional fused
// stable sort + top-k (the tensor-level QuantumReranker.rerank, quantum.py:44-78).
//
// For amplitude encoding the 2^n-amplitude state of a vector x is zero-pad(x)/|x|,
// so |<q^|d^>|^2 = (q.d)^2 / (|q|^2 |d|^2).  The kernel streams every candidate row
// from HBM exactly once (4*D bytes per score), keeps the query state staged in
// shared memory as fp64, accumulates the overlap in fp64 and reduces with warp
// shuffles.  HBM-bound: algorithmic bytes per score = 4*D + 4*D/C + 8.
//
// Work decomposition: one CTA owns one query at a time (persistent over queries),
// its 8 warps take candidate rows in batches of 4 so that 4 independent rows (up to
// 12 x 16 B per lane) are in flight per warp.  The per-row reduction tree is fixed
// (same lane->element map, same butterfly), so duplicate candidates produce
// bit-identical scores and tie exactly, as the reference's stable sort needs.
#include "common.cuh"
#include "sort.cuh"

namespace qrag {

constexpr int AMP_THREADS = 256;
constexpr int AMP_WARPS = AMP_THREADS / 32;
constexpr int AMP_ROWS = 4;          // rows per warp batch
constexpr int AMP_RANK_SORT_MAX = 128;

struct AmpParams {
    const float* Q;        // [nq, D]
    const float* cand;     // [nq, C, D] or null
    const float* X;        // [N, D] or null
    int64_t N;             // rows of X: an idx outside [0, N) is padding (scores -inf), never an out-of-bounds read
    const int64_t* idx;    // [nq, C] or null
    int nq;
    int64_t C;
    int D;
    double* out64;         // [nq, C]       (unfused)
    float* out32;          // [nq, C]       (unfused, optional)
    int top_k;             // fused
    double* out_scores;    // [nq, top_k]   (fused)
    int32_t* out_pos;      // [nq, top_k]   (fused)
    int64_t* out_ids;      // [nq, top_k]   (fused, optional)
};

__device__ __forceinline__ double fidelity_from(double dot, double nd2, double nq2) {
    const double den = nq2 * nd2;
    return den > 0.0 ? (dot * dot) / den : 0.0;
}

// Transposed butterfly: 8 per-lane partials (v[2*r+0] = dot of row r, v[2*r+1] = |d|^2
// of row r) -> lane L (L & 3 == 0) ends with the full sum of value index (L >> 2) & 7.
// 9 double shuffles instead of 40.  Every value goes through the same balanced tree
// (fp add is commutative), so the result does not depend on the slot a row sits in.
__device__ __forceinline__ double reduce8(const double (&v)[8], int lane) {
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
    {
        const double send = b2 ? w2[0] : w2[1];
        const double keep = b2 ? w2[1] : w2[0];
        w1 = keep + __shfl_xor_sync(FULL_MASK, send, 4);
    }
    w1 += __shfl_xor_sync(FULL_MASK, w1, 2);
    w1 += __shfl_xor_sync(FULL_MASK, w1, 1);
    return w1;   // value index = (b4 << 2) | (b3 << 1) | b2
}

__device__ __forceinline__ void fma4(const float4& d, const double* q, double& dot, double& nrm) {
    const double d0 = (double)d.x, d1 = (double)d.y, d2 = (double)d.z, d3 = (double)d.w;
    dot = fma(q[0], d0, dot); nrm = fma(d0, d0, nrm);
    dot = fma(q[1], d1, dot); nrm = fma(d1, d1, nrm);
    dot = fma(q[2], d2, dot); nrm = fma(d2, d2, nrm);
    dot = fma(q[3], d3, dot); nrm = fma(d3, d3, nrm);
}

// NCHUNK > 0: D == NCHUNK * 128 exactly, query slice cached in registers.
// NCHUNK == 0: any D % 4 == 0 (16-byte aligned rows), query read from shared memory.
// NCHUNK == -1: scalar path for any D / alignment.
template <int NCHUNK, bool FUSED>
__global__ void __launch_bounds__(AMP_THREADS) amp_fidelity_kernel(const AmpParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* qs = reinterpret_cast<double*>(smem_raw);                 // [D] query as fp64
    const int D = p.D;
    const int Dpad = (D + 3) & ~3;
    double* red = qs + Dpad;                                          // [AMP_WARPS] block reduce scratch
    double* skey = red + AMP_WARPS;                                   // [P] fused sort keys / scores
    const int64_t C = p.C;
    int* stag = nullptr;
    int P = 0;
    if (FUSED) {
        P = 1;
        while (P < C) P <<= 1;
        stag = reinterpret_cast<int*>(skey + P);
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D4 = D >> 2;

    for (int q = blockIdx.x; q < p.nq; q += gridDim.x) {
        // ---- stage the query state (unnormalised amplitudes) as fp64, and |q|^2 ----
        const float* qrow = p.Q + (size_t)q * D;
        double part = 0.0;
        for (int i = tid; i < D; i += AMP_THREADS) {
            const double v = (double)qrow[i];
            qs[i] = v;
            part = fma(v, v, part);
        }
        part = warp_sum(part);
        if (lane == 0) red[warp] = part;
        __syncthreads();
        double nq2 = 0.0;
#pragma unroll
        for (int w = 0; w < AMP_WARPS; ++w) nq2 += red[w];

        double qreg[NCHUNK > 0 ? NCHUNK * 4 : 1];
        if (NCHUNK > 0) {
#pragma unroll
            for (int t = 0; t < (NCHUNK > 0 ? NCHUNK : 0); ++t)
#pragma unroll
                for (int e = 0; e < 4; ++e) qreg[t * 4 + e] = qs[(lane + 32 * t) * 4 + e];
        }

        // ---- stream candidate rows: AMP_ROWS rows per warp per iteration ----
        for (int64_t r0 = (int64_t)warp * AMP_ROWS; r0 < C; r0 += AMP_WARPS * AMP_ROWS) {
            const float* rp[AMP_ROWS];
            unsigned missing = 0;   // bit i: row i of the batch has idx < 0 (padding)
#pragma unroll
            for (int i = 0; i < AMP_ROWS; ++i) {
                int64_t r = r0 + i;
                if (r >= C) r = r0;                     // clamp: result discarded below
                if (p.cand) {
                    rp[i] = p.cand + ((size_t)q * C + r) * D;
                } else {
                    int64_t id = p.idx[(size_t)q * C + r];
                    const bool pad = id < 0 || id >= p.N;
                    missing |= (pad ? 1u : 0u) << i;
                    rp[i] = p.X + (size_t)(pad ? 0 : id) * D;
                }
            }
            double acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.0;

            if (NCHUNK > 0) {
                float4 v[AMP_ROWS][NCHUNK > 0 ? NCHUNK : 1];
#pragma unroll
                for (int i = 0; i < AMP_ROWS; ++i)
#pragma unroll
                    for (int t = 0; t < (NCHUNK > 0 ? NCHUNK : 0); ++t)
                        v[i][t] = ldg_stream(reinterpret_cast<const float4*>(rp[i]) + lane + 32 * t);
#pragma unroll
                for (int t = 0; t < (NCHUNK > 0 ? NCHUNK : 0); ++t)
#pragma unroll
                    for (int i = 0; i < AMP_ROWS; ++i) fma4(v[i][t], &qreg[t * 4], acc[2 * i], acc[2 * i + 1]);
            } else if (NCHUNK == 0) {
#pragma unroll 2
                for (int j = lane; j < D4; j += 32) {
                    float4 v[AMP_ROWS];
#pragma unroll
                    for (int i = 0; i < AMP_ROWS; ++i) v[i] = ldg_stream(reinterpret_cast<const float4*>(rp[i]) + j);
                    const double2 qa = *reinterpret_cast<const double2*>(qs + 4 * j);
                    const double2 qb = *reinterpret_cast<const double2*>(qs + 4 * j + 2);
                    const double qv[4] = {qa.x, qa.y, qb.x, qb.y};
#pragma unroll
                    for (int i = 0; i < AMP_ROWS; ++i) fma4(v[i], qv, acc[2 * i], acc[2 * i + 1]);
                }
            } else {
                for (int j = lane; j < D; j += 32) {
                    const double qv = qs[j];
#pragma unroll
                    for (int i = 0; i < AMP_ROWS; ++i) {
                        const double d = (double)__ldg(rp[i] + j);
                        acc[2 * i] = fma(qv, d, acc[2 * i]);
                        acc[2 * i + 1] = fma(d, d, acc[2 * i + 1]);
                    }
                }
            }

            const double tot = reduce8(acc, lane);
            const double nd2 = __shfl_down_sync(FULL_MASK, tot, 4);   // |d|^2 sits 4 lanes up
            if ((lane & 7) == 0) {
                const int i = lane >> 3;
                const int64_t r = r0 + i;
                if (r < C) {
                    double f = fidelity_from(tot, nd2, nq2);
                    if ((missing >> i) & 1u) f = -pos_inf();
                    if (FUSED) {
                        skey[r] = f;
                    } else {
                        p.out64[(size_t)q * C + r] = f;
                        if (p.out32) p.out32[(size_t)q * C + r] = (float)f;
                    }
                }
            }
        }

        if (FUSED) {
            __syncthreads();
            const int top_k = p.top_k;
            double* os = p.out_scores + (size_t)q * top_k;
            int32_t* op = p.out_pos + (size_t)q * top_k;
            int64_t* oi = p.out_ids ? p.out_ids + (size_t)q * top_k : nullptr;
            if (C <= AMP_RANK_SORT_MAX) {
                // rank by counting: position of row i in (score desc, pos asc) order
                if (tid < C) {
                    const double si = skey[tid];
                    int rank = 0;
                    for (int j = 0; j < (int)C; ++j) {
                        const double sj = skey[j];
                        rank += (sj > si) || (sj == si && j < tid);
                    }
                    if (rank < top_k) {
                        os[rank] = si;
                        op[rank] = tid;
                        if (oi) oi[rank] = p.idx[(size_t)q * C + tid];
                    }
                }
            } else {
                for (int i = tid; i < P; i += AMP_THREADS) {
                    const bool real = i < C;
                    const double s = real ? skey[i] : 0.0;
                    skey[i] = real ? -s : pos_inf();
                    stag[i] = real ? i : TagPad<int>::value();
                }
                __syncthreads();
                block_bitonic_sort<int>(skey, stag, P);
                for (int i = tid; i < top_k; i += AMP_THREADS) {
                    os[i] = -skey[i];
                    op[i] = stag[i];
                    if (oi) oi[i] = p.idx[(size_t)q * C + stag[i]];
                }
            }
        }
        __syncthreads();   // qs / skey are reused by the next query
    }
}

template <int NCHUNK, bool FUSED>
static int launch_amp(const AmpParams& p, size_t smem, int grid, cudaStream_t st) {
    auto kern = amp_fidelity_kernel<NCHUNK, FUSED>;
    if (smem > 48 * 1024)
        QRAG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, AMP_THREADS, smem, st>>>(p);
    QRAG_LAUNCH_CHECK("amp_fidelity_kernel");
    return QRAG_OK;
}

template <bool FUSED>
static int dispatch_amp(const AmpParams& p, cudaStream_t st) {
    const DeviceProps& dp = device_props();
    QRAG_REQUIRE(dp.ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    const int Dpad = (p.D + 3) & ~3;
    size_t smem = (size_t)(Dpad + AMP_WARPS) * sizeof(double);
    if (FUSED) smem += (size_t)next_pow2(p.C) * (sizeof(double) + sizeof(int));
    QRAG_REQUIRE(smem <= (size_t)dp.max_smem_optin, QRAG_ERR_UNSUPPORTED,
                 "amp_fidelity: D=%d C=%lld needs %zu B of shared memory", p.D, (long long)p.C, smem);
    // one CTA per query, handed out by the hardware block scheduler; the in-kernel
    // query loop only engages for very large batches
    int grid = dp.sm_count * 64;
    if (grid > p.nq) grid = p.nq;

    const bool aligned = (p.D % 4 == 0) &&
                         (((uintptr_t)(p.cand ? p.cand : p.X)) % 16 == 0);
    if (!aligned) return launch_amp<-1, FUSED>(p, smem, grid, st);
    switch (p.D) {
        case 128: return launch_amp<1, FUSED>(p, smem, grid, st);
        case 256: return launch_amp<2, FUSED>(p, smem, grid, st);
        case 384: return launch_amp<3, FUSED>(p, smem, grid, st);
        case 512: return launch_amp<4, FUSED>(p, smem, grid, st);
        default:  return launch_amp<0, FUSED>(p, smem, grid, st);
    }
}

int amp_fidelity_plain(const AmpParams& p, cudaStream_t st) { return dispatch_amp<false>(p, st); }
int amp_fidelity_fused(const AmpParams& p, cudaStream_t st) { return dispatch_amp<true>(p, st); }

}  // namespace qrag

using namespace qrag;

static int check_amp_common(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx,
                            int64_t C, int D, int n_qubits) {
    QRAG_REQUIRE(Q != nullptr, QRAG_ERR_INVALID, "Q is null");
    QRAG_REQUIRE(nq >= 0 && C >= 0 && D > 0, QRAG_ERR_INVALID, "bad sizes nq=%d C=%lld D=%d", nq, (long long)C, D);
    QRAG_REQUIRE((cand != nullptr) != (X != nullptr && idx != nullptr), QRAG_ERR_INVALID,
                 "pass either cand, or X together with idx");
    QRAG_REQUIRE(cand != nullptr || N >= 1, QRAG_ERR_INVALID, "X needs its row count N >= 1 (got %lld)", (long long)N);
    QRAG_REQUIRE(n_qubits >= 1 && n_qubits <= 30, QRAG_ERR_INVALID, "n_qubits=%d out of range", n_qubits);
    QRAG_REQUIRE((int64_t)D <= ((int64_t)1 << n_qubits), QRAG_ERR_INVALID,
                 "D=%d does not fit %d qubits (2^n=%lld amplitudes)", D, n_qubits, (long long)1 << n_qubits);
    return QRAG_OK;
}

namespace qrag {
int amp_stream_try(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx, int64_t C, int D,
                   bool fused, double* out64, float* out32, int top_k, double* out_scores, int32_t* out_pos,
                   int64_t* out_ids, cudaStream_t st, bool* handled);                          // amp_stream.cu
int fmap_fidelity(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx, int64_t C, int D,
                  int n_qubits, int layers, double* out64, float* out32, cudaStream_t st);   // sv_kernels.cu
}

extern "C" int qrag_amp_fidelity(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx,
                                 int64_t C, int D, int n_qubits, int layers, double* out64, float* out32,
                                 void* stream) {
    int rc = check_amp_common(Q, nq, cand, X, N, idx, C, D, n_qubits);
    if (rc) return rc;
    QRAG_REQUIRE(out64 != nullptr, QRAG_ERR_INVALID, "out64 is null");
    QRAG_REQUIRE(layers >= 0, QRAG_ERR_INVALID, "layers=%d", layers);
    if (nq == 0 || C == 0) return QRAG_OK;
    if (layers > 0)
        return fmap_fidelity(Q, nq, cand, X, N, idx, C, D, n_qubits, layers, out64, out32, (cudaStream_t)stream);
    bool handled = false;
    rc = amp_stream_try(Q, nq, cand, X, N, idx, C, D, false, out64, out32, 0, nullptr, nullptr, nullptr,
                        (cudaStream_t)stream, &handled);
    if (rc || handled) return rc;
    AmpParams p{};           // rows that are not 16-byte aligned: plain load kernel
    p.Q = Q; p.cand = cand; p.X = X; p.N = N; p.idx = idx; p.nq = nq; p.C = C; p.D = D;
    p.out64 = out64; p.out32 = out32;
    return amp_fidelity_plain(p, (cudaStream_t)stream);
}

extern "C" int qrag_amp_rerank(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx,
                               int64_t C, int D, int n_qubits, int top_k, double* out_scores, int32_t* out_pos,
                               int64_t* out_ids, void* stream) {
    int rc = check_amp_common(Q, nq, cand, X, N, idx, C, D, n_qubits);
    if (rc) return rc;
    QRAG_REQUIRE(out_scores && out_pos, QRAG_ERR_INVALID, "out_scores/out_pos is null");
    QRAG_REQUIRE(C >= 1 && C <= QRAG_MAX_SORT_LEN, QRAG_ERR_UNSUPPORTED,
                 "fused rerank needs 1 <= C <= %d (got %lld)", QRAG_MAX_SORT_LEN, (long long)C);
    QRAG_REQUIRE(top_k >= 1 && top_k <= C, QRAG_ERR_INVALID, "top_k=%d outside [1, C=%lld]", top_k, (long long)C);
    QRAG_REQUIRE(out_ids == nullptr || idx != nullptr, QRAG_ERR_INVALID, "out_ids needs idx");
    if (nq == 0) return QRAG_OK;
    bool handled = false;
    rc = amp_stream_try(Q, nq, cand, X, N, idx, C, D, true, nullptr, nullptr, top_k, out_scores, out_pos, out_ids,
                        (cudaStream_t)stream, &handled);
    if (rc || handled) return rc;
    AmpParams p{};
    p.Q = Q; p.cand = cand; p.X = X; p.N = N; p.idx = idx; p.nq = nq; p.C = C; p.D = D;
    p.top_k = top_k; p.out_scores = out_scores; p.out_pos = out_pos; p.out_ids = out_ids;
    return amp_fidelity_fused(p, (cudaStream_t)stream);
}

// ---- the rerank called with HOST buffers: copies in, kernel, copies out, all enqueued by one call ----
namespace {
inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }
struct HostWs { size_t q, idx, s, o, p, total; };
HostWs host_ws(int nq, int64_t C, int D, int top_k) {
    HostWs w;
    w.q = 0;
    w.idx = w.q + up256((size_t)nq * D * sizeof(float));
    w.s = w.idx + up256((size_t)nq * C * sizeof(int64_t));
    w.o = w.s + up256((size_t)nq * top_k * sizeof(double));
    w.p = w.o + up256((size_t)nq * top_k * sizeof(int64_t));
    w.total = w.p + up256((size_t)nq * top_k * sizeof(int32_t));
    return w;
}
}  // namespace

extern "C" int qrag_amp_rerank_host_workspace(int nq, int64_t C, int D, int top_k, size_t* bytes) {
    QRAG_REQUIRE(bytes != nullptr, QRAG_ERR_INVALID, "bytes is null");
    QRAG_REQUIRE(nq >= 0 && C >= 1 && D >= 1 && top_k >= 1, QRAG_ERR_INVALID, "bad sizes nq=%d C=%lld D=%d top_k=%d", nq,
                 (long long)C, D, top_k);
    *bytes = host_ws(nq, C, D, top_k).total;
    return QRAG_OK;
}

extern "C" int qrag_amp_rerank_host(const float* hQ, int nq, const int64_t* hIdx, int64_t C, const float* X, int64_t N,
                                    int D, int n_qubits, int top_k, void* workspace, size_t workspace_bytes,
                                    double* hScores, int64_t* hIds, void* stream) {
    QRAG_REQUIRE(hQ && hIdx && X && workspace && hScores && hIds, QRAG_ERR_INVALID, "null pointer argument");
    QRAG_REQUIRE(nq >= 0 && C >= 1 && D >= 1 && top_k >= 1, QRAG_ERR_INVALID, "bad sizes nq=%d C=%lld D=%d top_k=%d", nq,
                 (long long)C, D, top_k);
    QRAG_REQUIRE(((uintptr_t)workspace & 255) == 0, QRAG_ERR_INVALID, "workspace must be 256-byte aligned");
    const HostWs w = host_ws(nq, C, D, top_k);
    QRAG_REQUIRE(workspace_bytes >= w.total, QRAG_ERR_INVALID, "workspace of %zu bytes, %zu needed", workspace_bytes, w.total);
    if (nq == 0) return QRAG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = static_cast<char*>(workspace);
    float* dQ = reinterpret_cast<float*>(base + w.q);
    int64_t* dI = reinterpret_cast<int64_t*>(base + w.idx);
    double* dS = reinterpret_cast<double*>(base + w.s);
    int64_t* dO = reinterpret_cast<int64_t*>(base + w.o);
    int32_t* dP = reinterpret_cast<int32_t*>(base + w.p);
    QRAG_CUDA_CHECK(cudaMemcpyAsync(dQ, hQ, (size_t)nq * D * sizeof(float), cudaMemcpyHostToDevice, st));
    QRAG_CUDA_CHECK(cudaMemcpyAsync(dI, hIdx, (size_t)nq * C * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    const int rc = qrag_amp_rerank(dQ, nq, nullptr, X, N, dI, C, D, n_qubits, top_k, dS, dP, dO, stream);
    if (rc) return rc;
    QRAG_CUDA_CHECK(cudaMemcpyAsync(hScores, dS, (size_t)nq * top_k * sizeof(double), cudaMemcpyDeviceToHost, st));
    QRAG_CUDA_CHECK(cudaMemcpyAsync(hIds, dO, (size_t)nq * top_k * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    return QRAG_OK;
}
