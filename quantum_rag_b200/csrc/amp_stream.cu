// K2 (streaming form): amplitude-encoded fidelity with the candidate rows staged through a
// shared-memory ring by the TMA bulk-copy engine.
//
//   one persistent CTA per SM = 1 producer warp + 16 consumer warps
//   producer : cp.async.bulk (SASS UBLKCP) of whole tiles (R candidate rows, plus the
//              query row when the query changes) into a ring of S stages, completion
//              signalled on mbarriers; up to S*R*D*4 bytes (~200 KB) in flight per SM
//   consumers: wait on the stage's "full" barrier, read rows with conflict-free LDS.128,
//              accumulate q.d and |d|^2 in fp64 against the query state cached in
//              registers, transposed warp-shuffle reduction, release the stage
//   fused    : when a query's last tile is done its C scores sit in shared memory and are
//              ranked there ((score desc, position asc), quantum.py:70-76); only top_k
//              (score, position) pairs leave the SM.
//
// HBM traffic = every candidate row exactly once; nothing else of size touches DRAM.
// Per-row numerics follow amp_fidelity.cu (same lane->element map and reduction tree); every
// score of a query is computed the same way in fused and unfused mode, so the two agree bit
// for bit and duplicate candidates tie exactly.
#include "common.cuh"
#include "sort.cuh"
#include "tma.cuh"

namespace qrag {

constexpr int AS_CWARPS = 16;
constexpr int AS_CONSUMERS = AS_CWARPS * 32;
constexpr int AS_THREADS = AS_CONSUMERS + 32;
constexpr int AS_MAX_STAGES = 8;
constexpr int AS_BAR = 1;                 // named barrier id for the consumer warps
constexpr int AS_RANK_SORT_MAX = 128;

struct AmpStreamParams {
    const float* Q; const float* cand; const float* X; const int64_t* idx;
    int nq; int64_t C; int D;
    int R, tpq, stages, fused;
    size_t stage_bytes;
    double* out64; float* out32;
    int top_k; double* out_scores; int32_t* out_pos; int64_t* out_ids;
};

// Transposed butterfly over the 2*RB per-lane partials v[2*r] = q.d of row r, v[2*r+1] = |d|^2 of
// row r.  Halving steps exchange half of the values with the partner lane, so 2*RB values
// cost (2*RB - 1) + (5 - log2(2*RB)) double shuffles instead of 5 * 2 * RB.  Afterwards every lane
// holds the complete sum of value index (lane >> (5 - log2(2*RB))).  Every value goes through the
// same balanced tree (fp add commutes), so a row's result does not depend on its slot.
template <int NV>
__device__ __forceinline__ double reduce_vals(const double (&v)[NV], int lane) {
    static_assert(NV == 2 || NV == 4 || NV == 8, "2, 4 or 8 values");
    double w[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) w[i] = v[i];
    int off = 16;
#pragma unroll
    for (int n = NV; n > 1; n >>= 1, off >>= 1) {
        const bool up = lane & off;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const double send = up ? w[i] : w[i + n / 2];
            const double keep = up ? w[i + n / 2] : w[i];
            w[i] = keep + __shfl_xor_sync(FULL_MASK, send, off);
        }
    }
    double r = w[0];
    for (; off > 0; off >>= 1) r += __shfl_xor_sync(FULL_MASK, r, off);
    return r;
}

__device__ __forceinline__ void fma4s(const float4& d, const double* q, double& dot, double& nrm) {
    const double d0 = (double)d.x, d1 = (double)d.y, d2 = (double)d.z, d3 = (double)d.w;
    dot = fma(q[0], d0, dot); nrm = fma(d0, d0, nrm);
    dot = fma(q[1], d1, dot); nrm = fma(d1, d1, nrm);
    dot = fma(q[2], d2, dot); nrm = fma(d2, d2, nrm);
    dot = fma(q[3], d3, dot); nrm = fma(d3, d3, nrm);
}

// NCHUNK > 0: D == 128 * NCHUNK, query cached in registers.  NCHUNK == 0: any D % 4 == 0,
// query converted once per query into shared memory.  RB = rows per consumer warp per tile
// (tile height R = 16 * RB).
template <int NCHUNK, int RB>
__global__ void __launch_bounds__(AS_THREADS, 1) amp_stream_kernel(const AmpStreamParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + AS_MAX_STAGES;
    unsigned char* ring = smem + 128;
    const int D = p.D, R = p.R, S = p.stages;
    const int64_t C = p.C;
    double* qd = reinterpret_cast<double*>(ring + (size_t)S * p.stage_bytes);          // [D] (NCHUNK == 0 only)
    double* sc = qd + (NCHUNK == 0 ? ((D + 1) & ~1) : 0);                               // [P] fused scores / keys
    int P = 1;
    while (P < C) P <<= 1;
    int* stag = reinterpret_cast<int*>(sc + P);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], AS_CWARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    // tile range of this CTA
    const int tpq = p.tpq;
    int64_t g0, g1;
    if (p.fused) {
        g0 = ((int64_t)blockIdx.x * p.nq / gridDim.x) * tpq;
        g1 = ((int64_t)(blockIdx.x + 1) * p.nq / gridDim.x) * tpq;
    } else {
        const int64_t T = (int64_t)p.nq * tpq;
        g0 = (int64_t)blockIdx.x * T / gridDim.x;
        g1 = (int64_t)(blockIdx.x + 1) * T / gridDim.x;
    }
    const uint32_t row_bytes = (uint32_t)D * 4u;

    if (warp == AS_CWARPS) {
        // ------------------------------------------------------------------ producer
        int64_t q = g0 / tpq;
        int ti = (int)(g0 - q * tpq);
        int s = 0;
        uint32_t par = 0;
        for (int64_t g = g0; g < g1; ++g) {
            const int r0 = ti * R;
            const int nr = (int)((C - r0) < R ? (C - r0) : R);
            const bool newq = (g == g0) || (ti == 0);
            unsigned char* st = ring + (size_t)s * p.stage_bytes;
            mbar_wait(&empty[s], par ^ 1u);
            if (p.cand) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full[s], (uint32_t)nr * row_bytes + (newq ? row_bytes : 0u));
                    if (newq) bulk_g2s(st + (size_t)R * row_bytes, p.Q + (size_t)q * D, row_bytes, &full[s]);
                    bulk_g2s(st, p.cand + ((size_t)q * C + r0) * D, (uint32_t)nr * row_bytes, &full[s]);
                }
            } else {
                int64_t id = -1;
                if (lane < nr) id = p.idx[(size_t)q * C + r0 + lane];
                const unsigned valid = __ballot_sync(FULL_MASK, id >= 0);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full[s], (uint32_t)__popc(valid) * row_bytes + (newq ? row_bytes : 0u));
                    if (newq) bulk_g2s(st + (size_t)R * row_bytes, p.Q + (size_t)q * D, row_bytes, &full[s]);
                }
                __syncwarp();
                if (id >= 0) bulk_g2s(st + (size_t)lane * row_bytes, p.X + (size_t)id * D, row_bytes, &full[s]);
            }
            if (++ti == tpq) { ti = 0; ++q; }
            if (++s == S) { s = 0; par ^= 1u; }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    double qreg[NCHUNK > 0 ? NCHUNK * 4 : 1];
    double nq2 = 0.0;
    const int D4 = D >> 2;
    int64_t q = g0 / tpq;
    int ti = (int)(g0 - q * tpq);
    int s = 0;
    uint32_t par = 0;
    for (int64_t g = g0; g < g1; ++g) {
        const int r0 = ti * R;
        const int nr = (int)((C - r0) < R ? (C - r0) : R);
        const bool newq = (g == g0) || (ti == 0);
        const unsigned char* st = ring + (size_t)s * p.stage_bytes;
        mbar_wait(&full[s], par);

        if (newq) {
            const float* qslot = reinterpret_cast<const float*>(st + (size_t)R * row_bytes);
            double part = 0.0;
            if (NCHUNK > 0) {
#pragma unroll
                for (int t = 0; t < (NCHUNK > 0 ? NCHUNK : 0); ++t) {
                    const float4 v = reinterpret_cast<const float4*>(qslot)[lane + 32 * t];
                    qreg[4 * t + 0] = (double)v.x; qreg[4 * t + 1] = (double)v.y;
                    qreg[4 * t + 2] = (double)v.z; qreg[4 * t + 3] = (double)v.w;
#pragma unroll
                    for (int e = 0; e < 4; ++e) part = fma(qreg[4 * t + e], qreg[4 * t + e], part);
                }
            } else {
                named_bar_sync(AS_BAR, AS_CONSUMERS);          // everyone is done with the previous query's qd
                for (int i = tid; i < D; i += AS_CONSUMERS) qd[i] = (double)qslot[i];
                named_bar_sync(AS_BAR, AS_CONSUMERS);
                for (int i = lane; i < D; i += 32) part = fma(qd[i], qd[i], part);
            }
            nq2 = warp_sum(part);
        }

        // rows of this tile owned by the warp: slot i -> row (warp + ti) % 16 + 16 i
        const int rbase = (warp + ti) & (AS_CWARPS - 1);
        double acc[2 * RB];
#pragma unroll
        for (int i = 0; i < 2 * RB; ++i) acc[i] = 0.0;
        const float4* rp[RB];
#pragma unroll
        for (int i = 0; i < RB; ++i) {
            const int rr = rbase + AS_CWARPS * i;
            rp[i] = reinterpret_cast<const float4*>(st + (size_t)(rr < nr ? rr : 0) * row_bytes);
        }
        if (NCHUNK > 0) {
#pragma unroll
            for (int t = 0; t < (NCHUNK > 0 ? NCHUNK : 0); ++t)
#pragma unroll
                for (int i = 0; i < RB; ++i) fma4s(rp[i][lane + 32 * t], &qreg[4 * t], acc[2 * i], acc[2 * i + 1]);
        } else {
            for (int j = lane; j < D4; j += 32) {
                const double2 qa = *reinterpret_cast<const double2*>(qd + 4 * j);
                const double2 qb = *reinterpret_cast<const double2*>(qd + 4 * j + 2);
                const double qv[4] = {qa.x, qa.y, qb.x, qb.y};
#pragma unroll
                for (int i = 0; i < RB; ++i) fma4s(rp[i][j], qv, acc[2 * i], acc[2 * i + 1]);
            }
        }
        // lane L ends with value index L / LPV: even index = q.d, odd = |d|^2 of row slot index / 2
        constexpr int LPV = 32 / (2 * RB);
        const double tot = reduce_vals<2 * RB>(acc, lane);
        const double nd2 = __shfl_down_sync(FULL_MASK, tot, LPV);
        // all shared-memory reads of this stage are complete (their values were consumed above)
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);

        if ((lane & (2 * LPV - 1)) == 0) {
            const int i = lane / (2 * LPV);
            const int rr = rbase + AS_CWARPS * i;
            if (rr < nr) {
                const int64_t c = r0 + rr;
                const double den = nq2 * nd2;
                double f = den > 0.0 ? (tot * tot) / den : 0.0;
                if (p.idx && p.idx[(size_t)q * C + c] < 0) f = -pos_inf();
                if (p.fused) {
                    sc[c] = f;
                } else {
                    p.out64[(size_t)q * C + c] = f;
                    if (p.out32) p.out32[(size_t)q * C + c] = (float)f;
                }
            }
        }

        if (p.fused && ti == tpq - 1) {
            named_bar_sync(AS_BAR, AS_CONSUMERS);              // all C scores of query q are in sc[]
            const int top_k = p.top_k;
            double* os = p.out_scores + (size_t)q * top_k;
            int32_t* op = p.out_pos + (size_t)q * top_k;
            int64_t* oi = p.out_ids ? p.out_ids + (size_t)q * top_k : nullptr;
            if (C <= AS_RANK_SORT_MAX) {
                if (tid < C) {
                    const double si = sc[tid];
                    int rank = 0;
                    for (int j = 0; j < (int)C; ++j) {
                        const double sj = sc[j];
                        rank += (sj > si) || (sj == si && j < tid);
                    }
                    if (rank < top_k) {
                        os[rank] = si;
                        op[rank] = tid;
                        if (oi) oi[rank] = p.idx[(size_t)q * C + tid];
                    }
                }
            } else {
                for (int i = tid; i < P; i += AS_CONSUMERS) {
                    const bool real = i < C;
                    const double v = real ? sc[i] : 0.0;
                    sc[i] = real ? -v : pos_inf();
                    stag[i] = real ? i : TagPad<int>::value();
                }
                named_bar_sync(AS_BAR, AS_CONSUMERS);
                block_bitonic_sort<int>(sc, stag, P, tid, AS_CONSUMERS, AS_BAR);
                for (int i = tid; i < top_k; i += AS_CONSUMERS) {
                    os[i] = -sc[i];
                    op[i] = stag[i];
                    if (oi) oi[i] = p.idx[(size_t)q * C + stag[i]];
                }
            }
            named_bar_sync(AS_BAR, AS_CONSUMERS);              // sc[] is free for the next query
        }
        if (++ti == tpq) { ti = 0; ++q; }
        if (++s == S) { s = 0; par ^= 1u; }
    }
}

template <int NCHUNK, int RB>
static int launch_stream(const AmpStreamParams& p, size_t smem_bytes, int grid, cudaStream_t st) {
    auto kern = amp_stream_kernel<NCHUNK, RB>;
    QRAG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    kern<<<grid, AS_THREADS, smem_bytes, st>>>(p);
    QRAG_LAUNCH_CHECK("amp_stream_kernel");
    return QRAG_OK;
}

// Returns QRAG_OK and sets *handled = true if the streaming kernel took the job.
int amp_stream_try(const float* Q, int nq, const float* cand, const float* X, const int64_t* idx, int64_t C, int D,
                   bool fused, double* out64, float* out32, int top_k, double* out_scores, int32_t* out_pos,
                   int64_t* out_ids, cudaStream_t st, bool* handled) {
    *handled = false;
    const DeviceProps& dp = device_props();
    QRAG_REQUIRE(dp.ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    if (D % 4 != 0 || D > 4096) return QRAG_OK;
    if (((uintptr_t)Q | (uintptr_t)(cand ? cand : X)) % 16 != 0) return QRAG_OK;
    if (nq < 1 || C < 1) return QRAG_OK;

    AmpStreamParams p{};
    p.Q = Q; p.cand = cand; p.X = X; p.idx = idx; p.nq = nq; p.C = C; p.D = D;
    p.out64 = out64; p.out32 = out32; p.top_k = top_k; p.out_scores = out_scores; p.out_pos = out_pos;
    p.out_ids = out_ids; p.fused = fused ? 1 : 0;
    // tile height: 32 rows while a stage stays <= 64 KB, else 16
    int rb = 2;
    while (rb > 1 && (size_t)(AS_CWARPS * rb + 1) * D * 4 > 64 * 1024) rb >>= 1;
    p.R = AS_CWARPS * rb;
    p.tpq = (int)ceil_div(C, p.R);
    p.stage_bytes = ((size_t)(p.R + 1) * D * 4 + 127) / 128 * 128;
    const bool nchunk_path = (D == 128 || D == 256 || D == 384 || D == 512);
    size_t fixed = 128 + (nchunk_path ? 0 : (size_t)((D + 1) & ~1) * 8);
    if (fused) fixed += (size_t)next_pow2(C) * 12;
    const size_t budget = (size_t)dp.max_smem_optin;
    if (fixed + 2 * p.stage_bytes > budget) return QRAG_OK;
    int stages = (int)((budget - fixed) / p.stage_bytes);
    if (stages > AS_MAX_STAGES) stages = AS_MAX_STAGES;
    p.stages = stages;
    const size_t smem_bytes = fixed + (size_t)stages * p.stage_bytes;

    int grid = dp.sm_count;
    if (fused) {
        if (grid > nq) grid = nq;
    } else {
        const int64_t T = (int64_t)nq * p.tpq;
        if (grid > T) grid = (int)T;
    }
    *handled = true;
    if (nchunk_path) {
        switch (D / 128) {
            case 1: return launch_stream<1, 2>(p, smem_bytes, grid, st);
            case 2: return launch_stream<2, 2>(p, smem_bytes, grid, st);
            case 3: return launch_stream<3, 2>(p, smem_bytes, grid, st);
            default: return launch_stream<4, 2>(p, smem_bytes, grid, st);
        }
    }
    if (rb == 2) return launch_stream<0, 2>(p, smem_bytes, grid, st);
    return launch_stream<0, 1>(p, smem_bytes, grid, st);
}

}  // namespace qrag
