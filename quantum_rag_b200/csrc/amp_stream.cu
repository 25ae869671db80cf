// K2 (streaming form): amplitude-encoded fidelity with the candidate rows staged through a
// shared-memory ring by the TMA bulk-copy engine, and the per-query ranking taken off the
// streaming path.
//
//   one persistent CTA per SM, warp-specialised, no CTA-wide barrier after start-up:
//   producer  (1 warp) : cp.async.bulk (SASS UBLKCP) of tiles of RB candidate rows into a ring
//                        of S stages (~200 KB in flight per SM) and of each query row into one
//                        of QS query slots; completion lands on mbarriers as transaction bytes
//   converter (1 warp) : turns a query slot into fp64 once (and |q|^2), so that the consumers
//                        never convert the query
//   consumers (16 warps): warp w owns tiles w, w+16, ... of the CTA's tile sequence: waits for
//                        its stage, reads the RB rows with conflict-free LDS.128, accumulates
//                        q.d and |d|^2 in fp64 (transposed warp-shuffle reduction), releases the
//                        stage, and drops (q.d, |q|^2 |d|^2) into the query's score buffer
//   rankers   (2 warps) : fused mode only.  When all tiles of a query have reported, compute
//                        F = (q.d)^2 / (|q|^2 |d|^2) and rank the query's C scores in shared
//                        memory ((score desc, position asc), quantum.py:70-76); only top_k
//                        (score, position) pairs leave the SM.  Score buffers are double-buffered,
//                        so ranking query i overlaps streaming query i+1.
//
// HBM traffic = every candidate row exactly once; nothing else of size touches DRAM.
// Every score is computed by the same instruction sequence (same lane->element map, same
// reduction tree) whichever slot its row sits in, so duplicate candidates tie exactly and the
// fused and unfused forms agree bit for bit.
#include "common.cuh"
#include "sort.cuh"
#include "tma.cuh"

#include <stdlib.h>

namespace qrag {

// Two CTA shapes.  CW = 16 consumer warps (640 threads, the whole SM's shared memory, one CTA per SM), and the
// INTERLEAVED shape CW = 8 (384 threads, half the shared memory, two CTAs per SM) whose grid is still one CTA per SM:
// the second slot of every SM is taken by the NEXT launch on the stream (programmatic dependent launch), so
// consecutive batches run staggered by half a period and one batch's start-up and drain (first HBM round trip, last
// ranking: ~3 us of a 29 us batch) are covered by the other's streaming.  Only with QRAG_OVERLAP_INPUTS_STABLE
// (reads may start before the previous kernel completes); writes still wait for it.
constexpr int AS_CWARPS = 16;
constexpr int AS_CWARPS_HALF = 8;
constexpr int AS_RWARPS = 2;
constexpr int AS_RTHREADS = AS_RWARPS * 32;
__host__ __device__ constexpr int as_threads(int cw) { return (cw + AS_RWARPS + 2) * 32; }      // consumers, rankers, converter, producer
constexpr int AS_MAX_STAGES = 64;
constexpr int AS_MAX_QSLOTS = 8;
constexpr int AS_MAX_SBUFS = 8;
constexpr int AS_BAR_RANK = 1;                           // named barrier of the ranker warps
constexpr int AS_RANK_COUNT_MAX = 2 * AS_RTHREADS;       // rank-by-counting up to this many candidates
constexpr int AS_BAR_BYTES = (2 * AS_MAX_STAGES + 2 * AS_MAX_QSLOTS + 2 * AS_MAX_SBUFS) * 8;   // 1280
constexpr int AS_HDR_BYTES = 1280;

struct AmpStreamParams {
    const float* Q; const float* cand; const float* X; const int64_t* idx;
    int64_t N;          // rows of X: an idx outside [0, N) is padding
    int nq; int64_t C; int D;
    int G;              // consumer warps per team; a team drains one tile of R = G * RB rows together
    int R;              // rows per tile (one bulk copy, one pair of barriers)
    int tpq;            // tiles per query = ceil(C / R)
    int stages;         // S (a multiple of teams)
    int teams;          // teams that take tiles = min(16 / G, S); tile t belongs to team t % teams
    int qslots;         // QS
    int nbuf;           // score buffers (fused): rankers may lag the stream by nbuf - 1 queries
    int fused;
    int P;              // next_pow2(C) (fused)
    uint32_t stage_bytes, qslot_bytes;
    double* out64; float* out32;
    int top_k; double* out_scores; int32_t* out_pos; int64_t* out_ids;
    int overlap;        // QRAG_OVERLAP_*: where the kernel orders itself after the previous kernel on the stream
    int out_stage_q;    // interleaved shape: results of up to this many queries per CTA are staged in shared memory and
                        // written after the previous kernel has completed (0 = write each query's result at once)
};

// Transposed butterfly over the NV = 2*RB per-lane partials v[2*r] = q.d of row r, v[2*r+1] = |d|^2
// of row r.  Halving steps exchange half of the values with the partner lane, so NV values cost
// (NV - 1) + (5 - log2 NV) double shuffles instead of 5 * NV.  Afterwards every lane holds the
// complete sum of value index lane / (32 / NV).  Every value goes through the same balanced
// tree (fp add commutes), so a row's result does not depend on its slot.
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

// score from the pair the consumers leave behind: den < 0 marks a padding candidate (idx < 0)
__device__ __forceinline__ double fidelity_pair(double dot, double den) {
    if (den < 0.0) return -pos_inf();
    return den > 0.0 ? (dot * dot) / den : 0.0;
}

// bitonic sort of P (power of two) {key, tag-as-int64-bits} pairs laid out as double2, ascending
// by (key, tag); nt threads on named barrier `bar`
__device__ void bitonic_sort_aos(double2* a, int P, int tid, int nt, int bar) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += nt) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int hi = lo | j;
                const double2 l = a[lo], h = a[hi];
                const bool ascending = (lo & k) == 0;
                const bool hi_first = pair_less<long long>(h.x, __double_as_longlong(h.y), l.x, __double_as_longlong(l.y));
                if (hi_first == ascending) { a[lo] = h; a[hi] = l; }
            }
            named_bar_sync(bar, nt);
        }
    }
}

// NCHUNK > 0: D == 128 * NCHUNK, query cached in registers.  NCHUNK == 0: any D % 4 == 0, query read
// from its fp64 slot in shared memory.  RB = candidate rows per tile (one consumer warp per tile).
template <int NCHUNK, int RB, int CW>
__global__ void __launch_bounds__(as_threads(CW), CW == AS_CWARPS ? 1 : 2) amp_stream_kernel(const AmpStreamParams p) {
    constexpr int AS_WARP_RANK = CW;                         // first ranker warp
    constexpr int AS_WARP_CONV = CW + AS_RWARPS;
    constexpr int AS_WARP_PROD = AS_WARP_CONV + 1;
    constexpr int AS_THREADS = as_threads(CW);
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + AS_MAX_STAGES;
    uint64_t* qfull = empty + AS_MAX_STAGES;
    uint64_t* qready = qfull + AS_MAX_QSLOTS;
    uint64_t* sfull = qready + AS_MAX_QSLOTS;
    uint64_t* sempty = sfull + AS_MAX_SBUFS;
    const int D = p.D, S = p.stages, QS = p.qslots, tpq = p.tpq, P = p.P, NT = p.teams, NB = p.nbuf, G = p.G, R = p.R;
    const int64_t C = p.C;
    const uint32_t row_bytes = (uint32_t)D * 4u;
    // query slot: [D] fp32 (TMA destination) | [D] fp64 | |q|^2
    unsigned char* qslot0 = smem + AS_HDR_BYTES;
    const uint32_t qd_off = (row_bytes + 15u) & ~15u;
    const uint32_t qn_off = qd_off + (uint32_t)D * 8u;
    double2* sc = reinterpret_cast<double2*>(qslot0 + (size_t)QS * p.qslot_bytes);     // [NB][P] (fused)
    unsigned char* ring = reinterpret_cast<unsigned char*>(sc + (p.fused ? (size_t)NB * P : 0));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Start-up: the PRODUCER warp initialises the barriers (one per lane, not one thread in a loop), makes them visible
    // and goes straight to issuing the first bulk copies; it only ARRIVES on the start barrier, the other warps wait on
    // it.  The first HBM round trip therefore overlaps the rest of the CTA's prologue instead of following it.
    constexpr int AS_BAR_START = 2;
    if (warp == AS_WARP_PROD) {
        for (int s = lane; s < S; s += 32) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], (uint32_t)G);
        }
        for (int s = lane; s < QS; s += 32) {
            mbar_init(&qfull[s], 1);
            mbar_init(&qready[s], 1);
        }
        for (int s = lane; s < NB; s += 32) {
            mbar_init(&sfull[s], p.fused ? (uint32_t)(tpq * G) : 1u);
            mbar_init(&sempty[s], AS_RWARPS);
        }
        fence_barrier_init();
        __syncwarp();
        asm volatile("bar.arrive %0, %1;" ::"r"(AS_BAR_START), "r"(AS_THREADS) : "memory");
    } else {
        named_bar_sync(AS_BAR_START, AS_THREADS);
    }
    // Programmatic dependent launch: let the next kernel on the stream start filling SMs as this
    // grid's CTAs retire.  QRAG_OVERLAP_SAFE orders every global access of this kernel after the
    // previous kernel (only launch latency and this prologue overlap); QRAG_OVERLAP_INPUTS_STABLE
    // orders only the writes, so streaming starts while the previous grid drains.
    // The interleaved shape triggers here too: two consecutive launches then start together, share every SM and end
    // together, so each pair pays ONE start-up and drain.  Delaying the trigger to the CTA's midpoint, which staggers
    // consecutive launches by half a batch, was measured and is worse (0.80-0.82 of the HBM peak against 0.88 in the same
    // run): a staggered launch cannot write its results until its predecessor completes half a batch later, and holds
    // its half of the SM idle meanwhile.
    if (p.overlap != QRAG_OVERLAP_NONE && tid == 0) griddep_launch_dependents();
    const bool wait_reads = p.overlap == QRAG_OVERLAP_SAFE;
    const bool wait_writes = p.overlap >= QRAG_OVERLAP_INPUTS_STABLE;

    // tile range [g0, g1) of this CTA in the global sequence (query-major, tpq tiles per query)
    int64_t g0, g1;
    if (p.fused) {
        g0 = ((int64_t)blockIdx.x * p.nq / gridDim.x) * tpq;
        g1 = ((int64_t)(blockIdx.x + 1) * p.nq / gridDim.x) * tpq;
    } else {
        const int64_t T = (int64_t)p.nq * tpq;
        g0 = (int64_t)blockIdx.x * T / gridDim.x;
        g1 = (int64_t)(blockIdx.x + 1) * T / gridDim.x;
    }
    const int ntiles = (int)(g1 - g0);
    if (ntiles <= 0) return;
    const int64_t qfirst = g0 / tpq;
    const int ti_first = (int)(g0 - qfirst * tpq);
    const int nqueries = (int)((g1 - 1) / tpq - qfirst) + 1;

    if (warp == AS_WARP_PROD) {
        // ---------------------------------------------------------------------- producer
        int consumed = 0, cst = 0;          // tiles [0, consumed) are known to be released
        uint32_t cpar = 0;
        auto ensure_consumed = [&](int upto) {
            while (consumed < upto) {
                mbar_wait(&empty[cst], cpar);
                ++consumed;
                if (++cst == S) { cst = 0; cpar ^= 1u; }
            }
        };
        int64_t q = qfirst;
        int ti = ti_first, st = 0, qs = 0, slot = 0;
        if (wait_reads) griddep_wait();
        for (int it = 0; it < ntiles; ++it) {
            if (it == 0 || ti == 0) {
                // the slot's previous query (qs - QS) must be fully consumed: all tiles before
                // the first tile of query qs - QS + 1
                if (qs >= QS) ensure_consumed((qs - QS + 1) * tpq - ti_first);
                // fused: the rankers must be done with query qs - NB before its score buffer is rewritten
                if (p.fused && qs >= NB) mbar_wait(&sempty[qs % NB], ((uint32_t)(qs / NB) & 1u) ^ 1u);
                if (lane == 0) {
                    mbar_arrive_expect_tx(&qfull[slot], row_bytes);
                    bulk_g2s(qslot0 + (size_t)slot * p.qslot_bytes, p.Q + (size_t)q * D, row_bytes, &qfull[slot]);
                }
            }
            if (it >= S) ensure_consumed(it - S + 1);
            const int r0 = ti * R;
            const int nr = (int)((C - r0) < R ? (C - r0) : R);
            unsigned char* dst = ring + (size_t)st * p.stage_bytes;
            if (p.cand) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full[st], (uint32_t)nr * row_bytes);
                    bulk_g2s(dst, p.cand + ((size_t)q * C + r0) * D, (uint32_t)nr * row_bytes, &full[st]);
                }
            } else {
                // R <= 64: two rows per lane at most
                int64_t ida = -1, idb = -1;
                if (lane < nr) ida = p.idx[(size_t)q * C + r0 + lane];
                if (lane + 32 < nr) idb = p.idx[(size_t)q * C + r0 + lane + 32];
                if (ida >= p.N) ida = -1;
                if (idb >= p.N) idb = -1;
                const int nvalid = __popc(__ballot_sync(FULL_MASK, ida >= 0)) + __popc(__ballot_sync(FULL_MASK, idb >= 0));
                if (lane == 0) mbar_arrive_expect_tx(&full[st], (uint32_t)nvalid * row_bytes);
                __syncwarp();
                if (ida >= 0) bulk_g2s(dst + (size_t)lane * row_bytes, p.X + (size_t)ida * D, row_bytes, &full[st]);
                if (idb >= 0) bulk_g2s(dst + (size_t)(lane + 32) * row_bytes, p.X + (size_t)idb * D, row_bytes, &full[st]);
            }
            if (++st == S) st = 0;
            if (++ti == tpq) {
                ti = 0; ++q; ++qs;
                if (++slot == QS) slot = 0;
            }
        }
        return;
    }

    if (warp == AS_WARP_CONV) {
        // --------------------------------------------------------------------- converter
        int slot = 0;
        uint32_t par = 0;
        const int D4 = D >> 2;
        for (int qs = 0; qs < nqueries; ++qs) {
            unsigned char* sl = qslot0 + (size_t)slot * p.qslot_bytes;
            const float4* qf = reinterpret_cast<const float4*>(sl);
            double2* qd = reinterpret_cast<double2*>(sl + qd_off);
            mbar_wait(&qfull[slot], par);
            double part = 0.0;
            for (int j = lane; j < D4; j += 32) {
                const float4 v = qf[j];
                const double a = (double)v.x, b = (double)v.y, c = (double)v.z, d = (double)v.w;
                qd[2 * j] = make_double2(a, b);
                qd[2 * j + 1] = make_double2(c, d);
                part = fma(a, a, part); part = fma(b, b, part); part = fma(c, c, part); part = fma(d, d, part);
            }
            part = warp_sum(part);
            if (lane == 0) *reinterpret_cast<double*>(sl + qn_off) = part;
            __syncwarp();
            if (lane == 0) mbar_arrive(&qready[slot]);
            if (++slot == QS) { slot = 0; par ^= 1u; }
        }
        return;
    }

    if (warp >= AS_WARP_RANK) {
        // ----------------------------------------------------------------------- rankers
        if (!p.fused) return;
        const int rt = tid - AS_WARP_RANK * 32;
        const int top_k = p.top_k;
        // Interleaved shape: this launch runs beside the previous one for most of its life, so its global writes (which
        // must follow that kernel's completion: both may write the same output arrays) are staged in shared memory --
        // top_k (score, position, id) per query, a few hundred bytes -- and flushed after griddepcontrol.wait at the
        // end, instead of stalling the ranking of the first query until the neighbour is done.
        const bool staged = p.out_stage_q > 0;
        double* st_s = reinterpret_cast<double*>(ring + (size_t)S * p.stage_bytes);
        long long* st_i = reinterpret_cast<long long*>(st_s + (size_t)p.out_stage_q * top_k);
        int32_t* st_p = reinterpret_cast<int32_t*>(st_i + (size_t)p.out_stage_q * top_k);
        bool ordered = !wait_writes || staged;
        for (int qs = 0; qs < nqueries; ++qs) {
            const int buf = qs % NB;
            const int64_t q = qfirst + qs;
            double2* sb = sc + (size_t)buf * P;
            mbar_wait(&sfull[buf], (uint32_t)(qs / NB) & 1u);
            double* os = staged ? st_s + (size_t)qs * top_k : p.out_scores + (size_t)q * top_k;
            int32_t* op = staged ? st_p + (size_t)qs * top_k : p.out_pos + (size_t)q * top_k;
            int64_t* oi = !p.out_ids ? nullptr
                                     : (staged ? reinterpret_cast<int64_t*>(st_i) + (size_t)qs * top_k : p.out_ids + (size_t)q * top_k);
            if (!ordered) { griddep_wait(); ordered = true; }       // first write of this thread
            if (C <= AS_RANK_COUNT_MAX) {
                // rank by counting; thread rt owns candidates rt and rt + 64
                const int c0 = rt, c1 = rt + AS_RTHREADS;
                double s0 = 0.0, s1 = 0.0;
                if (c0 < C) { const double2 v = sb[c0]; s0 = fidelity_pair(v.x, v.y); sb[c0].x = s0; }
                if (c1 < C) { const double2 v = sb[c1]; s1 = fidelity_pair(v.x, v.y); sb[c1].x = s1; }
                named_bar_sync(AS_BAR_RANK, AS_RTHREADS);
                int k0 = 0, k1 = 0;
                const int Ci = (int)C;
#pragma unroll 4
                for (int j = 0; j < Ci; ++j) {
                    const double sj = sb[j].x;
                    k0 += (sj > s0) || (sj == s0 && j < c0);
                    k1 += (sj > s1) || (sj == s1 && j < c1);
                }
                if (c0 < C && k0 < top_k) {
                    os[k0] = s0; op[k0] = c0;
                    if (oi) oi[k0] = p.idx[(size_t)q * C + c0];
                }
                if (c1 < C && k1 < top_k) {
                    os[k1] = s1; op[k1] = c1;
                    if (oi) oi[k1] = p.idx[(size_t)q * C + c1];
                }
            } else {
                for (int i = rt; i < P; i += AS_RTHREADS) {
                    double key = pos_inf();
                    long long tag = TagPad<long long>::value();
                    if (i < C) {
                        const double2 v = sb[i];
                        key = -fidelity_pair(v.x, v.y);
                        tag = i;
                    }
                    sb[i] = make_double2(key, __longlong_as_double(tag));
                }
                named_bar_sync(AS_BAR_RANK, AS_RTHREADS);
                bitonic_sort_aos(sb, P, rt, AS_RTHREADS, AS_BAR_RANK);
                for (int i = rt; i < top_k; i += AS_RTHREADS) {
                    const double2 v = sb[i];
                    const int pos = (int)__double_as_longlong(v.y);
                    os[i] = -v.x;
                    op[i] = pos;
                    if (oi) oi[i] = p.idx[(size_t)q * C + pos];
                }
            }
            __syncwarp();                                   // this warp's reads of sb are complete
            if (lane == 0) mbar_arrive(&sempty[buf]);
        }
        if (staged) {
            named_bar_sync(AS_BAR_RANK, AS_RTHREADS);       // both ranker warps have staged all their entries
            if (wait_writes) griddep_wait();                // the previous kernel is complete: its writes are behind ours
            const int total = nqueries * top_k;
            for (int i = rt; i < total; i += AS_RTHREADS) {
                const size_t o = (size_t)qfirst * top_k + i;
                p.out_scores[o] = st_s[i];
                p.out_pos[o] = st_p[i];
                if (p.out_ids) p.out_ids[o] = st_i[i];
            }
        }
        return;
    }

    // ------------------------------------------------------------------------- consumers
    double qreg[NCHUNK > 0 ? NCHUNK * 4 : 1];
    const double* qd = nullptr;
    double nq2 = 0.0;
    int qs_cached = -1;
    const int D4 = D >> 2;
    constexpr int NV = 2 * RB;
    constexpr int LPV = 32 / NV;

    // Each stage is always drained by the same team (S % NT == 0), and a warp waits for the tile
    // before it looks at the query slot, so no barrier is ever waited on more than one phase ahead.
    const int team = warp / G, sub = warp - team * G;
    if (team >= NT) return;
    bool ordered = !(wait_writes && !p.fused);
    int64_t q = qfirst;
    int ti = ti_first + team;
    while (ti >= tpq) { ti -= tpq; ++q; }
    int st = team;
    uint32_t par = 0;

    for (int it = team; it < ntiles; it += NT) {
        const int qs = (int)(q - qfirst);
        const int r0 = ti * R + sub * RB;                    // this warp's rows of the tile
        const int64_t left = C - r0;
        const int nr = left < 0 ? 0 : (left < RB ? (int)left : RB);
        const unsigned char* src = ring + (size_t)st * p.stage_bytes + (size_t)sub * RB * row_bytes;
        mbar_wait(&full[st], par);
        if (qs != qs_cached) {
            const int slot = qs % QS;
            const unsigned char* sl = qslot0 + (size_t)slot * p.qslot_bytes;
            mbar_wait(&qready[slot], (uint32_t)(qs / QS) & 1u);
            qd = reinterpret_cast<const double*>(sl + qd_off);
            nq2 = *reinterpret_cast<const double*>(sl + qn_off);
            if (NCHUNK > 0) {
#pragma unroll
                for (int t = 0; t < (NCHUNK > 0 ? NCHUNK : 0); ++t) {
                    const double2 a = *reinterpret_cast<const double2*>(qd + 4 * (lane + 32 * t));
                    const double2 b = *reinterpret_cast<const double2*>(qd + 4 * (lane + 32 * t) + 2);
                    qreg[4 * t + 0] = a.x; qreg[4 * t + 1] = a.y; qreg[4 * t + 2] = b.x; qreg[4 * t + 3] = b.y;
                }
            }
            qs_cached = qs;
        }

        double acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = 0.0;
        if (NCHUNK > 0) {
            if (nr == RB) {
#pragma unroll
                for (int t = 0; t < (NCHUNK > 0 ? NCHUNK : 0); ++t) {
                    float4 v[RB];
#pragma unroll
                    for (int i = 0; i < RB; ++i)
                        v[i] = reinterpret_cast<const float4*>(src + (size_t)i * row_bytes)[lane + 32 * t];
#pragma unroll
                    for (int i = 0; i < RB; ++i) fma4s(v[i], &qreg[4 * t], acc[2 * i], acc[2 * i + 1]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < RB; ++i) {
                    if (i < nr) {
#pragma unroll
                        for (int t = 0; t < (NCHUNK > 0 ? NCHUNK : 0); ++t)
                            fma4s(reinterpret_cast<const float4*>(src + (size_t)i * row_bytes)[lane + 32 * t],
                                  &qreg[4 * t], acc[2 * i], acc[2 * i + 1]);
                    }
                }
            }
        } else {
            for (int j = lane; j < D4; j += 32) {
                const double2 qa = *reinterpret_cast<const double2*>(qd + 4 * j);
                const double2 qb = *reinterpret_cast<const double2*>(qd + 4 * j + 2);
                const double qv[4] = {qa.x, qa.y, qb.x, qb.y};
#pragma unroll
                for (int i = 0; i < RB; ++i)
                    if (i < nr) fma4s(reinterpret_cast<const float4*>(src + (size_t)i * row_bytes)[j], qv, acc[2 * i],
                                      acc[2 * i + 1]);
            }
        }
        // every shared-memory read of this stage has completed (its value went through an FMA above), so the
        // stage goes back to the producer before the reduction, not after it
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        // lane L ends with value index L / LPV: even index = q.d, odd = |d|^2 of row index / 2
        const double tot = reduce_vals<NV>(acc, lane);
        const double nd2 = __shfl_down_sync(FULL_MASK, tot, LPV);

        const int i = lane / (2 * LPV);
        const bool owner = (lane & (2 * LPV - 1)) == 0 && i < nr;
        const int64_t c = r0 + i;
        double den = nq2 * nd2;
        if (owner && p.idx) {
            const int64_t id = p.idx[(size_t)q * C + c];
            if (id < 0 || id >= p.N) den = -1.0;
        }
        if (p.fused) {
            const int buf = qs % NB;           // free: the producer waited for the rankers before issuing this query
            if (owner) sc[(size_t)buf * P + c] = make_double2(tot, den);
            __syncwarp();
            if (lane == 0) mbar_arrive(&sfull[buf]);
        } else {
            if (!ordered) { griddep_wait(); ordered = true; }       // first write of this warp
            if (owner) {
                const double f = fidelity_pair(tot, den);
                p.out64[(size_t)q * C + c] = f;
                if (p.out32) p.out32[(size_t)q * C + c] = (float)f;
            }
        }
        ti += NT;
        while (ti >= tpq) { ti -= tpq; ++q; }
        st += NT;
        if (st >= S) { st -= S; par ^= 1u; }
    }
}

template <int NCHUNK, int RB, int CW>
static int launch_stream_cw(const AmpStreamParams& p, size_t smem_bytes, int grid, cudaStream_t st) {
    auto kern = amp_stream_kernel<NCHUNK, RB, CW>;
    QRAG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(as_threads(CW));
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p.overlap != QRAG_OVERLAP_NONE ? 1 : 0;
    QRAG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
    QRAG_LAUNCH_CHECK("amp_stream_kernel");
    return QRAG_OK;
}

template <int NCHUNK, int RB>
static int launch_stream(const AmpStreamParams& p, size_t smem_bytes, int grid, int cw, cudaStream_t st) {
    return cw == AS_CWARPS ? launch_stream_cw<NCHUNK, RB, AS_CWARPS>(p, smem_bytes, grid, st)
                           : launch_stream_cw<NCHUNK, RB, AS_CWARPS_HALF>(p, smem_bytes, grid, st);
}

// Shared-memory plan for a given tile height; returns the number of stages (0 = does not fit).
static int plan_stream(AmpStreamParams& p, int rb, int g, int cw, int nbuf_want, bool fused, size_t budget, size_t* smem_bytes) {
    const int D = p.D;
    p.G = g;
    p.R = g * rb;
    p.tpq = (int)ceil_div(p.C, p.R);
    p.stage_bytes = (uint32_t)(((size_t)p.R * D * 4 + 127) / 128 * 128);
    p.qslot_bytes = (uint32_t)((((size_t)D * 4 + 15) / 16 * 16 + (size_t)D * 8 + 16 + 127) / 128 * 128);
    p.P = fused ? next_pow2(p.C) : 0;
    // score buffers: enough that ranking never throttles the stream, within 32 KB (always >= 2)
    p.nbuf = 2;
    if (fused) {
        int want = (int)((size_t)(192 * 1024) / p.stage_bytes / p.tpq) + 3;
        if (nbuf_want > want) want = nbuf_want;
        if (want > AS_MAX_SBUFS) want = AS_MAX_SBUFS;
        while (want > 2 && (size_t)want * p.P * 16 > 32 * 1024) --want;
        p.nbuf = want;
    }
    const size_t scb = fused ? (size_t)p.nbuf * p.P * 16 : 0;
    int qs = 3, stages = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const size_t fixed = AS_HDR_BYTES + (size_t)qs * p.qslot_bytes + scb;
        if (fixed + 2 * (size_t)p.stage_bytes > budget) return 0;
        stages = (int)((budget - fixed) / p.stage_bytes);
        if (stages > AS_MAX_STAGES) stages = AS_MAX_STAGES;
        // enough query slots that the ring, not the slots, bounds what is in flight
        int want = stages / p.tpq + 2;
        if (want > AS_MAX_QSLOTS) want = AS_MAX_QSLOTS;
        if (want <= qs) break;
        qs = want;
    }
    const size_t fixed = AS_HDR_BYTES + (size_t)qs * p.qslot_bytes + scb;
    if (fixed + 2 * (size_t)p.stage_bytes > budget) return 0;
    stages = (int)((budget - fixed) / p.stage_bytes);
    if (stages > AS_MAX_STAGES) stages = AS_MAX_STAGES;
    p.teams = stages < cw / g ? stages : cw / g;
    stages -= stages % p.teams;
    p.qslots = qs;
    p.stages = stages;
    *smem_bytes = fixed + (size_t)stages * p.stage_bytes;
    return stages;
}

// Returns QRAG_OK and sets *handled = true if the streaming kernel took the job.
int amp_stream_try(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx, int64_t C, int D,
                   bool fused, double* out64, float* out32, int top_k, double* out_scores, int32_t* out_pos,
                   int64_t* out_ids, cudaStream_t st, bool* handled) {
    *handled = false;
    const DeviceProps& dp = device_props();
    QRAG_REQUIRE(dp.ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    if (D % 4 != 0 || D > 8192) return QRAG_OK;
#ifdef QRAG_TUNING
    if (getenv("QRAG_AMP_STREAM_OFF")) return QRAG_OK;          // force the plain-load kernel
#endif
    // Gathered candidates (X + idx) are one 4*D-byte bulk copy per row; the copy engine's per-operation cost
    // caps that form at ~2.5 TB/s for 1.5 KB rows, while the plain-load kernel (12 x 16 B loads in flight per
    // lane) reaches 3.2 - 4.1 TB/s on the same input.  The streaming kernel keeps the dense form.
#ifdef QRAG_TUNING
    if (cand == nullptr && !getenv("QRAG_AMP_STREAM_GATHER")) return QRAG_OK;
#else
    if (cand == nullptr) return QRAG_OK;
#endif
    if (((uintptr_t)Q | (uintptr_t)(cand ? cand : X)) % 16 != 0) return QRAG_OK;
    if (nq < 1 || C < 1) return QRAG_OK;

    AmpStreamParams p{};
    p.Q = Q; p.cand = cand; p.X = X; p.N = N; p.idx = idx; p.nq = nq; p.C = C; p.D = D;
    p.out64 = out64; p.out32 = out32; p.top_k = top_k; p.out_scores = out_scores; p.out_pos = out_pos;
    p.out_ids = out_ids; p.fused = fused ? 1 : 0;
    p.overlap = overlap_mode();

    const bool nchunk_path = (D == 128 || D == 256 || D == 384 || D == 512);
    // CTA shape: the interleaved half-size shape when the caller asked for it (QRAG_OVERLAP_INTERLEAVED) and every SM
    // gets a CTA; its shared-memory budget is half an SM's (two CTAs of consecutive launches share the SM)
    // CTAs per SM and launch in the interleaved shape (tuning build: QRAG_AMP_STREAM_CPS=2 fills both slots of every
    // SM from ONE launch, so a launch with no neighbour on the stream still uses the whole SM)
    int cps = 1;
#ifdef QRAG_TUNING
    if (const char* e = getenv("QRAG_AMP_STREAM_CPS")) { const int v = atoi(e); if (v == 1 || v == 2) cps = v; }
#endif
    if (nq < cps * dp.sm_count) cps = 1;
    const int q_per_cta = (int)ceil_div(nq, (int64_t)dp.sm_count * cps);
    const size_t out_stage = (size_t)q_per_cta * top_k * 20 + 16;      // staged results: fp64 score, int64 id, int32 position
    const bool half = p.overlap == QRAG_OVERLAP_INTERLEAVED && fused && nq >= dp.sm_count && out_stage <= 8 * 1024;
    const int cw = half ? AS_CWARPS_HALF : AS_CWARPS;
    const size_t budget = half ? ((size_t)dp.max_smem_optin + 1024) / 2 - 1024 - out_stage : (size_t)dp.max_smem_optin;
    const size_t tile_cap = half ? 12 * 1024 : 24 * 1024;
    // score buffers of the interleaved shape: one per query of the CTA, so that the stream never waits for the rankers
    const int nbuf_want = half ? q_per_cta + 1 : 0;
    p.out_stage_q = half ? q_per_cta : 0;
    // rows per warp: 4 while a warp's slice stays <= 8 KB, else 2, else 1; team size: the widest
    // tile (one bulk copy, one barrier pair) of <= 24 KB, so that >= 8 tiles are in flight per SM
    int rb = 4;
    while (rb > 1 && (size_t)rb * D * 4 > 8 * 1024) rb >>= 1;
    int g = cw;
    while (g > 1 && (size_t)g * rb * D * 4 > tile_cap) g >>= 1;
#ifdef QRAG_TUNING
    if (const char* e = getenv("QRAG_AMP_STREAM_RB")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) rb = v; }
    if (const char* e = getenv("QRAG_AMP_STREAM_G")) { const int v = atoi(e); if (v >= 1 && v <= cw && !(v & (v - 1))) g = v; }
#endif
    size_t smem_bytes = 0;
    while (plan_stream(p, rb, g, cw, nbuf_want, fused, budget, &smem_bytes) < 2) {
        if (g > 1) g >>= 1;
        else if (rb > 1 && !nchunk_path) rb >>= 1;
        else return QRAG_OK;
    }
    // per-CTA tile counts are kept in 32 bits
    if ((int64_t)p.tpq * nq / dp.sm_count > ((int64_t)1 << 30)) return QRAG_OK;
    if (half) smem_bytes += out_stage;                        // the staging area sits behind the ring

    int grid = half ? dp.sm_count * cps : dp.sm_count;
    if (fused) {
        if (grid > nq) grid = nq;
    } else {
        const int64_t T = (int64_t)nq * p.tpq;
        if (grid > T) grid = (int)T;
    }
    *handled = true;
    switch ((nchunk_path ? D / 128 : 0) * 8 + rb) {
        case 0 * 8 + 4: return launch_stream<0, 4>(p, smem_bytes, grid, cw, st);
        case 0 * 8 + 2: return launch_stream<0, 2>(p, smem_bytes, grid, cw, st);
        case 0 * 8 + 1: return launch_stream<0, 1>(p, smem_bytes, grid, cw, st);
        case 1 * 8 + 4: return launch_stream<1, 4>(p, smem_bytes, grid, cw, st);
        case 1 * 8 + 2: return launch_stream<1, 2>(p, smem_bytes, grid, cw, st);
        case 2 * 8 + 4: return launch_stream<2, 4>(p, smem_bytes, grid, cw, st);
        case 2 * 8 + 2: return launch_stream<2, 2>(p, smem_bytes, grid, cw, st);
        case 3 * 8 + 4: return launch_stream<3, 4>(p, smem_bytes, grid, cw, st);
        case 3 * 8 + 2: return launch_stream<3, 2>(p, smem_bytes, grid, cw, st);
        case 4 * 8 + 4: return launch_stream<4, 4>(p, smem_bytes, grid, cw, st);
        case 4 * 8 + 2: return launch_stream<4, 2>(p, smem_bytes, grid, cw, st);
        default: break;
    }
    *handled = false;        // (NCHUNK, RB) pair without an instantiation: the caller uses the plain kernel
    return QRAG_OK;
}

}  // namespace qrag
