// K3 / K4: brute-force search over a flat index on the 5th-generation tensor cores, with results
// identical to the exact CUDA-core search (search_exact.cu).
//
// The index is the faiss IndexFlat* the ingest tool writes (store_in_faiss.py:99-109); `.search` is
// never called by the reference, so the contract is faiss's API made canonical: fp32 inputs, fp64
// accumulation, order (best score, smaller id).
//
// A bf16 tensor-core product cannot give fp32-exact order, so the GEMM is used as a FILTER with a
// proven error bound, and only the survivors are scored exactly:
//
//   prepare (once per index)  Xb = 16-bit shadow of the corpus, shaped so that every metric is a plain
//                             inner product: IP: x (bf16); cosine: x/|x| against q/|q| (fp16: both sides are
//                             normalised, so the range is safe and the rounding error 8x smaller than bf16's);
//                             L2: [x, hi(|x|^2), lo(|x|^2)] against [2q, -1, -1] (bf16), i.e. s' = 2 q.x - |x|^2.
//   pass 1  sim_gemm<BUCKET>  S = Qb Xb^T over every `sample`-th 256-document tile; the epilogue keeps
//                             only the maximum of each 32-document bucket.
//   tau     per query, m_k = k-th largest bucket maximum.  At least k documents have approximate
//           score >= m_k, so the exact k-th best is >= m_k - eps and every member of the exact
//           top-k has approximate score >= tau = m_k - 2 eps  (eps = bound on |approx - exact|).
//   pass 2  sim_gemm<FILTER>  S over all tiles; the epilogue compares each 32-score chunk's maximum
//                             with tau and appends the (score, row) pairs >= tau to the query's list.
//                             S itself never leaves TMEM.
//   final   per query: a_k = k-th best approximate score of the list; candidates = entries
//           >= a_k - 2 eps (a superset of the exact top-k, typically k + a few hundred); exact fp64
//           rescoring with the code of the CUDA-core path (exact_score.cuh), sort by (score, id).
//   status  a query whose list overflowed is flagged (never silent); the caller reruns it exactly.
//
// sim_gemm: one persistent CTA per SM, warp-specialised.  Warp 0 = TMA producer (the query tile
// A = 128 x K stays resident in shared memory, document tiles B = 256 x 64 stream through a ring),
// warp 1 = MMA issuer (tcgen05.mma, M=128 N=256 K=16, fp32 accumulators in TMEM, two accumulator
// buffers of 256 columns so that the epilogue of tile i overlaps the MMAs of tile i+1), warp 2 =
// TMEM allocator, warps 4-7 = epilogue (tcgen05.ld 32x32b: one thread per query row).
#include "common.cuh"
#include "exact_score.cuh"
#include "sort.cuh"
#include "tma.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace qrag {

constexpr int TC_BM = 128;            // queries per tile (UMMA M)
constexpr int TC_BN = 256;            // documents per tile (UMMA N)
constexpr int TC_BK = 64;             // K elements per shared-memory chunk (128 B rows, SWIZZLE_128B)
constexpr int TC_UK = 16;             // K per tcgen05.mma (bf16)
constexpr int TC_EPI_WARPS = 16;       // four per TMEM lane quarter: each takes a quarter of the tile's columns
constexpr int TC_EPI_SPLIT = TC_EPI_WARPS / 4;
constexpr int TC_EPI_COLS = TC_BN / TC_EPI_SPLIT;
constexpr int TC_THREADS = (4 + TC_EPI_WARPS) * 32;
constexpr int TC_A_CHUNK = TC_BM * TC_BK * 2;     // 16 KB
constexpr int TC_B_STAGE = TC_BN * TC_BK * 2;     // 32 KB
constexpr int TC_BUCKET = 32;         // documents per bucket maximum (one tcgen05.ld chunk)
constexpr int TC_CAP_MIN = 16384;     // survivor list capacity per query: 64 k rounded up to a power of two,
constexpr int TC_CAP_MAX = 131072;    // within these bounds (TcPlan::cap)
constexpr int TC_MAX_CAND = 8192;     // largest exact-rescore capacity per query
constexpr int TC_MAX_SEGS = 640;      // survivor-list segments per query (TC_EPI_SPLIT per CTA of the query's group)
constexpr int TC_MAX_STAGES = 10;
constexpr int TC_MODE_BUCKET = 0, TC_MODE_FILTER = 1, TC_MODE_DUMP = 2;
constexpr int TC_HIST_BINS = QRAG_TC_HIST_BINS;   // survivor-score histogram per query (filter pass -> a_k)

struct TcGemmParams {
    int kchunks;            // ceil(Kp / 64)
    int ksteps_last;        // tcgen05.mma K-steps in the last chunk (1..4)
    int stages;             // ring depth
    int fp16;               // operands are fp16 (cosine: both sides normalised, so the range is safe) instead of bf16
    int a_resident;         // 1: the query tile (all K chunks) stays in shared memory; 0: its chunks stream with B
    int stage_bytes;        // 32 KB (B chunk) or 48 KB (B chunk + A chunk)
    int groups;             // query groups (of 128) in this launch
    int group0;             // first query group of this launch
    int nq;                 // total queries
    int64_t N;              // documents
    int ntiles;             // ceil(N / 256)
    int sample;             // pass 1 visits tiles 0, sample, 2*sample, ...
    int nbuckets;           // buckets per query in pass 1 = nsample_tiles * 8
    const float* tau;       // [nq] (filter)
    float* bmax;            // [nq, nbuckets] (bucket)
#ifdef QRAG_TUNING
    int debug_skip;         // kernel-tuning builds only (-DQRAG_TUNING): 1 / 3 / 4 skip parts of the epilogue
#endif
    int cap;                // survivor slots per query
    int seg_cap;            // survivor slots per (query, CTA, column quarter) segment
    unsigned int* cnt;      // [nq, TC_MAX_SEGS] survivors found per segment, may exceed seg_cap (filter)
    float2* surv;           // [nq, cap] (score, row as int bits), segment s at s * seg_cap (filter)
    float* dump;            // [nq, N] every approximate score (TC_MODE_DUMP: the diagnostic qrag_search_tc_scores)
};

// ------------------------------------------------------------------ PTX wrappers (tcgen05 / TMA)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile with 128-byte swizzled rows: 8-row groups are 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// ---- cta_group::2 forms (a CTA pair = one cluster of 2; rank 0 is the leader and issues the MMAs)
constexpr uint32_t TC_PEER_MASK = 0xFEFFFFFFu;     // clears the CTA-rank bit of a shared::cluster address: rank 0's copy
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(rank)
        : "memory");
}
// TMA load whose completion bytes are counted on the LEADER's barrier (same offset in rank 0)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & TC_PEER_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair when the MMAs issued so far are done
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// one lane of a converged warp; keeps the surrounding control flow warp-uniform, so that descriptors and barrier
// addresses stay in uniform registers (a branch on lane == 0 makes the compiler wrap every tcgen05 / TMA
// instruction in a per-lane loop of ~25 instructions, more than the 128 cycles an MMA lasts)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float neg_inf_f() { return __int_as_float(0xff800000); }

// ------------------------------------------------------------------------------- the GEMM
// First statement of every per-batch kernel (see launch_chained): let the next kernel of the stream start launching,
// then wait until everything before this kernel has completed and is visible.  No-ops in a plain launch.
__device__ __forceinline__ void pdl_enter() {
    griddep_launch_dependents();
    griddep_wait();
}

// CG == 1: one CTA computes 128 queries x 256 documents per tile.  CG == 2: a CTA pair computes 256 queries x 256
// documents per tile with tcgen05.mma.cta_group::2: each CTA keeps its own 128 query rows resident and loads HALF of
// every document tile (128 rows); the pair's tensor cores read both halves.  Per SM that halves the L2->SM bytes per
// flop and doubles the MMA work each in-flight stage feeds, which is what the single-CTA mainloop waits on.
template <int MODE, int CG>
__global__ void __launch_bounds__(TC_THREADS, 1)
sim_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const TcGemmParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = smem;                                             // [kchunks][128 rows][128 B]
    unsigned char* sB = sA + (p.a_resident ? (size_t)p.kchunks * TC_A_CHUNK : 0);   // [stages]{[256 rows][128 B], A chunk}
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)p.stages * p.stage_bytes);
    uint64_t* a_full = bars;
    uint64_t* full = bars + 1;
    uint64_t* empty = full + TC_MAX_STAGES;
    uint64_t* acc_full = empty + TC_MAX_STAGES;       // [2]
    uint64_t* acc_empty = acc_full + 2;               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.stages, KC = p.kchunks;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;      // position in the CTA pair; 0 issues the MMAs
    constexpr int B_ROWS = TC_BN / CG;                            // document rows this CTA loads per tile

    if (tid == 0) {
        // full / a_full / acc_empty are only used in the leader: both CTAs' producers and epilogues report there
        mbar_init(a_full, CG);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], CG); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], CG * TC_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (CG == 2) tmem_alloc_pair(tmem_slot, 512);
        else tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();                              // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_enter();                                                  // barriers and TMEM are set up while the previous kernel drains

    // work of this CTA (pair): query group g, tiles u0, u0 + cpg, ... (pass 1: of the sampled tiles)
    const int unit_id = blockIdx.x / CG;                          // CTA (CG 1) or pair (CG 2) index
    const int ugroups = p.groups / CG;                            // query groups are handled CG at a time
    const int g = (unit_id % ugroups) * CG + (int)rank;
    const int u0 = unit_id / ugroups;
    const int cpg = (gridDim.x / CG) / ugroups;
    const int step = MODE == TC_MODE_BUCKET ? p.sample : 1;       // the filter and dump passes visit every tile
    const int nunits = (p.ntiles + step - 1) / step;          // tiles this pass visits

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer
            if (p.a_resident) {
                if (CG == 2) {
                    if (rank == 0) mbar_arrive_expect_tx(a_full, (uint32_t)KC * TC_A_CHUNK * 2);
                    else mbar_arrive_cta(a_full, 0);
                    for (int kc = 0; kc < KC; ++kc)
                        tma_load_2d_pair(sA + (size_t)kc * TC_A_CHUNK, &mapA, a_full, kc * TC_BK, (p.group0 + g) * TC_BM);
                } else {
                    mbar_arrive_expect_tx(a_full, (uint32_t)KC * TC_A_CHUNK);
                    for (int kc = 0; kc < KC; ++kc)
                        tma_load_2d(sA + (size_t)kc * TC_A_CHUNK, &mapA, a_full, kc * TC_BK, (p.group0 + g) * TC_BM);
                }
            }
            int s = 0;
            uint32_t par = 0;
            for (int u = u0; u < nunits; u += cpg) {
                const int tile = u * step;
                for (int kc = 0; kc < KC; ++kc) {
                    mbar_wait(&empty[s], par ^ 1u);
                    if (CG == 2) {
                        // both halves of the tile report their bytes to the leader's barrier
                        if (rank == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)p.stage_bytes * 2);
                        else mbar_arrive_cta(&full[s], 0);
                        tma_load_2d_pair(sB + (size_t)s * p.stage_bytes, &mapB, &full[s], kc * TC_BK,
                                         tile * TC_BN + (int)rank * B_ROWS);
                    } else {
                        mbar_arrive_expect_tx(&full[s], (uint32_t)p.stage_bytes);
                        tma_load_2d(sB + (size_t)s * p.stage_bytes, &mapB, &full[s], kc * TC_BK, tile * TC_BN);
                        if (!p.a_resident)
                            tma_load_2d(sB + (size_t)s * p.stage_bytes + TC_B_STAGE, &mapA, &full[s], kc * TC_BK,
                                        (p.group0 + g) * TC_BM);
                    }
                    if (++s == S) { s = 0; par ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ------------------------------- MMA issuer: the whole warp runs the loop, one elected lane issues
            // instruction descriptor: D fp32, A/B bf16 or fp16, both K-major, N = 256, M = 128 (256 across a pair)
            // (bit 4: D fp32; bits 7 / 10: A / B format, 0 = fp16, 1 = bf16)
            const uint32_t idesc = (1u << 4) | (p.fp16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(TC_BN >> 3) << 17) |
                                   ((uint32_t)((TC_BM * CG) >> 4) << 24);
            // The issuing thread has 128 cycles per MMA: descriptors are built once and advanced by adding the
            // byte offset (>> 4) to their address field, the four K-steps of a chunk are unrolled.
            const uint64_t adesc0 = umma_desc_sw128(smem_u32(sA)), bdesc0 = umma_desc_sw128(smem_u32(sB));
            const uint32_t b_bytes = (uint32_t)B_ROWS * TC_BK * 2;               // this CTA's share of a document chunk
            const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4;
            const uint32_t a_step16 = p.a_resident ? (uint32_t)TC_A_CHUNK >> 4 : 0u;      // per k-chunk
            const uint64_t a_base = p.a_resident ? adesc0 : bdesc0 + (b_bytes >> 4);      // streamed A sits behind B
            const uint32_t a_stage16 = p.a_resident ? 0u : stage16;
            const int ks_last = p.ksteps_last;
            if (p.a_resident) mbar_wait(a_full, 0);
            int s = 0;
            uint32_t par = 0;
            int n = 0;
            for (int u = u0; u < nunits; u += cpg, ++n) {
                const int buf = n & 1;
                mbar_wait(&acc_empty[buf], (((uint32_t)n >> 1) & 1u) ^ 1u);      // epilogue drained this buffer
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)buf * TC_BN;
                uint32_t accumulate = 0;
                for (int kc = 0; kc < KC; ++kc) {
                    mbar_wait(&full[s], par);
                    tc_fence_after();
                    const uint64_t bd = bdesc0 + (uint64_t)((uint32_t)s * stage16);
                    const uint64_t ad = a_base + (uint64_t)((uint32_t)kc * a_step16 + (uint32_t)s * a_stage16);
                    const int ks = (kc == KC - 1) ? ks_last : TC_BK / TC_UK;
                    if (elect_one()) {
                        if (ks == TC_BK / TC_UK) {
#pragma unroll
                            for (int k = 0; k < TC_BK / TC_UK; ++k) {
                                const uint64_t off = (uint64_t)(k * TC_UK * 2 >> 4);
                                if (CG == 2) umma_bf16_pair(d, ad + off, bd + off, idesc, k == 0 ? accumulate : 1u);
                                else umma_bf16(d, ad + off, bd + off, idesc, k == 0 ? accumulate : 1u);
                            }
                        } else {
                            for (int k = 0; k < ks; ++k) {
                                const uint64_t off = (uint64_t)(k * TC_UK * 2 >> 4);
                                if (CG == 2) umma_bf16_pair(d, ad + off, bd + off, idesc, k == 0 ? accumulate : 1u);
                                else umma_bf16(d, ad + off, bd + off, idesc, k == 0 ? accumulate : 1u);
                            }
                        }
                        // frees the stage (in both CTAs of a pair) when these MMAs have read it
                        if (CG == 2) umma_commit_pair(&empty[s]);
                        else umma_commit(&empty[s]);
                    }
                    __syncwarp();
                    accumulate = 1;
                    if (++s == S) { s = 0; par ^= 1u; }
                }
                __syncwarp();
                if (elect_one()) {
                    if (CG == 2) umma_commit_pair(&acc_full[buf]);               // accumulator complete
                    else umma_commit(&acc_full[buf]);
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue
        const int ew = warp & 3;                              // TMEM lanes 32*ew .. 32*ew+31 (fixed by warp % 4)
        const int half = (warp - 4) >> 2;                      // columns [TC_EPI_COLS*half, +TC_EPI_COLS)
        const int row = ew * 32 + lane;
        const int q = (p.group0 + g) * TC_BM + row;
        const bool qvalid = q < p.nq;
        float tau = __int_as_float(0x7f800000);               // +inf: rows beyond nq never emit
        if (MODE == TC_MODE_FILTER && qvalid) tau = fmaxf(p.tau[q], -3.0e38f);   // -inf (rows past the end) never passes
        // this thread is the only writer of segment u0 of query q: no atomics on the hot path
        float2* seg = p.surv + (size_t)(qvalid ? q : 0) * p.cap + (size_t)(TC_EPI_SPLIT * u0 + half) * p.seg_cap;
        unsigned int found = 0;
        int n = 0;
        for (int u = u0; u < nunits; u += cpg, ++n) {
            const int buf = n & 1;
            const int tile = u * step;
            const int64_t doc0 = (int64_t)tile * TC_BN;
            const bool ragged = doc0 + TC_BN > p.N;
            mbar_wait(&acc_full[buf], ((uint32_t)n >> 1) & 1u);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)buf * TC_BN + (uint32_t)half * TC_EPI_COLS;
            // one 32-column chunk of this thread's row: bucket maximum, or threshold test and survivor append
            auto process = [&](uint32_t (&r)[32], const int c0) {
#ifdef QRAG_TUNING
                if (p.debug_skip == 3) { if ((r[0] ^ r[31]) == 0x12345u) found += 1; return; }
#endif
                if (MODE == TC_MODE_DUMP) {
                    if (qvalid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (doc0 + c0 + j < p.N) p.dump[(size_t)q * p.N + doc0 + c0 + j] = __uint_as_float(r[j]);
                    }
                    return;
                }
                if (ragged) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (doc0 + c0 + j >= p.N) r[j] = 0xff800000u;             // -inf: rows past the end
                }
                float t[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) t[j] = fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 16]));
#pragma unroll
                for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                    for (int j = 0; j < w; ++j) t[j] = fmaxf(t[j], t[j + w]);
                const float m = t[0];
#ifdef QRAG_TUNING
                if (p.debug_skip == 4) { if (m == 1234.5f) found += 1; return; }
#endif
                if (MODE == TC_MODE_BUCKET) {
                    if (qvalid) p.bmax[(size_t)q * p.nbuckets + (size_t)u * (TC_BN / TC_BUCKET) + (c0 >> 5)] = m;
                } else if (m >= tau) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = __uint_as_float(r[j]);
                        if (v >= tau) {
                            if (found < (unsigned)p.seg_cap) seg[found] = make_float2(v, __int_as_float((int)(doc0 + c0 + j)));
                            ++found;
                        }
                    }
                }
            };
            static_assert(TC_EPI_COLS == 64, "the epilogue below issues both of a warp's chunk loads up front");
#ifdef QRAG_TUNING
            if (p.debug_skip != 1)
#endif
            {
                uint32_t ra[32], rb[32];
                tmem_ld32(t0, ra);
                tmem_ld32(t0 + 32, rb);
                tmem_ld_wait();
                process(ra, half * TC_EPI_COLS);
                process(rb, half * TC_EPI_COLS + 32);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cta(&acc_empty[buf], 0);                // the leader issues the next MMAs
                else mbar_arrive(&acc_empty[buf]);
            }
        }
        if (MODE == TC_MODE_FILTER && qvalid) p.cnt[(size_t)q * TC_MAX_SEGS + TC_EPI_SPLIT * u0 + half] = found;
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();                              // neither CTA leaves while the other may still signal it
    if (warp == 2) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------- operand preparation
// The 16-bit operand format per metric.  Cosine normalises BOTH sides (rows in the shadow, the query in
// query_prepare), so every component is in [-1, 1] and fp16 -- 11 significant bits against bf16's 8 -- is safe; its
// rounding error, and with it the filter margin 2 eps and the number of rows rescored exactly, is 8x smaller.
// Inner product and L2 work on raw data of any scale and keep bf16's fp32 range.
__host__ __device__ __forceinline__ bool tc_operand_fp16(int metric) { return metric == QRAG_METRIC_COSINE; }

// v rounded to the operand format: its 16 bits and the value they stand for.  fp16 values below the smallest normal
// (2^-14) are flushed to zero here, so that no subnormal reaches the tensor core and the MEASURED rounding error is
// the error of what is really multiplied.
__device__ __forceinline__ unsigned short tc_round_operand(float v, bool fp16, float* back) {
    if (fp16) {
        if (fabsf(v) < 6.103515625e-05f) { *back = 0.f; return 0; }
        const __half h = __float2half_rn(v);
        *back = __half2float(h);
        return __half_as_ushort(h);
    }
    const __nv_bfloat16 r = __float2bfloat16_rn(v);
    *back = __bfloat162float(r);
    return __bfloat16_as_ushort(r);
}

// one warp per corpus row: |x|^2 in fp64, then the 16-bit shadow row for the metric
__global__ void __launch_bounds__(256) index_prepare_kernel(const float* __restrict__ X, int64_t N, int D, int Kp, int metric,
                                                            unsigned short* __restrict__ Xb, float* __restrict__ aux) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= N) return;
    const bool fp16 = tc_operand_fp16(metric);
    const float* x = X + (size_t)row * D;
    double part = 0.0;
    for (int j = lane; j < D; j += 32) { const double v = (double)x[j]; part = fma(v, v, part); }
    const double n2 = warp_sum(part);
    const double nrm = sqrt(n2);
    unsigned short* o = Xb + (size_t)row * Kp;
    double err2 = 0.0;                                           // |b - round(b)|^2 over the D data columns (exact: the
    for (int j = lane; j < Kp; j += 32) {                        // difference of an fp32 and its 16-bit rounding is an fp32)
        float v = 0.f;
        if (j < D) {
            v = x[j];
            if (metric == QRAG_METRIC_COSINE) v = nrm > 0.0 ? (float)((double)v / nrm) : 0.f;
        } else if (metric == QRAG_METRIC_L2 && j < D + 2) {
            const float n2f = (float)n2;
            const float hi = __bfloat162float(__float2bfloat16_rn(n2f));
            v = (j == D) ? hi : (n2f - hi);
        }
        float back;
        o[j] = tc_round_operand(v, fp16, &back);
        if (j < D) { const double d = (double)(v - back); err2 = fma(d, d, err2); }
    }
    err2 = warp_sum(err2);
    if (lane == 0) {
        const float up = __double2float_ru(nrm);                 // max |x| (rounded up); floats >= 0 order like ints
        atomicMax(reinterpret_cast<int*>(aux), __float_as_int(up));
        const float eup = __double2float_ru(sqrt(err2) * 1.000001);   // max over rows of the rounding-error norm
        atomicMax(reinterpret_cast<int*>(aux) + 1, __float_as_int(eup));
    }
}

// one warp per query row: 16-bit query operand for the metric, its norm (rounded up) and its rounding-error norm.
// Cosine: the operand is q / |q| (the score is then the cosine itself, and the operand fits fp16 whatever the
// scale of q); inner product: q; L2: [2q, -1, -1].
__global__ void __launch_bounds__(256) query_prepare_kernel(const float* __restrict__ Q, int nq, int nq_pad, int D, int Kp,
                                                            int metric, unsigned short* __restrict__ Qb,
                                                            float* __restrict__ qnorm, float* __restrict__ qerr) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= nq_pad) return;
    unsigned short* o = Qb + (size_t)row * Kp;
    if (row >= nq) {
        for (int j = lane; j < Kp; j += 32) o[j] = 0;
        return;
    }
    const bool fp16 = tc_operand_fp16(metric);
    const float* x = Q + (size_t)row * D;
    double nrm = 1.0;
    if (metric == QRAG_METRIC_COSINE) {
        double p2 = 0.0;
        for (int j = lane; j < D; j += 32) { const double v = (double)x[j]; p2 = fma(v, v, p2); }
        nrm = sqrt(warp_sum(p2));
    }
    double part = 0.0, err2 = 0.0;
    for (int j = lane; j < Kp; j += 32) {
        float v = 0.f;
        unsigned short bits = 0;
        if (j < D) {
            v = x[j];
            if (metric == QRAG_METRIC_COSINE) v = nrm > 0.0 ? (float)((double)v / nrm) : 0.f;
            part = fma((double)v, (double)v, part);
            float back;
            bits = tc_round_operand(v, fp16, &back);
            const double d = (double)(v - back);                 // a - round(a), exact
            err2 = fma(d, d, err2);
            if (metric == QRAG_METRIC_L2) bits = tc_round_operand(2.f * v, fp16, &back);   // exact: bf16(2q) == 2 bf16(q)
        } else if (metric == QRAG_METRIC_L2 && j < D + 2) {
            float back;
            bits = tc_round_operand(-1.f, fp16, &back);
        }
        o[j] = bits;
    }
    part = warp_sum(part);
    err2 = warp_sum(err2);
    if (lane == 0) {
        qnorm[row] = __double2float_ru(sqrt(part) * 1.000001);
        qerr[row] = __double2float_ru(sqrt(err2) * 1.000001);
    }
}

// Bound on |approximate - exact| in the units of the approximate score.  With a = the query operand, b = the
// document operand, a~ = bf16(a), b~ = bf16(b):  a~.b~ - a.b = (a~ - a).b~ + a.(b~ - b), so by Cauchy-Schwarz
//     |a~.b~ - a.b| <= |a~ - a| |b~| + |a| |b~ - b|.
// The rounding-error norms are MEASURED, not assumed: qe = |q - bf16(q)| per query (query_prepare_kernel), xe = the
// maximum over the corpus rows of |b - bf16(b)| (index_prepare_kernel, aux[1]; reduced over shards like aux[0]).
// A worst-case constant would be 2^-8 per operand (bf16 keeps 8 significant bits), twice what random data shows
// (~1.65e-3 |x|), and 2^-9 -- what this bound assumed before -- is not a bound at all.  On top: fp32 accumulation of
// Kp exact bf16 products inside the tensor core, Kp * 2^-22 of |a~||b~| (generous).
__device__ __forceinline__ float tc_eps(int metric, int Kp, float qn, float qe, float xmax, float xe) {
    const float g = (float)Kp * 2.4e-7f;
    float e;
    if (metric == QRAG_METRIC_IP) {
        const float xt = xmax + xe;                              // |x~| <= |x| + |x~ - x|
        e = qe * xt + qn * xe + g * (qn + qe) * xt;
    } else if (metric == QRAG_METRIC_COSINE) {
        // both operands are normalised in fp32 before the 16-bit rounding (2.4e-7 each: the fp32 division and the
        // norm); qn is the norm of the normalised query operand, ~1
        const float xt = 1.0000002f + xe;
        e = (qe + 2.4e-7f * qn) * xt + qn * (xe + 2.4e-7f) + g * (qn + qe) * xt;
    } else {
        // s' = 2 q.x - |x|^2 with |x|^2 split into hi + lo (relative error 2^-15, generous) in two extra columns
        const float xt = xmax + xe;
        e = 2.f * (qe * xt + qn * xe) + 3.1e-5f * xmax * xmax + g * (2.f * (qn + qe) * xt + 1.01f * xmax * xmax);
    }
    return e * 1.0001f + 1e-30f;
}

// order-preserving map float -> uint32 (larger float <-> larger integer) and back
__device__ __forceinline__ uint32_t f2sortable(float f) {
    const uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float sortable2f(uint32_t u) {
    return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xffffffffu));
}

struct RadixSel { int hist[256]; uint32_t prefix; int remaining; int count; };     // lives in shared memory

// f(load(j)) for j = start, start + stride, ... < n, with U independent loads in flight per thread
// (one load at a time makes the select kernels latency-bound: ncu showed 28 % of the samples on the
// first use of each loaded value).
template <int U, class Load, class F>
__device__ __forceinline__ void for_strided(int start, int n, int stride, Load load, F f) {
    for (int i = start; i < n; i += U * stride) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = i + u * stride;
            v[u] = j < n ? load(j) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < n) f(v[u]);
    }
}

// k-th largest (1-based) of the values `for_each` enumerates (each thread its own part; the
// enumeration must be repeatable and hold at least k values).  Radix select, 4 x 8 bits.
template <class ForEach>
__device__ float block_kth_largest(RadixSel& rs, int k, ForEach for_each) {
    if (threadIdx.x == 0) { rs.prefix = 0; rs.remaining = k; }
    __syncthreads();
    for (int pass = 3; pass >= 0; --pass) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) rs.hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = rs.prefix;
        const uint32_t mask = pass == 3 ? 0u : (0xffffffffu << (8 * (pass + 1)));
        // scores cluster (same sign and exponent): in the upper passes nearly every value falls into one bin and
        // per-value atomics on one shared address serialise.  Each thread run-length-merges its consecutive values.
        uint32_t last_bin = 0xffffffffu;
        int run = 0;
        for_each([&](float v) {
            const uint32_t u = f2sortable(v);
            if ((u & mask) == prefix) {
                const uint32_t bin = (u >> (8 * pass)) & 255u;
                if (bin == last_bin) {
                    ++run;
                } else {
                    if (run) atomicAdd(&rs.hist[last_bin], run);
                    last_bin = bin;
                    run = 1;
                }
            }
        });
        if (run) atomicAdd(&rs.hist[last_bin], run);
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp 0: the bin d (from the top) where the running count reaches `remaining`.  Lane l owns bins
            // 255-8l .. 248-8l; an exclusive scan over lanes gives the count above each lane's bins.
            const int lane = threadIdx.x;
            int h[8], mine = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) { h[u] = rs.hist[255 - 8 * lane - u]; mine += h[u]; }
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o) incl += t;
            }
            const int remaining = rs.remaining;
            const int above = incl - mine;
            const bool here = above < remaining && incl >= remaining;        // exactly one lane (total >= remaining)
            const unsigned who = __ballot_sync(FULL_MASK, here);
            const int total = __shfl_sync(FULL_MASK, incl, 31);
            if (who == 0) {                                                  // fewer than `remaining` values: bin 0
                if (lane == 0) { rs.remaining = remaining - (total - rs.hist[0]); rs.prefix = prefix; }
            } else if (here) {
                int rem = remaining - above, d = 255 - 8 * lane;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (h[u] >= rem) break;
                    rem -= h[u];
                    --d;
                }
                if (d < 255 - 8 * lane - 7) d = 255 - 8 * lane - 7;
                rs.remaining = rem;
                rs.prefix = prefix | ((uint32_t)d << (8 * pass));
            }
        }
        __syncthreads();
    }
    return sortable2f(rs.prefix);
}

// out[0..k) = the k largest enumerated values as a multiset (any order), -inf padded when fewer exist
template <class ForEach>
__device__ void block_topk_values(RadixSel& rs, int k, int total, ForEach for_each, float* __restrict__ out) {
    float t = neg_inf_f();
    if (total >= k) t = block_kth_largest(rs, k, for_each);
    if (threadIdx.x == 0) rs.count = 0;
    __syncthreads();
    const bool all = total < k;
    for_each([&](float v) {
        if (all || v > t) {
            const unsigned am = __activemask();                 // one atomic per converged group
            const int lane = threadIdx.x & 31, leader = __ffs(am) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&rs.count, __popc(am));
            base = __shfl_sync(am, base, leader);
            const int pos = base + __popc(am & ((1u << lane) - 1u));
            if (pos < k) out[pos] = v;
        }
    });
    __syncthreads();
    for (int i = rs.count + threadIdx.x; i < k; i += blockDim.x) out[i] = t;       // ties with the k-th / padding
}

// The select kernels make four radix passes (and the top-k form a fifth, collecting) over the same few thousand
// values: read them from global memory once into shared memory when they fit (ncu r02: tau_union barrier- and
// scoreboard-bound, 40 registers = 6 CTAs per SM = 1.15 waves of 1024 queries; now 8 CTAs per SM, one wave).
constexpr int SEL_CACHE = 5120;                // 20 KB: eight CTAs of a select kernel still share an SM
template <class Load>
__device__ __forceinline__ void sel_fill(float* sv, int total, Load load) {
    for (int i = threadIdx.x; i < total; i += 8 * blockDim.x) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = i + u * blockDim.x;
            v[u] = j < total ? load(j) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = i + u * blockDim.x;
            if (j < total) sv[j] = v[u];
        }
    }
    __syncthreads();
}

// per query: the k largest bucket maxima of this shard -> bm_top [nq, k]
__global__ void __launch_bounds__(256, 8) bucket_topk_kernel(const float* __restrict__ bmax, int nbuckets, int k,
                                                             float* __restrict__ bm_top) {
    pdl_enter();
    __shared__ RadixSel rs;
    __shared__ float sv[SEL_CACHE];
    const int q = blockIdx.x;
    const float* row = bmax + (size_t)q * nbuckets;
    const bool cached = nbuckets <= SEL_CACHE;
    if (cached) sel_fill(sv, nbuckets, [&](int j) { return row[j]; });
    auto fe = [&](auto f) {
        if (cached) {
            for (int j = threadIdx.x; j < nbuckets; j += blockDim.x) f(sv[j]);
        } else {
            for_strided<8>(threadIdx.x, nbuckets, blockDim.x, [&](int j) { return row[j]; }, f);
        }
    };
    block_topk_values(rs, k, nbuckets, fe, bm_top + (size_t)q * k);
}

// per query: m_k = k-th largest bucket maximum over all G shards' lists; tau = m_k - 2 eps, rounded down
// (single shard: bm_top_all == nullptr and the k-th largest is taken straight from this shard's bucket maxima).
// Also lays out the query's survivor histogram (surv_hist_kernel): TC_HIST_BINS uniform bins from tau up to the
// largest sampled bucket maximum (scores above it land in the last bin).
__global__ void __launch_bounds__(256, 8) tau_union_kernel(const float* __restrict__ bm_top_all, int G, int nq, int k, int kt,
                                                           int metric,
                                                           int Kp, const float* __restrict__ qnorm, const float* __restrict__ qerr,
                                                           const float* __restrict__ aux,
                                                           const float* __restrict__ bmax, int nbuckets,
                                                           float* __restrict__ tau, float* __restrict__ eps,
                                                           float* __restrict__ hinv) {
    pdl_enter();
    __shared__ RadixSel rs;
    __shared__ float s_max[8];
    __shared__ float sv[SEL_CACHE];
    const int q = blockIdx.x;
    float mk = neg_inf_f(), top = neg_inf_f();
    auto track = [&](auto f) { return [&top, f](float v) { top = fmaxf(top, v); f(v); }; };
    const int64_t total = bm_top_all != nullptr ? (int64_t)G * kt : (int64_t)nbuckets;
    const bool cached = total <= SEL_CACHE;
    if (cached) {
        if (bm_top_all != nullptr) {
            sel_fill(sv, (int)total, [&](int j) {
                const int g = j / kt;
                return bm_top_all[((size_t)g * nq + q) * kt + (j - g * kt)];
            });
        } else {
            const float* row = bmax + (size_t)q * nbuckets;
            sel_fill(sv, (int)total, [&](int j) { return row[j]; });
        }
    }
    auto fe = [&](auto f) {
        if (cached) {
            auto tf = track(f);
            for (int j = threadIdx.x; j < (int)total; j += blockDim.x) tf(sv[j]);
        } else if (bm_top_all != nullptr) {
            for (int g = 0; g < G; ++g) {
                const float* row = bm_top_all + ((size_t)g * nq + q) * kt;
                for_strided<4>(threadIdx.x, kt, blockDim.x, [&](int j) { return row[j]; }, track(f));
            }
        } else {
            const float* row = bmax + (size_t)q * nbuckets;
            for_strided<8>(threadIdx.x, nbuckets, blockDim.x, [&](int j) { return row[j]; }, track(f));
        }
    };
    if (total >= k) mk = block_kth_largest(rs, k, fe);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) top = fmaxf(top, __shfl_xor_sync(FULL_MASK, top, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = top;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) top = fmaxf(top, s_max[w]);
        const float e = tc_eps(metric, Kp, qnorm[q], qerr[q], aux[0], aux[1]);
        eps[q] = e;
        float t = mk - 2.f * e;                                   // -inf stays -inf (fewer than k buckets)
        t = t - fabsf(t) * 2.4e-7f - 1e-37f;                      // the subtraction above rounds to nearest: step down
        tau[q] = t;
        // The bins reach 1.6x beyond the largest SAMPLED maximum: the corpus holds ~ `sample` times more documents
        // above it than the sample does, and all of them share the last bin -- if that bin alone held k survivors the
        // cut could not be placed any higher (measured: 1 % of the queries at 1/39 sampling without the extension).
        const float width = 1.6f * (top - t);                     // > 0 whenever both are finite (top >= m_k > tau)
        hinv[q] = (width > 0.f && width < 3.0e38f) ? (float)(TC_HIST_BINS - 1) / width : 0.f;
    }
}

// per query: histogram of this shard's survivors over their approximate score, bin = min(BINS - 1, (v - tau) * hinv).
// One pass over the survivor lists.  It replaces a select over them and, sharded, the exchange of k scores per query
// and shard by an all-reduce(SUM) of TC_HIST_BINS counters: the bins are laid out identically on every shard (tau and
// hinv come from the all-gathered bucket maxima), and the highest bin edge with k survivors at or above it bounds
// the k-th best approximate score from below, which is all the candidate cut needs.  (int)NaN == 0, +inf saturates.
// (Counting in the filter GEMM's epilogue instead was measured: +14 % on the GEMM, whose epilogue is its critical path.)
// f(entry) for every survivor of one query.  seg_n [nseg] (shared memory) = entries per segment.  The lists are many
// short segments (config 3: 72 segments of ~22 survivors), so walking them one segment per warp at a time is a chain
// of dependent-latency loads (ncu r02: both kernels below latency-bound at 12-15 us for 13 MB).  Here a warp has the
// first 32 entries of FOUR segments in flight at once, then mops up the rare longer segments four loads deep.
template <class F>
__device__ __forceinline__ void for_each_survivor(const float2* __restrict__ sv, const int* seg_n, int nseg, int seg_cap, F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int s0 = 4 * warp; s0 < nseg; s0 += 4 * nwarps) {
        float2 e[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int sg = s0 + u;
            ok[u] = sg < nseg && lane < seg_n[sg];
            e[u] = ok[u] ? sv[(size_t)sg * seg_cap + lane] : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (ok[u]) f(e[u]);
    }
    for (int sg = warp; sg < nseg; sg += nwarps) {
        const int n = seg_n[sg];
        const float2* sp = sv + (size_t)sg * seg_cap;
        for (int i0 = 32 + lane; i0 < n; i0 += 128) {
            float2 e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) e[u] = (i0 + 32 * u < n) ? sp[i0 + 32 * u] : make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + 32 * u < n) f(e[u]);
        }
    }
}

// CTA size of the two kernels that walk one query's survivor segments.  A few queries over a large shard (nq <= 128:
// the single-CTA GEMM form, one segment per (SM, column quarter) = 592 of them) leave each walking CTA a long chain of
// dependent-latency loads -- 23 + 28 us of a 211 us one-query search at 8 warps (profiles/r02b_search_nq1_launches.csv)
// -- so those run 32 warps wide; many queries with short segment lists keep 8 warps (more CTAs per SM).
static inline int seg_walk_threads(int nseg) { return nseg >= 256 ? 1024 : 256; }

__global__ void __launch_bounds__(1024) surv_hist_kernel(const unsigned int* __restrict__ cnt, const float2* __restrict__ surv,
                                                        int q0, int nseg, int seg_cap, int cap,
                                                        const float* __restrict__ tau, const float* __restrict__ hinv,
                                                        int* __restrict__ hist) {
    pdl_enter();
    __shared__ int h[TC_HIST_BINS];
    __shared__ int seg_n[TC_MAX_SEGS];
    const int q = q0 + blockIdx.x;
    const int tid = threadIdx.x;
    for (int i = tid; i < TC_HIST_BINS; i += blockDim.x) h[i] = 0;
    for (int sgi = tid; sgi < nseg; sgi += blockDim.x) {
        const unsigned int c = cnt[(size_t)q * TC_MAX_SEGS + sgi];
        seg_n[sgi] = c < (unsigned)seg_cap ? (int)c : seg_cap;
    }
    const float t = tau[q], hv = hinv[q];
    __syncthreads();
    for_each_survivor(surv + (size_t)q * cap, seg_n, nseg, seg_cap, [&](float2 e) {
        atomicAdd(&h[max(0, min(TC_HIST_BINS - 1, (int)((e.x - t) * hv)))], 1);
    });
    __syncthreads();
    for (int i = tid; i < TC_HIST_BINS; i += blockDim.x) hist[(size_t)q * TC_HIST_BINS + i] = h[i];
}

// ---------------------------------------------------------------------------------- final stage
struct TcFinalParams {
    const float* Q; const float* X; int q0; int nq; int64_t N; int D; int k; int metric; int64_t id_base;
    int nseg, seg_cap, cap, cand_cap;
    const int* hist;                         // [nq, TC_HIST_BINS] survivors per score bin, summed over ALL shards
    const float* tau; const float* hinv;     // [nq] the histogram's origin and bins per unit of score
    const unsigned int* cnt; const float2* surv; const float* eps;
    double* out_scores; int64_t* out_ids; int32_t* status;
    int* cand_rows; double* cand_key; int* cand_m;     // [nq, cand_cap], [nq, cand_cap], [nq, 2] (count, overflow)
    double* cand_fid;                                  // [nq, cand_cap] amplitude fidelity of the row (packed form) or null
    long long* pack; int kk;                           // packed form: [nq, 3 kk + 1] records cut to kk entries, or null
    int fuse_sort;                                     // tc_rescore_bulk_kernel sorts and writes the list itself (grid.y == 1)
#ifdef QRAG_TUNING
    int tune;                                          // 2: register form of tc_rescore even where the bulk form fits; 4: never fuse the sort
#endif
};

template <typename K, typename T>
__device__ void bitonic_sort_kt(K* key, T* tag, int P) {
    for (int kk = 2; kk <= P; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int hi = lo | j;
                const K kl = key[lo], kh = key[hi];
                const T tl = tag[lo], th = tag[hi];
                const bool asc = (lo & kk) == 0;
                const bool hi_first = (kh < kl) || (kh == kl && th < tl);
                if (hi_first == asc) { key[lo] = kh; key[hi] = kl; tag[lo] = th; tag[hi] = tl; }
            }
            __syncthreads();
        }
    }
}

// The final stage is three kernels, so that each runs at the occupancy its work allows (as one kernel, 55 KB of
// shared memory and 78 registers held it to 3 CTAs per SM and its phases could not overlap: ncu, k = 1000).
//
// (1) one CTA per query: a_k = a lower bound of the k-th best approximate score over all shards, read off the
// survivor histogram the filter pass built (summed over the shards by the caller): the highest bin edge that still
// has k survivors at or above it.  Candidates = this shard's survivors >= a_k - 2 eps -> cand_rows, cand_m.
__global__ void __launch_bounds__(1024) tc_collect_kernel(const TcFinalParams p) {
    pdl_enter();
    __shared__ int seg_n[TC_MAX_SEGS];
    __shared__ int s_m, s_bad;
    __shared__ float s_ak;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = p.q0 + blockIdx.x;
    const int k = p.k;
    const float2* sv = p.surv + (size_t)q * p.cap;
    int* cand = p.cand_rows + (size_t)q * p.cand_cap;

    if (tid == 0) { s_m = 0; s_bad = 0; }
    __syncthreads();
    for (int sgi = tid; sgi < p.nseg; sgi += blockDim.x) {
        const unsigned int c = p.cnt[(size_t)q * TC_MAX_SEGS + sgi];
        seg_n[sgi] = c < (unsigned)p.seg_cap ? (int)c : p.seg_cap;
        if (c > (unsigned)p.seg_cap) s_bad = 1;
    }
    if (warp == 0) {
        // lane l owns bins 8l .. 8l+7; suffix counts from the top bin down; j* = the largest bin index with at least k
        // survivors in bins >= j*.  Every survivor v in a bin >= j has (v - tau) * hinv >= j up to two fp32 roundings,
        // so tau + (j / hinv)(1 - 4e-6), stepped down once more for the final add, is <= all of them.
        static_assert(TC_HIST_BINS == 256, "8 bins per lane");
        const int* h = p.hist + (size_t)q * TC_HIST_BINS + 8 * lane;
        int c[8], mine = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) { c[u] = h[u]; mine += c[u]; }
        int incl = mine;                                    // sum over lanes >= this one (suffix scan)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_down_sync(FULL_MASK, incl, o);
            if (lane + o < 32) incl += t;
        }
        int above = incl - mine, jbest = -1;                // survivors in the bins above this lane's
#pragma unroll
        for (int u = 7; u >= 0; --u) {
            above += c[u];
            if (jbest < 0 && above >= k) jbest = 8 * lane + u;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) jbest = max(jbest, __shfl_xor_sync(FULL_MASK, jbest, o));
        if (lane == 0) {
            float ak = neg_inf_f();                         // fewer than k survivors in total: keep them all
            if (jbest >= 0) {
                const float t = p.tau[q], hv = p.hinv[q];
                ak = t;
                if (jbest > 0 && hv > 0.f) {
                    ak = t + ((float)jbest / hv) * (1.f - 4e-6f);
                    ak = ak - fabsf(ak) * 2.4e-7f - 1e-37f;
                    ak = fmaxf(ak, t);
                }
            }
            s_ak = ak;
        }
    }
    __syncthreads();
    const float ak = s_ak;
    float thr = ak - 2.f * p.eps[q];                                        // -inf stays -inf
    thr = thr - fabsf(thr) * 2.4e-7f - 1e-37f;
    // candidates (any order: the final sort is a total order on (score, id))
    for_each_survivor(sv, seg_n, p.nseg, p.seg_cap, [&](float2 e) {
        if (e.x >= thr) {
            const unsigned am = __activemask();                             // one atomic per converged group
            const int leader = __ffs(am) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&s_m, __popc(am));
            base = __shfl_sync(am, base, leader);
            const int pos = base + __popc(am & ((1u << lane) - 1u));
            if (pos < p.cand_cap) cand[pos] = __float_as_int(e.y);
        }
    });
    __syncthreads();
    if (tid == 0) {
        int m = s_m, bad = s_bad;
        if (m > p.cand_cap) { m = p.cand_cap; bad = 1; }
        p.cand_m[2 * q] = m;
        p.cand_m[2 * q + 1] = bad;
    }
}

// (3) one CTA per query: sort the (key, row) pairs by (score, id), write the shard's list and its status.  The body
// is shared by the stand-alone kernel and by the rescoring kernel, which runs it itself when one CTA scored the
// whole list of its query (tc_rescore_bulk_kernel, fuse_sort).  `smem` holds 12 bytes per (power-of-two) list slot.
__device__ __forceinline__ void tc_sort_body(const TcFinalParams& p, int q, unsigned char* smem) {
    const int tid = threadIdx.x;
    const int k = p.k;
    const bool l2 = p.metric == QRAG_METRIC_L2;
    const int m = p.cand_m[2 * q];
    int P2 = 1;
    while (P2 < m) P2 <<= 1;
    double* ckey = reinterpret_cast<double*>(smem);                         // [P2]
    int* ctag = reinterpret_cast<int*>(ckey + P2);                          // [P2] corpus rows = the sort tags
    const int* cand = p.cand_rows + (size_t)q * p.cand_cap;
    const double* gkey = p.cand_key + (size_t)q * p.cand_cap;
    for (int i = tid; i < P2; i += blockDim.x) {
        ckey[i] = i < m ? gkey[i] : pos_inf();
        ctag[i] = i < m ? cand[i] : 0x7fffffff;
    }
    __syncthreads();
    bitonic_sort_kt<double, int>(ckey, ctag, P2);
    for (int i = tid; i < k; i += blockDim.x) {
        double kv = pos_inf();
        long long tv = 0x7fffffffffffffffLL;
        if (i < m && ctag[i] != 0x7fffffff) { kv = ckey[i]; tv = p.id_base + ctag[i]; }
        if (tv == 0x7fffffffffffffffLL) { tv = -1; kv = l2 ? pos_inf() : -pos_inf(); }
        else if (!l2) kv = -kv;
        p.out_scores[(size_t)q * k + i] = kv;
        p.out_ids[(size_t)q * k + i] = tv;
    }
    if (tid == 0) p.status[q] = p.cand_m[2 * q + 1];
}

// (3') packed form of (3) for the search + rerank path: the sorted list goes straight into the record the exchange
// sends to the query's owner -- [0] header (valid entries | bad << 32), then kk score bits, kk ids, kk fidelity bits
// -- cut to the kk best entries.  The sort tag is row << 13 | candidate slot (rows are unique within a list, so the
// order is still (score, id)); the slot finds the row's fidelity after the sort.  16 bytes of `smem` per list slot.
constexpr int TC_SLOT_BITS = 13;
static_assert((1 << TC_SLOT_BITS) >= TC_MAX_CAND, "a candidate slot must fit the tag");
__device__ __forceinline__ void tc_sort_pack_body(const TcFinalParams& p, int q, unsigned char* smem) {
    const int tid = threadIdx.x;
    const int k = p.k, kk = p.kk;
    const bool l2 = p.metric == QRAG_METRIC_L2;
    const int m = p.cand_m[2 * q];
    int P2 = 1;
    while (P2 < m) P2 <<= 1;
    double* ckey = reinterpret_cast<double*>(smem);                         // [P2]
    long long* ctag = reinterpret_cast<long long*>(ckey + P2);              // [P2]
    const int* cand = p.cand_rows + (size_t)q * p.cand_cap;
    const double* gkey = p.cand_key + (size_t)q * p.cand_cap;
    const double* gfid = p.cand_fid + (size_t)q * p.cand_cap;
    for (int i = tid; i < P2; i += blockDim.x) {
        ckey[i] = i < m ? gkey[i] : pos_inf();
        ctag[i] = i < m ? (((long long)cand[i] << TC_SLOT_BITS) | i) : 0x7fffffffffffffffLL;
    }
    __syncthreads();
    bitonic_sort_kt<double, long long>(ckey, ctag, P2);
    const int valid = m < k ? m : k;                                        // this shard's list has min(m, k) entries
    long long* rec = p.pack + (size_t)q * (3 * (size_t)kk + 1);
    for (int i = tid; i < kk; i += blockDim.x) {
        double sv = l2 ? pos_inf() : -pos_inf(), fv = -pos_inf();
        long long id = -1;
        if (i < valid) {
            const long long t = ctag[i];
            sv = l2 ? ckey[i] : -ckey[i];
            id = p.id_base + (t >> TC_SLOT_BITS);
            fv = gfid[(int)(t & ((1 << TC_SLOT_BITS) - 1))];
        }
        rec[1 + i] = __double_as_longlong(sv);
        rec[1 + kk + i] = id;
        rec[1 + 2 * kk + i] = __double_as_longlong(fv);
    }
    if (tid == 0) {
        const long long bad = (p.cand_m[2 * q + 1] != 0 || valid > kk) ? 1 : 0;
        rec[0] = (long long)(valid < kk ? valid : kk) | (bad << 32);
    }
}

__global__ void __launch_bounds__(1024) tc_sort_kernel(const TcFinalParams p) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tc_sort_body(p, p.q0 + blockIdx.x, smem_raw);
}

__global__ void __launch_bounds__(1024) tc_sort_pack_kernel(const TcFinalParams p) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tc_sort_pack_body(p, p.q0 + blockIdx.x, smem_raw);
}

// (2) exact rescoring, grid (queries, y): CTA (q, y) takes the chunks y, y + gridDim.y, ... of XS_RESCORE_ROWS
// candidates of query q (y = 1 when the queries alone fill the machine: no CTAs that only find nothing to do).  A warp
// scores batches of XS_ROWS rows; lane l reads the row number of the warp's batch l >> 2, slot l & 3 of the chunk up
// front.  Same device code as the CUDA-core search (exact_score.cuh): a row's key does not depend on which kernel or
// batch scored it.
//
// Two forms.  tc_rescore_bulk_kernel (rows a multiple of 16 bytes and short enough for two CTAs per SM): a warp
// brings its batch into shared memory with one bulk async copy per row (UBLKCP, completion on the warp's own
// mbarrier) and scores it from there.  The gather is latency-bound when every load instruction waits for its own
// lines -- ncu r02 on the register form at k = 100: three exposed waits per batch (the compiler keeps 4 of the 12
// LDG.128 of a batch in flight at 80 registers), 66 % of the stall samples in that loop, 3.3 TB/s -- while a bulk
// copy has the whole batch (6 KB at D = 384) in flight per warp and needs no registers for it: 4 CTAs per SM,
// 32 batches = 192 KB in flight per SM, one wait per batch.  The first batch is requested before the query is staged.
// tc_rescore_kernel is the register form (any D, any alignment).
constexpr int XS_RESCORE_ROWS = 128;
constexpr int XS_RESCORE_BATCHES = XS_RESCORE_ROWS / (XS_WARPS * XS_ROWS);      // per warp and chunk
static_assert(XS_RESCORE_BATCHES * XS_ROWS <= 32, "one lane per row of the warp's share of a chunk");

__device__ __forceinline__ int rescore_fetch(const int* cand, int c, int m, int lane, int warp) {
    const int off = XS_ROWS * (warp + XS_WARPS * (lane >> 2)) + (lane & 3);     // this lane's row within a chunk
    return (lane < XS_RESCORE_BATCHES * XS_ROWS && c + off < m) ? cand[c + off] : -1;
}

__device__ __forceinline__ void rescore_emit(const TcFinalParams& p, int q, int r0, int c1, int lane, double tot, double nd2,
                                             double dot, double nq2) {
    if ((lane & 7) == 0) {
        const int r = r0 + (lane >> 3);
        if (r < c1) {
            p.cand_key[(size_t)q * p.cand_cap + r] = xs_key(p.metric, tot, nd2, nq2);
            // the rerank's fidelity from the same read of the row (bit-identical to amp_fidelity.cu)
            if (p.cand_fid != nullptr) p.cand_fid[(size_t)q * p.cand_cap + r] = xs_fidelity(dot, nd2, nq2);
        }
    }
}

static size_t rescore_bulk_smem(int D) {                       // staged query + reduction scratch + barriers + row stages
    return (size_t)(D + 2 * XS_WARPS) * 8 + (size_t)XS_WARPS * XS_ROWS * D * 4;
}

__global__ void __launch_bounds__(XS_THREADS, 4) tc_rescore_bulk_kernel(const TcFinalParams p) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int q = p.q0 + blockIdx.x;
    const int m = p.cand_m[2 * q];
    int cc = blockIdx.y * XS_RESCORE_ROWS;
    if (cc >= m && !p.fuse_sort) return;                        // (fused: an empty list still writes its padding)
    const int D = p.D;                                          // a multiple of 4 (the caller checks)
    double* qs = reinterpret_cast<double*>(smem_raw);
    double* red = qs + D;
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + XS_WARPS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* stage = reinterpret_cast<float*>(bars + XS_WARPS) + (size_t)warp * XS_ROWS * D;
    uint64_t* bar = bars + warp;
    const uint32_t row_bytes = (uint32_t)D * 4u;
    const bool l2 = p.metric == QRAG_METRIC_L2;
    const int* cand = p.cand_rows + (size_t)q * p.cand_cap;
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncwarp();
    int my = rescore_fetch(cand, cc, m, lane, warp);
    int c1 = (cc + XS_RESCORE_ROWS < m) ? cc + XS_RESCORE_ROWS : m;
    // lane 0 starts the copies of the warp's batch t of the chunk at cc (the stage is free: see the __syncwarp below)
    auto issue = [&](int t) {
        int id[XS_ROWS];
#pragma unroll
        for (int i = 0; i < XS_ROWS; ++i) id[i] = __shfl_sync(FULL_MASK, my, XS_ROWS * t + i);
        if (lane == 0) {
            const int left = c1 - (cc + XS_ROWS * (warp + XS_WARPS * t));
            const int n = left < XS_ROWS ? left : XS_ROWS;
            mbar_arrive_expect_tx(bar, (uint32_t)n * row_bytes);
#pragma unroll
            for (int i = 0; i < XS_ROWS; ++i)
                if (i < n) bulk_g2s(stage + (size_t)i * D, p.X + (size_t)id[i] * D, row_bytes, bar);
        }
    };
    bool ahead = cc + XS_ROWS * warp < c1;                      // the first batch is requested before the query is staged
    if (ahead) issue(0);
    const double nq2 = xs_stage_query(p.Q + (size_t)q * D, D, qs, red);     // ends with __syncthreads()
    uint32_t parity = 0;
    while (cc < m) {
#pragma unroll 1
        for (int t = 0; t < XS_RESCORE_BATCHES; ++t) {
            const int r0 = cc + XS_ROWS * (warp + XS_WARPS * t);
            if (r0 >= c1) break;                                // warp-uniform
            if (!ahead) issue(t);
            ahead = false;
            mbar_wait(bar, parity);
            parity ^= 1u;
            const float* rp[XS_ROWS];
#pragma unroll
            for (int i = 0; i < XS_ROWS; ++i) rp[i] = stage + (size_t)(r0 + i < c1 ? i : 0) * D;   // past the end: row 0 again, unused
            double nd2, dot, tot;
            if (l2 && p.cand_fid != nullptr) {
                tot = xs_score4_l2dot<true, false>(rp, qs, D, lane, nd2, dot);
            } else {
                tot = xs_score4<true, false>(rp, qs, D, l2, lane, nd2);
                dot = tot;
            }
            rescore_emit(p, q, r0, c1, lane, tot, nd2, dot, nq2);
            __syncwarp();                                       // every lane has read the stage before it is refilled
        }
        cc += gridDim.y * XS_RESCORE_ROWS;
        if (cc >= m) break;                                     // CTA-uniform
        c1 = (cc + XS_RESCORE_ROWS < m) ? cc + XS_RESCORE_ROWS : m;
        my = rescore_fetch(cand, cc, m, lane, warp);
    }
    if (p.fuse_sort) {
        // this CTA scored the query's whole list (grid.y == 1): sort it here, in the row stages (every copy into them
        // has been waited for), instead of in a kernel of its own
        __syncthreads();
        unsigned char* sort_smem = reinterpret_cast<unsigned char*>(bars + XS_WARPS);
        if (p.pack != nullptr) tc_sort_pack_body(p, q, sort_smem);
        else                   tc_sort_body(p, q, sort_smem);
    }
}

template <bool VEC>
__global__ void __launch_bounds__(XS_THREADS) tc_rescore_kernel(const TcFinalParams p) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int q = p.q0 + blockIdx.x;
    const int m = p.cand_m[2 * q];
    int cc = blockIdx.y * XS_RESCORE_ROWS;
    if (cc >= m) return;
    const int D = p.D, Dpad = (D + 3) & ~3;
    double* qs = reinterpret_cast<double*>(smem_raw);
    double* red = qs + Dpad;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool l2 = p.metric == QRAG_METRIC_L2;
    const int* cand = p.cand_rows + (size_t)q * p.cand_cap;
    int my = rescore_fetch(cand, cc, m, lane, warp);
    const double nq2 = xs_stage_query(p.Q + (size_t)q * D, D, qs, red);     // ends with __syncthreads()
    while (true) {
        const int c1 = (cc + XS_RESCORE_ROWS < m) ? cc + XS_RESCORE_ROWS : m;
#pragma unroll 1
        for (int t = 0; t < XS_RESCORE_BATCHES; ++t) {
            const int r0 = cc + XS_ROWS * (warp + XS_WARPS * t);
            if (r0 >= c1) break;                               // warp-uniform
            const float* rp[XS_ROWS];
            const int id0 = __shfl_sync(FULL_MASK, my, XS_ROWS * t);
#pragma unroll
            for (int i = 0; i < XS_ROWS; ++i) {
                int id = __shfl_sync(FULL_MASK, my, XS_ROWS * t + i);
                if (id < 0) id = id0;                          // past the end of the list: score row r0 again, unused
                rp[i] = p.X + (size_t)id * D;
            }
            double nd2, dot, tot;
            if (l2 && p.cand_fid != nullptr) {
                tot = xs_score4_l2dot<VEC>(rp, qs, D, lane, nd2, dot);
            } else {
                tot = xs_score4<VEC>(rp, qs, D, l2, lane, nd2);
                dot = tot;
            }
            rescore_emit(p, q, r0, c1, lane, tot, nd2, dot, nq2);
        }
        cc += gridDim.y * XS_RESCORE_ROWS;
        if (cc >= m) break;                                    // CTA-uniform
        my = rescore_fetch(cand, cc, m, lane, warp);
    }
}

// Owner side of the search + rerank exchange: one CTA per owned query.  recv [G, per, 3 kk + 1] holds, from every
// shard, the record tc_sort_pack_kernel wrote for this query.  The G sorted lists are merged by RANK -- an entry's
// position in the global (score, id) order is its own position plus, for every other list, the number of entries
// that sort before it (binary search) -- so nothing is moved; entries with rank < k1 are the global top-k1, and
// they are ordered by (fidelity desc, rank asc), the reference's stable sort (quantum.py:70-76).
// out [per, 2 k2 + 1]: k2 fidelity bits, k2 ids, status (some shard could not certify the query / cut its list).
struct OwnerParams {
    const long long* recv; int G, per, kk, k1, k2, l2, q_base, nq;
    long long* out;
};
constexpr int OWNER_MAX_G = 256;
__global__ void __launch_bounds__(1024) owner_finalize_kernel(const OwnerParams p) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_valid[OWNER_MAX_G];
    __shared__ int s_bad, s_total;
    const int tid = threadIdx.x;
    const int j = blockIdx.x, q = p.q_base + j;
    const int G = p.G, kk = p.kk, k1 = p.k1, k2 = p.k2;
    const size_t rec = 3 * (size_t)kk + 1;
    long long* out = p.out + (size_t)j * (2 * (size_t)k2 + 1);
    if (q >= p.nq) {                                            // padding query of the last owner
        for (int i = tid; i < k2; i += blockDim.x) { out[i] = __double_as_longlong(-pos_inf()); out[k2 + i] = -1; }
        if (tid == 0) out[2 * k2] = 0;
        return;
    }
    if (tid == 0) { s_bad = 0; s_total = 0; }
    __syncthreads();
    for (int g = tid; g < G; g += blockDim.x) {
        const long long h = p.recv[((size_t)g * p.per + j) * rec];
        int v = (int)(h & 0xffffffffLL);
        if (v < 0) v = 0;
        if (v > kk) v = kk;
        s_valid[g] = v;
        atomicAdd(&s_total, v);
        if (h >> 32) s_bad = 1;
    }
    __syncthreads();
    int members = s_total < k1 ? s_total : k1;
    int P = 1;
    while (P < members) P <<= 1;
    double* key = reinterpret_cast<double*>(smem_raw);                      // [G kk] ascending = better
    long long* ids = reinterpret_cast<long long*>(key + (size_t)G * kk);    // [G kk]
    double* mkey = reinterpret_cast<double*>(ids + (size_t)G * kk);         // [P] -fidelity of the member at a rank
    int* mtag = reinterpret_cast<int*>(mkey + P);                           // [P] the rank (sort tag)
    int* mslot = mtag + P;                                                  // [P] rank -> entry
    const int total_slots = G * kk;
    for (int e = tid; e < total_slots; e += blockDim.x) {
        const int g = e / kk, i = e - g * kk;
        if (i < s_valid[g]) {
            const long long* r = p.recv + ((size_t)g * p.per + j) * rec;
            const double sv = __longlong_as_double(r[1 + i]);
            key[e] = p.l2 ? sv : -sv;
            ids[e] = r[1 + kk + i];
        }
    }
    // every rank slot starts empty: with well-formed input (ids unique over the lists) ranks 0 .. members-1 are each
    // written exactly once below; malformed input leaves holes that come out as padding instead of wild reads
    for (int i = tid; i < P; i += blockDim.x) { mkey[i] = pos_inf(); mtag[i] = 0x7fffffff; mslot[i] = -1; }
    __syncthreads();
    for (int e = tid; e < total_slots; e += blockDim.x) {
        const int g = e / kk, i = e - g * kk;
        if (i >= s_valid[g]) continue;
        const double ke = key[e];
        const long long ie = ids[e];
        int rank = i;
        for (int h = 0; h < G; ++h) {
            if (h == g) continue;
            int lo = 0, hi = s_valid[h];                         // first entry of list h that does NOT sort before e
            const double* kh = key + (size_t)h * kk;
            const long long* ih = ids + (size_t)h * kk;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const double km = kh[mid];
                if (km < ke || (km == ke && ih[mid] < ie)) lo = mid + 1;
                else hi = mid;
            }
            rank += lo;
        }
        if (rank < members) {
            const double f = __longlong_as_double(p.recv[((size_t)g * p.per + j) * rec + 1 + 2 * kk + i]);
            mkey[rank] = -f;
            mtag[rank] = rank;
            mslot[rank] = e;
        }
    }
    __syncthreads();
    const int nwarps = blockDim.x >> 5, lane = tid & 31, warp = tid >> 5;
    if (k2 <= nwarps && P <= (int)blockDim.x) {
        // Few winners out of many members: no sort.  The k2-th largest of the per-warp maxima is a lower bound of the
        // k2-th best fidelity (k2 warps each hold a member at least that good), so only members reaching it -- a few
        // dozen -- can be winners; they are compacted and ranked among themselves by counting.  4 block barriers
        // instead of the ~55 of a bitonic sort of 1024 (ncu: the sort's barriers were half of this kernel's stalls).
        __shared__ double s_wmax[32];
        __shared__ double s_low;
        __shared__ int s_ns;
        double* skey = reinterpret_cast<double*>(mslot + P);                // [P] survivors' keys (space: see host side)
        int* stag = reinterpret_cast<int*>(skey + P);                      // [P] survivors' ranks in the merged list
        const double mine = tid < P ? mkey[tid] : pos_inf();               // -fidelity (ascending = better), +inf = empty
        double wbest = mine;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wbest = fmin(wbest, __shfl_xor_sync(FULL_MASK, wbest, o));
        if (lane == 0) s_wmax[warp] = wbest;
        if (tid == 0) s_ns = 0;
        __syncthreads();
        if (warp == 0) {
            const double v = lane < nwarps ? s_wmax[lane] : pos_inf();
            int better = 0;                                                // warps whose best beats this warp's best
            for (int j = 0; j < 32; ++j) {
                const double o = __shfl_sync(FULL_MASK, v, j);
                better += (o < v) || (o == v && j < lane);
            }
            if (better == k2 - 1) s_low = v;                               // exactly one lane: the k2-th best warp maximum
        }
        __syncthreads();
        const double low = s_low;
        if (tid < P && mine <= low && mtag[tid] != 0x7fffffff) {
            const int at = atomicAdd(&s_ns, 1);
            skey[at] = mine;
            stag[at] = mtag[tid];
        }
        __syncthreads();
        const int ns = s_ns;
        for (int i = tid; i < ns; i += blockDim.x) {
            const double ki = skey[i];
            const int ti = stag[i];
            int rank = 0;
            for (int j = 0; j < ns; ++j) {
                const double kj = skey[j];
                rank += (kj < ki) || (kj == ki && stag[j] < ti);
            }
            if (rank < k2) {
                out[rank] = __double_as_longlong(-ki);
                out[k2 + rank] = ids[mslot[ti]];
            }
        }
        for (int i = ns + tid; i < k2; i += blockDim.x) {                  // fewer members than k2: padding
            out[i] = __double_as_longlong(-pos_inf());
            out[k2 + i] = -1;
        }
    } else {
        bitonic_sort_kt<double, int>(mkey, mtag, P);
        for (int i = tid; i < k2; i += blockDim.x) {
            double f = -pos_inf();
            long long id = -1;
            if (i < members && mtag[i] != 0x7fffffff) { f = -mkey[i]; id = ids[mslot[mtag[i]]]; }
            out[i] = __double_as_longlong(f);
            out[k2 + i] = id;
        }
    }
    if (tid == 0) out[2 * k2] = s_bad;
}

// --------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// [rows, Kp] bf16 row-major, box = 64 x box_rows, 128-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int Kp, int box_rows, bool fp16) {
    EncodeTiledFn fn = encode_tiled_fn();
    QRAG_REQUIRE(fn != nullptr, QRAG_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    QRAG_REQUIRE(r == CUDA_SUCCESS, QRAG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld Kp=%d", (int)r,
                 (long long)rows, Kp);
    return QRAG_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Every per-batch kernel of the search is launched with programmatic stream serialization and begins with pdl_enter()
// (below): its CTAs may become resident -- launch latency, scheduling, and in the GEMM the barrier / TMEM set-up --
// while the kernel before it drains, and touch nothing that kernel wrote until it has completed.  A batch is 7-8
// short dependent kernels around one long one; the boundaries were ~20 us of it (a CUDA-graph replay measured 3 %
// faster than plain launches; this gets the same without asking the caller for fixed pointers).
template <class... P, class... A>
static cudaError_t launch_chained(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#ifdef QRAG_TUNING
    if (const char* e = getenv("QRAG_TC_TUNE")) if (atoi(e) & 8) cfg.numAttrs = 0;
#endif
    return cudaLaunchKernelEx(&cfg, kern, P(args)...);
}

static int tc_kp(int D, int metric) { return (int)align_up((size_t)D + (metric == QRAG_METRIC_L2 ? 2 : 0), 16); }

struct TcPlan {
    int Kp, kchunks, ksteps_last, stages, a_resident, stage_bytes, cg, nq_pad, groups, ntiles, sample, nsample_tiles, nbuckets;
    int cand_cap, cap;
    int kt;                 // entries of the threshold lists the shards exchange (tc_exchange_len)
    int fp16;               // 16-bit operand format of this metric (tc_operand_fp16)
    size_t smem_gemm, smem_final;
    size_t off_qb, off_qnorm, off_bmax, off_tau, off_eps, off_hinv, off_hist, off_cnt, off_surv, off_crow, off_ckey,
        off_cfid, off_cm, total;
};

// Lists exchanged between the shards of a G-way search travel cut to this many entries.  A shard holds ~ k / G of
// the global top-k; the k-th largest of the union of the G lists' first kt entries is still a valid threshold (every
// entry is a distinct document, so at least k documents reach it) and equals the exact one unless a single shard
// holds more than kt of the top k -- then it is merely lower (more survivors, never a wrong result).
static int tc_exchange_len(int k, int shards) {
    if (shards <= 1) return k;
    const int64_t kt = (2 * (int64_t)k / shards + 64 + 31) / 32 * 32;
    return kt < k ? (int)kt : k;
}

static int tc_sm_units(int cg);

static int tc_plan(int nq, int64_t N, int D, int k, int metric, int shards, TcPlan* pl) {
    QRAG_REQUIRE(nq >= 0 && N >= 0 && D > 0, QRAG_ERR_INVALID, "bad sizes nq=%d N=%lld D=%d", nq, (long long)N, D);
    QRAG_REQUIRE(metric >= 0 && metric <= 2, QRAG_ERR_INVALID, "unknown metric %d", metric);
    QRAG_REQUIRE(k >= 1 && k <= 2048, QRAG_ERR_UNSUPPORTED, "tensor-core search supports 1 <= k <= 2048 (got %d)", k);
    QRAG_REQUIRE(N < ((int64_t)1 << 31) - TC_BN, QRAG_ERR_UNSUPPORTED, "shard too large (N=%lld)", (long long)N);
    const DeviceProps& dp = device_props();
    QRAG_REQUIRE(dp.ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    QRAG_REQUIRE(dp.cc_major == 10, QRAG_ERR_UNSUPPORTED, "tcgen05 search needs compute capability 10.x (got %d.%d)",
                 dp.cc_major, dp.cc_minor);
    pl->Kp = tc_kp(D, metric);
    pl->kchunks = (pl->Kp + TC_BK - 1) / TC_BK;
    pl->ksteps_last = (pl->Kp - (pl->kchunks - 1) * TC_BK) / TC_UK;
    // the query tile stays resident when that still leaves a 3-deep ring for the document chunks;
    // otherwise (long rows) its chunks travel through the ring next to them
    const size_t a_bytes = (size_t)pl->kchunks * TC_A_CHUNK;
    const size_t budget = (size_t)dp.max_smem_optin;
    pl->a_resident = (1024 + 256 + a_bytes + 3 * (size_t)TC_B_STAGE <= budget) ? 1 : 0;
    // CTA pairs (cta_group::2) whenever there are at least two query groups and the query tile is resident
    pl->cg = (pl->a_resident && nq > TC_BM) ? 2 : 1;
#ifdef QRAG_TUNING
    if (getenv("QRAG_TC_NO_PAIR")) pl->cg = 1;
#endif
    pl->stage_bytes = pl->a_resident ? TC_B_STAGE / pl->cg : TC_B_STAGE + TC_A_CHUNK;
    const size_t fixed = 1024 + 256 + (pl->a_resident ? a_bytes : 0);
    int stages = (int)((budget - fixed) / pl->stage_bytes);
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    QRAG_REQUIRE(stages >= 2, QRAG_ERR_UNSUPPORTED, "not enough shared memory for the tensor-core search");
    pl->stages = stages;
    pl->smem_gemm = fixed + (size_t)stages * pl->stage_bytes;
    pl->nq_pad = (int)align_up((size_t)(nq > 0 ? nq : 1), (size_t)TC_BM * pl->cg);
    pl->groups = pl->nq_pad / TC_BM;
    pl->ntiles = (int)ceil_div(N > 0 ? N : 1, TC_BN);
    int cap = TC_CAP_MIN;
    while (cap < 64 * k && cap < TC_CAP_MAX) cap <<= 1;       // sized for the expected survivors (the sample rate below uses it)
    {
        // ... and for CLUSTERED corpora: the list is split into one segment per (CTA of the query's group, column
        // quarter), and near-duplicates stored in adjacent rows put most of a query's neighbours into one 256-row
        // tile, i.e. into one CTA's four segments.  Give every segment room for 256 survivors where memory allows
        // (few queries = many CTAs per group = many short segments: 27 slots each before this).
        const int su = tc_sm_units(pl->cg);
        int ug = pl->groups / pl->cg;
        if (ug > su) ug = su;
        const int nseg = TC_EPI_SPLIT * (su / (ug > 0 ? ug : 1));
        int64_t want = next_pow2((int64_t)256 * nseg);
        while (want > cap && (int64_t)pl->nq_pad * want * 8 > ((int64_t)1 << 30)) want >>= 1;   // at most 1 GiB of lists
        pl->cap = want > cap ? (int)want : cap;
    }
    // pass 1 samples every `sample`-th tile: survivors ~ sample * k per query, kept well under the capacity,
    // and the sample must hold many more buckets than k for the bound to be tight
    // (the threshold comes from the union of all shards' samples: a shard of a G-way search expects sample * k / G)
    QRAG_REQUIRE(shards >= 1 && shards <= 1024, QRAG_ERR_INVALID, "shards=%d", shards);
    int64_t sample = (int64_t)cap * shards / (4 * (int64_t)k);
    // every 16th tile at most, whatever the shard count: pass 1 costs ~ (filter pass) / sample, the survivors it leaves
    // the filter pass cost ~ 0.002 ms per unit of `sample` (k = 1000, measured per stage with tools/shard_stage_probe.py
    // --samples at G = 1, 2, 4, 8): the sum is flat between 12 and 24 and rises beyond (G = 4: 1.85 ms at 16, 1.93 at
    // 32, 2.06 at 64; round 2 sampled G x more sparsely on a G-way shard)
    if (sample > 16) sample = 16;
    const int64_t by_buckets = N * shards / ((int64_t)256 * k);
    if (sample > by_buckets) sample = by_buckets;
    if (sample < 1) sample = 1;
#ifdef QRAG_TUNING
    if (const char* e = getenv("QRAG_TC_SAMPLE")) sample = atoi(e) > 0 ? atoi(e) : sample;
#endif
    pl->sample = (int)sample;
    const int sample_i = (int)sample;
    pl->nsample_tiles = (pl->ntiles + sample_i - 1) / sample_i;
    pl->nbuckets = pl->nsample_tiles * (TC_BN / TC_BUCKET);
    // rescoring candidates per query: k + the 2-eps margin; a shard of a G-way search expects (k + margin) / G of
    // them, so its lists (and the shared memory that bounds the resident CTAs) are sized for 4x that share -- a
    // shard that holds more flags the query (status) and the caller reruns it exactly
    int cand_cap = next_pow2(2 * (int64_t)k + 512);
    if (shards > 1) {
        const int share = next_pow2(4 * (int64_t)k / shards + 512);
        if (share < cand_cap) cand_cap = share;
    }
    if (cand_cap > TC_MAX_CAND) cand_cap = TC_MAX_CAND;
    pl->cand_cap = cand_cap;
    pl->kt = tc_exchange_len(k, shards);
    pl->fp16 = tc_operand_fp16(metric) ? 1 : 0;
    const int Dpad = (D + 3) & ~3;
    pl->smem_final = (size_t)(Dpad + XS_WARPS) * 8;           // tc_rescore: the staged query (tc_sort: 12 B per candidate)
    QRAG_REQUIRE(pl->smem_final <= budget, QRAG_ERR_UNSUPPORTED, "D=%d too large for the rescoring stage", D);
    size_t off = 0;
    pl->off_qb = off; off = align_up(off + (size_t)pl->nq_pad * pl->Kp * 2, 256);
    pl->off_qnorm = off; off = align_up(off + (size_t)pl->nq_pad * 8, 256);      // |q| and |q - bf16(q)| per query
    pl->off_bmax = off; off = align_up(off + (size_t)pl->nq_pad * pl->nbuckets * 4, 256);
    pl->off_tau = off; off = align_up(off + (size_t)pl->nq_pad * 4, 256);
    pl->off_eps = off; off = align_up(off + (size_t)pl->nq_pad * 4, 256);
    pl->off_hinv = off; off = align_up(off + (size_t)pl->nq_pad * 4, 256);
    pl->off_hist = off; off = align_up(off + (size_t)pl->nq_pad * TC_HIST_BINS * 4, 256);
    pl->off_cnt = off; off = align_up(off + (size_t)pl->nq_pad * TC_MAX_SEGS * 4, 256);
    pl->off_surv = off; off = align_up(off + (size_t)pl->nq_pad * pl->cap * 8, 256);
    pl->off_crow = off; off = align_up(off + (size_t)pl->nq_pad * cand_cap * 4, 256);
    pl->off_ckey = off; off = align_up(off + (size_t)pl->nq_pad * cand_cap * 8, 256);
    pl->off_cfid = off; off = align_up(off + (size_t)pl->nq_pad * cand_cap * 8, 256);
    pl->off_cm = off; off = align_up(off + (size_t)pl->nq_pad * 2 * 4, 256);
    pl->total = off + 256;
    return QRAG_OK;
}

// workspace carved the same way by every phase of one search (the phases share state through it)
struct TcWs {
    TcPlan pl;
    unsigned short* Qb; float* qnorm; float* bmax; float* tau; float* eps; float* hinv; int* hist; unsigned int* cnt;
    float2* surv;
    int* crow; double* ckey; double* cfid; int* cm;
    float* dump = nullptr;  // qrag_search_tc_scores only
};

static int tc_ws(int nq, int64_t N, int D, int k, int metric, int shards, void* workspace, size_t workspace_bytes, TcWs* w) {
    int rc = tc_plan(nq, N, D, k, metric, shards, &w->pl);
    if (rc) return rc;
    QRAG_REQUIRE(N >= 1, QRAG_ERR_INVALID, "empty shard: nothing to search");
    QRAG_REQUIRE(workspace != nullptr && workspace_bytes >= w->pl.total, QRAG_ERR_WORKSPACE,
                 "workspace too small: need %zu bytes, got %zu", w->pl.total, workspace_bytes);
    unsigned char* ws = reinterpret_cast<unsigned char*>(align_up((size_t)workspace, 256));
    const TcPlan& pl = w->pl;
    w->Qb = reinterpret_cast<unsigned short*>(ws + pl.off_qb);
    w->qnorm = reinterpret_cast<float*>(ws + pl.off_qnorm);
    w->bmax = reinterpret_cast<float*>(ws + pl.off_bmax);
    w->tau = reinterpret_cast<float*>(ws + pl.off_tau);
    w->eps = reinterpret_cast<float*>(ws + pl.off_eps);
    w->hinv = reinterpret_cast<float*>(ws + pl.off_hinv);
    w->hist = reinterpret_cast<int*>(ws + pl.off_hist);
    w->cnt = reinterpret_cast<unsigned int*>(ws + pl.off_cnt);
    w->surv = reinterpret_cast<float2*>(ws + pl.off_surv);
    w->crow = reinterpret_cast<int*>(ws + pl.off_crow);
    w->ckey = reinterpret_cast<double*>(ws + pl.off_ckey);
    w->cfid = reinterpret_cast<double*>(ws + pl.off_cfid);
    w->cm = reinterpret_cast<int*>(ws + pl.off_cm);
    return QRAG_OK;
}

template <int MODE, int CG>
static int launch_gemm(const CUtensorMap& mapA, const CUtensorMap& mapB, const TcGemmParams& gp, size_t smem, int grid,
                       cudaStream_t st) {
    auto kern = sim_gemm_kernel<MODE, CG>;
    QRAG_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // see launch_chained
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = CG;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CG > 1 ? 2 : 1;
#ifdef QRAG_TUNING
    if (const char* e = getenv("QRAG_TC_TUNE")) if (atoi(e) & 8) { cfg.attrs = attr + 1; cfg.numAttrs = CG > 1 ? 1 : 0; }
#endif
    QRAG_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, mapA, mapB, gp));
    QRAG_LAUNCH_CHECK("sim_gemm_kernel");
    return QRAG_OK;
}

// How the query groups [g0, g0 + groups) of one launch share the SMs: `groups` groups (a multiple of cg), each
// visited by `cpg` CTAs (pairs), i.e. cpg * TC_EPI_SPLIT survivor segments per query.
struct TcLaunch { int g0, groups, cpg; };

static int tc_sm_units(int cg) {
    int n = device_props().sm_count / cg;
    const int cap = TC_MAX_SEGS / TC_EPI_SPLIT;
    return n < cap ? n : cap;
}

static TcLaunch tc_launch_at(const TcPlan& pl, int g0, int units) {
    const int su = tc_sm_units(pl.cg);                       // CTAs (cg 1) or pairs (cg 2) available
    int ug = (pl.groups - g0) / pl.cg;                       // group units left
    if (ug > su) ug = su;
    int cpg = su / ug;
    if (cpg > units) cpg = units;
    return TcLaunch{g0, ug * pl.cg, cpg};
}

// One GEMM pass over the shard for every query group; `after` runs once per launch (query range).
template <int MODE, class After>
static int tc_gemm_pass(const TcWs& w, int nq, int64_t N, const uint16_t* Xb, cudaStream_t st, After after) {
    const TcPlan& pl = w.pl;
    CUtensorMap mapA, mapB;
    int rc = make_map(&mapA, w.Qb, pl.nq_pad, pl.Kp, TC_BM, pl.fp16 != 0);
    if (rc) return rc;
    rc = make_map(&mapB, Xb, N, pl.Kp, TC_BN / pl.cg, pl.fp16 != 0);
    if (rc) return rc;
    TcGemmParams gp{};
    gp.kchunks = pl.kchunks; gp.ksteps_last = pl.ksteps_last; gp.stages = pl.stages;
    gp.a_resident = pl.a_resident; gp.stage_bytes = pl.stage_bytes; gp.fp16 = pl.fp16;
    gp.nq = nq; gp.N = N; gp.ntiles = pl.ntiles; gp.sample = pl.sample; gp.nbuckets = pl.nbuckets;
    gp.tau = w.tau; gp.bmax = w.bmax; gp.cnt = w.cnt; gp.surv = w.surv; gp.dump = w.dump;
#ifdef QRAG_TUNING
    if (const char* e = getenv("QRAG_TC_DEBUG_SKIP")) gp.debug_skip = atoi(e);
#endif
    const int units = MODE == TC_MODE_BUCKET ? pl.nsample_tiles : pl.ntiles;
    for (int g0 = 0; g0 < pl.groups;) {
        const TcLaunch L = tc_launch_at(pl, g0, units);
        gp.groups = L.groups;
        gp.group0 = L.g0;
        gp.cap = pl.cap;
        gp.seg_cap = pl.cap / (TC_EPI_SPLIT * L.cpg);
        const int grid = L.groups * L.cpg;
        rc = pl.cg == 2 ? launch_gemm<MODE, 2>(mapA, mapB, gp, pl.smem_gemm, grid, st)
                        : launch_gemm<MODE, 1>(mapA, mapB, gp, pl.smem_gemm, grid, st);
        if (rc) return rc;
        const int q0 = g0 * TC_BM;
        const int q1 = (g0 + L.groups) * TC_BM < nq ? (g0 + L.groups) * TC_BM : nq;
        if (q1 > q0) {
            rc = after(q0, q1, TC_EPI_SPLIT * L.cpg, gp.seg_cap);
            if (rc) return rc;
        }
        g0 += L.groups;
    }
    return QRAG_OK;
}

}  // namespace qrag

using namespace qrag;

extern "C" int qrag_index_prepared_dims(int D, int metric, int* Kp) {
    QRAG_REQUIRE(Kp != nullptr && D > 0 && metric >= 0 && metric <= 2, QRAG_ERR_INVALID, "bad arguments");
    *Kp = tc_kp(D, metric);
    return QRAG_OK;
}

extern "C" int qrag_index_prepare(const float* X, int64_t N, int D, int metric, uint16_t* Xb, float* aux, void* stream) {
    QRAG_REQUIRE(Xb && aux && (X || N == 0), QRAG_ERR_INVALID, "null pointer argument");
    QRAG_REQUIRE(N >= 0 && D > 0 && metric >= 0 && metric <= 2, QRAG_ERR_INVALID, "bad arguments");
    QRAG_REQUIRE(device_props().ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    cudaStream_t st = (cudaStream_t)stream;
    QRAG_CUDA_CHECK(cudaMemsetAsync(aux, 0, 4 * sizeof(float), st));
    if (N == 0) return QRAG_OK;
    const int Kp = tc_kp(D, metric);
    index_prepare_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, st>>>(X, N, D, Kp, metric, reinterpret_cast<unsigned short*>(Xb),
                                                                  aux);
    QRAG_LAUNCH_CHECK("index_prepare_kernel");
    return QRAG_OK;
}

extern "C" int qrag_search_tc_workspace(int nq, int64_t N, int D, int k, int metric, int shards, size_t* bytes) {
    QRAG_REQUIRE(bytes != nullptr, QRAG_ERR_INVALID, "bytes is null");
    TcPlan pl;
    int rc = tc_plan(nq, N, D, k, metric, shards, &pl);
    if (rc) return rc;
    *bytes = pl.total;
    return QRAG_OK;
}

// phase 1: operands, sampled bucket-maximum GEMM, the shard's k largest bucket maxima
extern "C" int qrag_search_tc_begin(const float* Q, int nq, const uint16_t* Xb, int64_t N, int D, int k, int metric,
                                    int shards, float* bm_top, void* workspace, size_t workspace_bytes, void* stream) {
    QRAG_REQUIRE(Q && Xb && (bm_top || shards == 1), QRAG_ERR_INVALID, "null pointer argument");
    QRAG_REQUIRE((uintptr_t)Xb % 16 == 0, QRAG_ERR_INVALID, "Xb must be 16-byte aligned");
    if (nq == 0) return QRAG_OK;
    TcWs w;
    int rc = tc_ws(nq, N, D, k, metric, shards, workspace, workspace_bytes, &w);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const TcPlan& pl = w.pl;
    QRAG_CUDA_CHECK(launch_chained(query_prepare_kernel, dim3((unsigned)ceil_div(pl.nq_pad, 8)), dim3(256), 0, st, Q, nq, pl.nq_pad, D, pl.Kp, metric, w.Qb, w.qnorm,
                                   w.qnorm + pl.nq_pad));
    QRAG_LAUNCH_CHECK("query_prepare_kernel");
    rc = tc_gemm_pass<TC_MODE_BUCKET>(w, nq, N, Xb, st, [](int, int, int, int) { return QRAG_OK; });
    if (rc) return rc;
    if (bm_top != nullptr) {                                   // single shard: the threshold kernel reads bmax itself
        QRAG_CUDA_CHECK(launch_chained(bucket_topk_kernel, dim3(nq), dim3(256), 0, st, w.bmax, pl.nbuckets, pl.kt, bm_top));
        QRAG_LAUNCH_CHECK("bucket_topk_kernel");
    }
    return QRAG_OK;
}

// diagnostic: the approximate score matrix of the filter GEMM, i.e. what tc_eps is a bound on
extern "C" int qrag_search_tc_scores(const float* Q, int nq, const uint16_t* Xb, int64_t N, int D, int metric, float* out,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    QRAG_REQUIRE(Q && Xb && out, QRAG_ERR_INVALID, "null pointer argument");
    QRAG_REQUIRE((uintptr_t)Xb % 16 == 0, QRAG_ERR_INVALID, "Xb must be 16-byte aligned");
    if (nq == 0) return QRAG_OK;
    TcWs w;
    int rc = tc_ws(nq, N, D, 1, metric, 1, workspace, workspace_bytes, &w);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const TcPlan& pl = w.pl;
    QRAG_CUDA_CHECK(launch_chained(query_prepare_kernel, dim3((unsigned)ceil_div(pl.nq_pad, 8)), dim3(256), 0, st, Q, nq, pl.nq_pad, D, pl.Kp, metric, w.Qb, w.qnorm,
                                   w.qnorm + pl.nq_pad));
    QRAG_LAUNCH_CHECK("query_prepare_kernel");
    w.dump = out;
    return tc_gemm_pass<TC_MODE_DUMP>(w, nq, N, Xb, st, [](int, int, int, int) { return QRAG_OK; });
}

// phase 2: threshold from all shards' bucket maxima, filter GEMM, then this shard's survivors counted into the
// per-query score histogram (hist [nq, QRAG_TC_HIST_BINS], or the workspace's own when hist == nullptr)
extern "C" int qrag_search_tc_filter(int nq, const uint16_t* Xb, const float* aux, int64_t N, int D, int k, int metric,
                                     const float* bm_top_all, int G, int32_t* hist, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    QRAG_REQUIRE(Xb && aux && G >= 1 && (bm_top_all || G == 1), QRAG_ERR_INVALID, "bad argument");
    if (nq == 0) return QRAG_OK;
    TcWs w;
    int rc = tc_ws(nq, N, D, k, metric, G, workspace, workspace_bytes, &w);
    if (rc) return rc;
    if (hist != nullptr) w.hist = hist;
    cudaStream_t st = (cudaStream_t)stream;
    QRAG_CUDA_CHECK(launch_chained(tau_union_kernel, dim3(nq), dim3(256), 0, st, bm_top_all, G, nq, k, w.pl.kt, metric, w.pl.Kp, w.qnorm, w.qnorm + w.pl.nq_pad, aux,
                                   w.bmax, w.pl.nbuckets, w.tau, w.eps, w.hinv));
    QRAG_LAUNCH_CHECK("tau_union_kernel");
    return tc_gemm_pass<TC_MODE_FILTER>(w, nq, N, Xb, st, [&](int q0, int q1, int nseg, int seg_cap) {
        QRAG_CUDA_CHECK(launch_chained(surv_hist_kernel, dim3(q1 - q0), dim3(seg_walk_threads(nseg)), 0, st, w.cnt, w.surv, q0, nseg, seg_cap, w.pl.cap, w.tau, w.hinv, w.hist));
        QRAG_LAUNCH_CHECK("surv_hist_kernel");
        return QRAG_OK;
    });
}

// phase 3: candidates against the global k-th best approximate score, exact rescoring, sorted shard list --
// either as (scores, ids, status) arrays or, packed != nullptr, as the per-query records of the search + rerank
// exchange (tc_sort_pack_kernel), which also carry the amplitude fidelity of every entry.
static int tc_finish_impl(const float* Q, int nq, const float* X, int64_t N, int D, int k, int metric, int64_t id_base,
                          const int32_t* hist_all, int G, double* out_scores, int64_t* out_ids, int32_t* status,
                          long long* pack, int kk, void* workspace, size_t workspace_bytes, void* stream) {
    if (nq == 0) return QRAG_OK;
    TcWs w;
    int rc = tc_ws(nq, N, D, k, metric, G, workspace, workspace_bytes, &w);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const TcPlan& pl = w.pl;
    const bool vec = (D % 4 == 0) && ((uintptr_t)X % 16 == 0);
    for (int g0 = 0; g0 < pl.groups;) {                        // same launch partition as the filter pass
        const TcLaunch L = tc_launch_at(pl, g0, pl.ntiles);
        const int q0 = g0 * TC_BM;
        const int q1 = (g0 + L.groups) * TC_BM < nq ? (g0 + L.groups) * TC_BM : nq;
        g0 += L.groups;
        if (q1 <= q0) continue;
        TcFinalParams fp{Q, X, q0, nq, N, D, k, metric, id_base, TC_EPI_SPLIT * L.cpg, pl.cap / (TC_EPI_SPLIT * L.cpg),
                         pl.cap, pl.cand_cap, hist_all ? hist_all : w.hist, w.tau, w.hinv, w.cnt, w.surv, w.eps, out_scores,
                         out_ids, status,
                         w.crow, w.ckey, w.cm, pack ? w.cfid : nullptr, pack, kk};
#ifdef QRAG_TUNING
        fp.tune = getenv("QRAG_TC_TUNE") ? atoi(getenv("QRAG_TC_TUNE")) : 0;
#endif
        QRAG_CUDA_CHECK(launch_chained(tc_collect_kernel, dim3(q1 - q0), dim3(seg_walk_threads(fp.nseg)), 0, st, fp));
        QRAG_LAUNCH_CHECK("tc_collect_kernel");
        // the bulk-copy form when the rows allow it and two CTAs still share an SM
        const size_t smem_bulk = rescore_bulk_smem(D);
        bool bulk = vec && 2 * (smem_bulk + 1024) <= (size_t)device_props().max_smem_optin + 1024;
#ifdef QRAG_TUNING
        if (fp.tune & 2) bulk = false;
#endif
        int per_sm = 3;                                        // resident CTAs per SM: registers / shared memory
        if (bulk) {
            per_sm = (int)((228 * 1024) / (smem_bulk + 1024));
            if (per_sm > 4) per_sm = 4;
        }
        // chunks of one query side by side only as far as it takes to fill the machine once
        int ry = (int)ceil_div((int64_t)per_sm * device_props().sm_count, q1 - q0);
        const int rchunks = (int)ceil_div(pl.cand_cap, XS_RESCORE_ROWS);
        if (ry > rchunks) ry = rchunks;
#ifdef QRAG_TUNING
        if (const char* e = getenv("QRAG_TC_RESCORE_Y")) ry = atoi(e) > 0 ? (atoi(e) < rchunks ? atoi(e) : rchunks) : ry;
#endif
        const dim3 rgrid((unsigned)(q1 - q0), (unsigned)ry);
        // one CTA per query and a list that fits the row stages: the rescoring kernel sorts too
        fp.fuse_sort = (bulk && ry == 1 && (size_t)pl.cand_cap * 16 <= (size_t)XS_WARPS * XS_ROWS * D * 4) ? 1 : 0;
#ifdef QRAG_TUNING
        if (fp.tune & 4) fp.fuse_sort = 0;
#endif
        if (bulk) {
            if (smem_bulk > 48 * 1024)
                QRAG_CUDA_CHECK(cudaFuncSetAttribute(tc_rescore_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bulk));
            QRAG_CUDA_CHECK(launch_chained(tc_rescore_bulk_kernel, rgrid, dim3(XS_THREADS), smem_bulk, st, fp));
        } else {
            if (pl.smem_final > 48 * 1024) {
                QRAG_CUDA_CHECK(cudaFuncSetAttribute(tc_rescore_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_final));
                QRAG_CUDA_CHECK(cudaFuncSetAttribute(tc_rescore_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_final));
            }
            if (vec) QRAG_CUDA_CHECK(launch_chained(tc_rescore_kernel<true>, rgrid, dim3(XS_THREADS), pl.smem_final, st, fp));
            else     QRAG_CUDA_CHECK(launch_chained(tc_rescore_kernel<false>, rgrid, dim3(XS_THREADS), pl.smem_final, st, fp));
        }
        QRAG_LAUNCH_CHECK("tc_rescore_kernel");
        // about one compare-exchange per thread and stage for the list length to EXPECT (this shard's share of k plus
        // the margin; the capacity is several times that), within [256, 1024] threads: small CTAs for short lists,
        // so that 1024 queries are one wave
        const int64_t expect = next_pow2(((int64_t)k + G - 1) / G * 5 / 4 + 16);
        int sort_threads = (int)(expect / 2 < 256 ? 256 : (expect / 2 > 1024 ? 1024 : expect / 2));
#ifdef QRAG_TUNING
        if (const char* e = getenv("QRAG_TC_SORT_THREADS")) sort_threads = atoi(e);
#endif
        if (fp.fuse_sort) continue;
        if (pack != nullptr) {
            const size_t smem_sort = (size_t)pl.cand_cap * 16;
            if (smem_sort > 48 * 1024)
                QRAG_CUDA_CHECK(cudaFuncSetAttribute(tc_sort_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sort));
            QRAG_CUDA_CHECK(launch_chained(tc_sort_pack_kernel, dim3(q1 - q0), dim3(sort_threads), smem_sort, st, fp));
            QRAG_LAUNCH_CHECK("tc_sort_pack_kernel");
        } else {
            const size_t smem_sort = (size_t)pl.cand_cap * 12;
            if (smem_sort > 48 * 1024)
                QRAG_CUDA_CHECK(cudaFuncSetAttribute(tc_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sort));
            QRAG_CUDA_CHECK(launch_chained(tc_sort_kernel, dim3(q1 - q0), dim3(sort_threads), smem_sort, st, fp));
            QRAG_LAUNCH_CHECK("tc_sort_kernel");
        }
    }
    return QRAG_OK;
}

extern "C" int qrag_search_tc_finish(const float* Q, int nq, const float* X, int64_t N, int D, int k, int metric,
                                     int64_t id_base, const int32_t* hist_all, int G, double* out_scores, int64_t* out_ids,
                                     int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
    QRAG_REQUIRE(Q && X && out_scores && out_ids && status && G >= 1 && (hist_all || G == 1), QRAG_ERR_INVALID,
                 "bad argument");
    return tc_finish_impl(Q, nq, X, N, D, k, metric, id_base, hist_all, G, out_scores, out_ids, status, nullptr, 0,
                          workspace, workspace_bytes, stream);
}

extern "C" int qrag_search_tc_exchange_len(int k, int G, int* len) {
    QRAG_REQUIRE(len != nullptr && k >= 1 && G >= 1, QRAG_ERR_INVALID, "bad argument");
    *len = tc_exchange_len(k, G);
    return QRAG_OK;
}

extern "C" int qrag_search_tc_finish_packed(const float* Q, int nq, const float* X, int64_t N, int D, int k, int metric,
                                            int64_t id_base, const int32_t* hist_all, int G, int kk, int64_t* pack,
                                            void* workspace, size_t workspace_bytes, void* stream) {
    QRAG_REQUIRE(Q && X && pack && G >= 1 && (hist_all || G == 1), QRAG_ERR_INVALID, "bad argument");
    QRAG_REQUIRE(kk >= 1 && kk <= k, QRAG_ERR_INVALID, "kk=%d outside [1, k=%d]", kk, k);
    return tc_finish_impl(Q, nq, X, N, D, k, metric, id_base, hist_all, G, nullptr, nullptr, nullptr,
                          reinterpret_cast<long long*>(pack), kk, workspace, workspace_bytes, stream);
}

extern "C" int qrag_owner_finalize(const int64_t* recv, int G, int per, int kk, int k1, int k2, int metric, int q_base,
                                   int nq, int64_t* out, void* stream) {
    QRAG_REQUIRE(recv && out, QRAG_ERR_INVALID, "null pointer argument");
    QRAG_REQUIRE(G >= 1 && G <= OWNER_MAX_G, QRAG_ERR_UNSUPPORTED, "owner merge supports 1 <= G <= %d (got %d)", OWNER_MAX_G, G);
    QRAG_REQUIRE(per >= 0 && kk >= 1 && k1 >= 1 && k2 >= 1 && k2 <= k1 && q_base >= 0 && nq >= 0, QRAG_ERR_INVALID,
                 "bad sizes per=%d kk=%d k1=%d k2=%d", per, kk, k1, k2);
    QRAG_REQUIRE(metric >= 0 && metric <= 2, QRAG_ERR_INVALID, "unknown metric %d", metric);
    if (per == 0) return QRAG_OK;
    const DeviceProps& dp = device_props();
    QRAG_REQUIRE(dp.ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    const int64_t slots = (int64_t)G * kk;
    const int members_max = (int)(slots < k1 ? slots : k1);
    const size_t smem = (size_t)slots * 16 + (size_t)next_pow2(members_max) * 28;   // (key, id) per slot; per member: key,
                                                                                       // rank, slot + the winners' (key, rank)
    QRAG_REQUIRE(smem + 2048 <= (size_t)dp.max_smem_optin, QRAG_ERR_UNSUPPORTED,
                 "owner merge of %d lists x %d entries needs %zu B of shared memory", G, kk, smem);
    if (smem > 40 * 1024)                                  // the kernel also holds ~1 KB of static shared memory
        QRAG_CUDA_CHECK(cudaFuncSetAttribute(owner_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OwnerParams op{reinterpret_cast<const long long*>(recv), G, per, kk, k1, k2, metric == QRAG_METRIC_L2 ? 1 : 0, q_base,
                   nq, reinterpret_cast<long long*>(out)};
    QRAG_CUDA_CHECK(launch_chained(owner_finalize_kernel, dim3(per), dim3(1024), smem, (cudaStream_t)stream, op));
    QRAG_LAUNCH_CHECK("owner_finalize_kernel");
    return QRAG_OK;
}

// single shard: the three phases back to back, exchanging through the workspace
extern "C" int qrag_search_topk_tc(const float* Q, int nq, const float* X, const uint16_t* Xb, const float* aux, int64_t N,
                                   int D, int k, int metric, int64_t id_base, double* out_scores, int64_t* out_ids,
                                   int32_t* status, void* workspace, size_t workspace_bytes, void* stream) {
    QRAG_REQUIRE(Q && X && Xb && aux && out_scores && out_ids && status, QRAG_ERR_INVALID, "null pointer argument");
    if (nq == 0) return QRAG_OK;
    TcWs w;
    int rc = tc_ws(nq, N, D, k, metric, 1, workspace, workspace_bytes, &w);
    if (rc) return rc;
    rc = qrag_search_tc_begin(Q, nq, Xb, N, D, k, metric, 1, nullptr, workspace, workspace_bytes, stream);
    if (rc) return rc;
    rc = qrag_search_tc_filter(nq, Xb, aux, N, D, k, metric, nullptr, 1, nullptr, workspace, workspace_bytes, stream);
    if (rc) return rc;
    return qrag_search_tc_finish(Q, nq, X, N, D, k, metric, id_base, nullptr, 1, out_scores, out_ids, status, workspace,
                                 workspace_bytes, stream);
}
