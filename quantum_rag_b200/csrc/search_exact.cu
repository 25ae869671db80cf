// K3a / K5: exact brute-force search over a flat index and top-k list merging.
//
// The flat index is what the ingest tool builds with faiss.IndexFlatL2
// (/root/reference/mcp/server/tools/store_in_faiss.py:99-109); the reference never
// searches it, so the semantics are faiss's API contract made canonical:
// fp32 inputs, fp64 accumulation, order (best score, smaller id), ids int64,
// padding id -1 when k > N.
//
// Phase 1 (search_chunk_kernel): one CTA scores one chunk of rows against one query
//   (query staged in shared memory as fp64, 4 rows in flight per warp, warp-shuffle
//   reductions), sorts the chunk in shared memory and emits its best kk entries.
// Phase 2 (merge_lists_kernel): groups of lists are sorted together in shared memory
//   until one list per query remains.  The same kernel merges the per-shard lists
//   after the NCCL all-gather (qrag_topk_merge).
#include "common.cuh"
#include "exact_score.cuh"
#include "sort.cuh"

namespace qrag {

constexpr int SE_THREADS = XS_THREADS;
constexpr int SE_WARPS = XS_WARPS;
constexpr int SE_ROWS = XS_ROWS;
constexpr int MERGE_MAX = 8192;          // pairs sorted at once by merge_lists_kernel (128 KB smem)
constexpr long long TAG_PAD = 0x7fffffffffffffffLL;

struct ChunkParams {
    const float* Q; const float* X;
    int nq; int64_t N; int D; int metric; int64_t id_base;
    int chunk, kk, nchunks;
    double* ws_key; long long* ws_tag;    // [nq, nchunks, kk]
};

template <bool VEC>
__global__ void __launch_bounds__(SE_THREADS) search_chunk_kernel(const ChunkParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = p.D, Dpad = (D + 3) & ~3;
    double* qs = reinterpret_cast<double*>(smem_raw);
    double* red = qs + Dpad;
    double* key = red + SE_WARPS;
    long long* tag = reinterpret_cast<long long*>(key + p.chunk);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x;
    const int64_t row0 = (int64_t)c * p.chunk;
    const int rows = (int)((p.N - row0) < p.chunk ? (p.N - row0) : p.chunk);
    int P = 1;
    while (P < rows) P <<= 1;
    const bool l2 = p.metric == QRAG_METRIC_L2;

    for (int q = blockIdx.y; q < p.nq; q += gridDim.y) {
        const double nq2 = xs_stage_query(p.Q + (size_t)q * D, D, qs, red);

        for (int r0 = warp * SE_ROWS; r0 < P; r0 += SE_WARPS * SE_ROWS) {
            if (r0 >= rows) {                                   // padding slots of the sort
                if (lane < SE_ROWS && r0 + lane < P) { key[r0 + lane] = pos_inf(); tag[r0 + lane] = TAG_PAD; }
                continue;
            }
            const float* rp[SE_ROWS];
#pragma unroll
            for (int i = 0; i < SE_ROWS; ++i) {
                const int r = (r0 + i < rows) ? r0 + i : r0;
                rp[i] = p.X + (size_t)(row0 + r) * D;
            }
            double nd2;
            const double tot = xs_score4<VEC>(rp, qs, D, l2, lane, nd2);
            if ((lane & 7) == 0) {
                const int i = lane >> 3, r = r0 + i;
                if (r < rows) {
                    const double kv = xs_key(p.metric, tot, nd2, nq2);
                    key[r] = kv;
                    tag[r] = p.id_base + row0 + r;
                } else if (r < P) {
                    key[r] = pos_inf();
                    tag[r] = TAG_PAD;
                }
            }
        }
        __syncthreads();
        block_bitonic_sort<long long>(key, tag, P);
        double* ok = p.ws_key + ((size_t)q * p.nchunks + c) * p.kk;
        long long* ot = p.ws_tag + ((size_t)q * p.nchunks + c) * p.kk;
        for (int i = tid; i < p.kk; i += SE_THREADS) {
            const bool have = i < P;
            ok[i] = have ? key[i] : pos_inf();
            ot[i] = have ? tag[i] : TAG_PAD;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// merge: for each query, lists [l0, l0+g) of length kin -> one list of length kout
// ---------------------------------------------------------------------------
struct MergeParams {
    const double* in_key; const long long* in_tag;      // element (l, q, i) at l*stride_l + q*stride_q + i
    int64_t stride_l, stride_q;
    int nlists, kin, group, nq;
    int raw_scores;         // inputs are scores/ids (id < 0 = padding) rather than keys/tags
    int metric;
    double* out_key; long long* out_tag;                // [nq, ngroups, kout]
    int kout;
    int final_out;          // write scores/ids (convert keys back, pads -> id -1)
    int pmax;               // capacity (pairs) of the shared-memory sort arrays
};

__global__ void __launch_bounds__(256) merge_lists_kernel(const MergeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ngroups = (p.nlists + p.group - 1) / p.group;
    const int gi = blockIdx.x;
    const int l0 = gi * p.group;
    const int nl = (p.nlists - l0) < p.group ? (p.nlists - l0) : p.group;
    const int total = nl * p.kin;
    double* key = reinterpret_cast<double*>(smem_raw);
    long long* tag = reinterpret_cast<long long*>(key + p.pmax);
    const bool l2 = p.metric == QRAG_METRIC_L2;
    __shared__ int s_cnt;
    for (int q = blockIdx.y; q < p.nq; q += gridDim.y) {
        // compact: padding entries (short shards, per-shard lists of a sharded search) are dropped before
        // the sort, so the sort size follows the number of real entries, not G * k
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int l = l0 + i / p.kin, e = i % p.kin;
            const size_t off = (size_t)l * p.stride_l + (size_t)q * p.stride_q + e;
            double kv = p.in_key[off];
            const long long tv = p.in_tag[off];
            const bool pad = p.raw_scores ? tv < 0 : tv == TAG_PAD;
            if (!pad) {
                if (p.raw_scores && !l2) kv = -kv;
                const int pos = atomicAdd(&s_cnt, 1);
                key[pos] = kv;
                tag[pos] = tv;
            }
        }
        __syncthreads();
        const int real = s_cnt;
        int P = 1;
        while (P < real) P <<= 1;
        for (int i = real + threadIdx.x; i < P; i += blockDim.x) { key[i] = pos_inf(); tag[i] = TAG_PAD; }
        __syncthreads();
        block_bitonic_sort<long long>(key, tag, P);
        double* ok = p.out_key + ((size_t)q * ngroups + gi) * p.kout;
        long long* ot = p.out_tag + ((size_t)q * ngroups + gi) * p.kout;
        for (int i = threadIdx.x; i < p.kout; i += blockDim.x) {
            double kv = i < P ? key[i] : pos_inf();
            long long tv = i < P ? tag[i] : TAG_PAD;
            if (p.final_out) {
                if (tv == TAG_PAD) { tv = -1; kv = l2 ? pos_inf() : -pos_inf(); }
                else if (!l2) kv = -kv;
            }
            ok[i] = kv;
            ot[i] = tv;
        }
        __syncthreads();
    }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int search_plan(int nq, int64_t N, int k, int* chunk, int* kk, int* nchunks, int* group) {
    QRAG_REQUIRE(k >= 1 && k <= 2048, QRAG_ERR_UNSUPPORTED, "exact search supports 1 <= k <= 2048 (got %d)", k);
    *chunk = k > 1024 ? 4096 : 2048;
    *kk = k;
    const int64_t nc = N > 0 ? ceil_div(N, *chunk) : 1;
    QRAG_REQUIRE(nc <= 0x7fffffff, QRAG_ERR_UNSUPPORTED, "N too large");
    *nchunks = (int)nc;
    int g = MERGE_MAX / *kk;
    if (g < 2) g = 2;
    *group = g;
    (void)nq;
    return QRAG_OK;
}

static int launch_merge(const MergeParams& p, cudaStream_t st) {
    const int ngroups = (p.nlists + p.group - 1) / p.group;
    const size_t P = (size_t)next_pow2((int64_t)p.group * p.kin);
    const size_t smem = P * (sizeof(double) + sizeof(long long));
    QRAG_REQUIRE(smem <= (size_t)device_props().max_smem_optin, QRAG_ERR_UNSUPPORTED,
                 "merge of %d lists x %d entries needs %zu B shared memory", p.group, p.kin, smem);
    if (smem > 48 * 1024)
        QRAG_CUDA_CHECK(cudaFuncSetAttribute(merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ngroups, p.nq < 65535 ? p.nq : 65535);
    MergeParams pp = p;
    pp.pmax = (int)P;
    merge_lists_kernel<<<grid, 256, smem, st>>>(pp);
    QRAG_LAUNCH_CHECK("merge_lists_kernel");
    return QRAG_OK;
}

}  // namespace qrag

using namespace qrag;

extern "C" int qrag_search_workspace(int nq, int64_t N, int D, int k, size_t* bytes) {
    QRAG_REQUIRE(bytes != nullptr, QRAG_ERR_INVALID, "bytes is null");
    QRAG_REQUIRE(nq >= 0 && N >= 0 && D > 0, QRAG_ERR_INVALID, "bad sizes");
    int chunk = 0, kk = 0, nchunks = 0, group = 0;
    int rc = search_plan(nq, N, k, &chunk, &kk, &nchunks, &group);
    if (rc) return rc;
    const size_t lvl0 = (size_t)nq * nchunks * kk;
    const size_t lvl1 = (size_t)nq * ceil_div(nchunks, group) * kk;
    *bytes = align_up(lvl0 * 16, 256) + align_up(lvl1 * 16, 256) + 256;
    return QRAG_OK;
}

extern "C" int qrag_search_topk(const float* Q, int nq, const float* X, int64_t N, int D, int k, int metric,
                                int64_t id_base, double* out_scores, int64_t* out_ids, void* workspace,
                                size_t workspace_bytes, void* stream) {
    QRAG_REQUIRE(Q && out_scores && out_ids, QRAG_ERR_INVALID, "null pointer argument");
    QRAG_REQUIRE(X != nullptr || N == 0, QRAG_ERR_INVALID, "X is null");
    QRAG_REQUIRE(nq >= 0 && N >= 0 && D > 0, QRAG_ERR_INVALID, "bad sizes nq=%d N=%lld D=%d", nq, (long long)N, D);
    QRAG_REQUIRE(metric >= 0 && metric <= 2, QRAG_ERR_INVALID, "unknown metric %d", metric);
    int chunk = 0, kk = 0, nchunks = 0, group = 0;
    int rc = search_plan(nq, N, k, &chunk, &kk, &nchunks, &group);
    if (rc) return rc;
    if (nq == 0) return QRAG_OK;
    size_t need;
    rc = qrag_search_workspace(nq, N, D, k, &need);
    if (rc) return rc;
    QRAG_REQUIRE(workspace != nullptr && workspace_bytes >= need, QRAG_ERR_WORKSPACE,
                 "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    const DeviceProps& dp = device_props();
    QRAG_REQUIRE(dp.ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    cudaStream_t st = (cudaStream_t)stream;

    // carve workspace: two ping-pong levels of (key, tag)
    const size_t lvl0 = (size_t)nq * nchunks * kk;
    const size_t lvl1 = (size_t)nq * ceil_div(nchunks, group) * kk;
    unsigned char* base = reinterpret_cast<unsigned char*>(align_up((size_t)workspace, 256));
    double* key0 = reinterpret_cast<double*>(base);
    long long* tag0 = reinterpret_cast<long long*>(key0 + lvl0);
    unsigned char* base1 = base + align_up(lvl0 * 16, 256);
    double* key1 = reinterpret_cast<double*>(base1);
    long long* tag1 = reinterpret_cast<long long*>(key1 + lvl1);

    ChunkParams cp{Q, X, nq, N, D, metric, id_base, chunk, kk, nchunks, key0, tag0};
    const int Dpad = (D + 3) & ~3;
    const size_t smem = (size_t)(Dpad + SE_WARPS) * sizeof(double) + (size_t)chunk * 16;
    QRAG_REQUIRE(smem <= (size_t)dp.max_smem_optin, QRAG_ERR_UNSUPPORTED, "D=%d too large for the exact search", D);
    const bool vec = (D % 4 == 0) && ((uintptr_t)X % 16 == 0);
    dim3 grid(nchunks, nq < 65535 ? nq : 65535);
    if (vec) {
        if (smem > 48 * 1024)
            QRAG_CUDA_CHECK(cudaFuncSetAttribute(search_chunk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        search_chunk_kernel<true><<<grid, SE_THREADS, smem, st>>>(cp);
    } else {
        if (smem > 48 * 1024)
            QRAG_CUDA_CHECK(cudaFuncSetAttribute(search_chunk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        search_chunk_kernel<false><<<grid, SE_THREADS, smem, st>>>(cp);
    }
    QRAG_LAUNCH_CHECK("search_chunk_kernel");

    // merge levels
    const double* in_key = key0; const long long* in_tag = tag0;
    double* out_key = key1; long long* out_tag = tag1;
    int nlists = nchunks;
    while (true) {
        const int ngroups = (nlists + group - 1) / group;
        MergeParams mp{};
        mp.in_key = in_key; mp.in_tag = in_tag;
        mp.stride_q = (int64_t)nlists * kk; mp.stride_l = kk;
        mp.nlists = nlists; mp.kin = kk; mp.group = group; mp.nq = nq; mp.metric = metric;
        mp.kout = kk;
        if (ngroups == 1) {
            mp.final_out = 1;
            mp.out_key = out_scores; mp.out_tag = reinterpret_cast<long long*>(out_ids);
            return launch_merge(mp, st);
        }
        mp.out_key = out_key; mp.out_tag = out_tag;
        rc = launch_merge(mp, st);
        if (rc) return rc;
        // swap
        const double* nk = out_key; const long long* nt = out_tag;
        out_key = const_cast<double*>(in_key); out_tag = const_cast<long long*>(in_tag);
        in_key = nk; in_tag = nt;
        nlists = ngroups;
    }
}

// levels of a merge of G lists of k entries when G * k exceeds what one CTA sorts at once: groups of `group` lists
// are merged to min(k_out, ...) entries each until one list remains
static int merge_group(int k) {
    int g = MERGE_MAX / (k > 0 ? k : 1);
    return g < 2 ? 2 : g;
}

extern "C" int qrag_topk_merge_workspace(int G, int nq, int k, int k_out, size_t* bytes) {
    QRAG_REQUIRE(bytes != nullptr, QRAG_ERR_INVALID, "bytes is null");
    QRAG_REQUIRE(G >= 1 && nq >= 0 && k >= 1 && k_out >= 1, QRAG_ERR_INVALID, "bad sizes G=%d nq=%d k=%d k_out=%d", G, nq, k,
                 k_out);
    *bytes = 0;
    if ((int64_t)G * k <= MERGE_MAX) return QRAG_OK;
    const int kl = k_out < k ? k_out : k;                    // entries a group keeps: no list can contribute more
    QRAG_REQUIRE(2 * (int64_t)(k > kl ? k : kl) <= MERGE_MAX, QRAG_ERR_UNSUPPORTED, "lists of %d entries are too long to merge", k);
    const int g1 = merge_group(k);
    const size_t l1 = (size_t)nq * ceil_div(G, g1) * kl;
    const size_t l2 = (size_t)nq * ceil_div(ceil_div(G, g1), merge_group(kl)) * kl;
    *bytes = align_up(l1 * 16, 256) + align_up(l2 * 16, 256) + 256;
    return QRAG_OK;
}

extern "C" int qrag_topk_merge(const double* scores, const int64_t* ids, int G, int nq, int k, int k_out, int metric,
                               double* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream) {
    QRAG_REQUIRE(scores && ids && out_scores && out_ids, QRAG_ERR_INVALID, "null pointer argument");
    QRAG_REQUIRE(G >= 1 && nq >= 0 && k >= 1 && k_out >= 1, QRAG_ERR_INVALID, "bad sizes G=%d nq=%d k=%d k_out=%d", G,
                 nq, k, k_out);
    QRAG_REQUIRE(metric >= 0 && metric <= 2, QRAG_ERR_INVALID, "unknown metric %d", metric);
    if (nq == 0) return QRAG_OK;
    QRAG_REQUIRE(device_props().ok, QRAG_ERR_CUDA, "no CUDA device available (libqrag has no CPU fallback)");
    cudaStream_t st = (cudaStream_t)stream;
    MergeParams mp{};
    mp.in_key = scores; mp.in_tag = reinterpret_cast<const long long*>(ids);
    mp.stride_l = (int64_t)nq * k; mp.stride_q = k;
    mp.nlists = G; mp.kin = k; mp.nq = nq; mp.raw_scores = 1; mp.metric = metric;
    if ((int64_t)G * k <= MERGE_MAX) {                       // one level
        mp.group = G;
        mp.out_key = out_scores; mp.out_tag = reinterpret_cast<long long*>(out_ids);
        mp.kout = k_out; mp.final_out = 1;
        return launch_merge(mp, st);
    }
    // several levels, ping-ponging between two workspace buffers of (key, tag) lists
    size_t need = 0;
    int rc = qrag_topk_merge_workspace(G, nq, k, k_out, &need);
    if (rc) return rc;
    QRAG_REQUIRE(workspace != nullptr && workspace_bytes >= need, QRAG_ERR_WORKSPACE,
                 "merge of %d lists x %d entries needs a workspace of %zu bytes (qrag_topk_merge_workspace), got %zu", G, k,
                 need, workspace_bytes);
    const int kl = k_out < k ? k_out : k;
    const int g1 = merge_group(k);
    const size_t l1 = (size_t)nq * ceil_div(G, g1) * kl;
    unsigned char* base = reinterpret_cast<unsigned char*>(align_up((size_t)workspace, 256));
    double* key_a = reinterpret_cast<double*>(base);
    long long* tag_a = reinterpret_cast<long long*>(key_a + l1);
    unsigned char* base_b = base + align_up(l1 * 16, 256);
    const size_t l2 = (size_t)nq * ceil_div(ceil_div(G, g1), merge_group(kl)) * kl;
    double* key_b = reinterpret_cast<double*>(base_b);
    long long* tag_b = reinterpret_cast<long long*>(key_b + l2);
    // level 1: raw (score, id) lists -> keys/tags
    mp.group = g1; mp.kout = kl; mp.final_out = 0;
    mp.out_key = key_a; mp.out_tag = tag_a;
    rc = launch_merge(mp, st);
    if (rc) return rc;
    int nlists = (int)ceil_div(G, g1);
    const double* in_key = key_a; const long long* in_tag = tag_a;
    double* ok = key_b; long long* ot = tag_b;
    const int g2 = merge_group(kl);
    while (true) {
        const int ngroups = (nlists + g2 - 1) / g2;
        MergeParams m2{};
        m2.in_key = in_key; m2.in_tag = in_tag;
        m2.stride_q = (int64_t)nlists * kl; m2.stride_l = kl;
        m2.nlists = nlists; m2.kin = kl; m2.group = g2; m2.nq = nq; m2.metric = metric;
        if (ngroups == 1) {
            m2.kout = k_out; m2.final_out = 1;
            m2.out_key = out_scores; m2.out_tag = reinterpret_cast<long long*>(out_ids);
            return launch_merge(m2, st);
        }
        m2.kout = kl;
        m2.out_key = ok; m2.out_tag = ot;
        rc = launch_merge(m2, st);
        if (rc) return rc;
        const double* nk = ok; const long long* nt = ot;
        ok = const_cast<double*>(in_key); ot = const_cast<long long*>(in_tag);
        in_key = nk; in_tag = nt;
        nlists = ngroups;
    }
}
