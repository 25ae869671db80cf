/*
 * qrag.h -- C ABI of libqrag.so, the B200 (sm_100a) reranking hot path behind
 * the quantum-rag reranker API.
 *
 * The reference (jon-fox/quantum-rag) has no native/FFI interface on this path:
 * its boundary is the Python class API of src/reranker (app.py:12-13 imports
 * it).  These entry points are what a ctypes binding inside those classes
 * calls; each one names the reference code it replaces.  INTEGRATION.md shows
 * the binding.
 *
 * Conventions
 *  - every function returns 0 (QRAG_OK) or a negative QRAG_ERR_* code;
 *    qrag_last_error() returns a thread-local, human-readable message;
 *  - all data pointers are DEVICE pointers owned by the caller (row-major,
 *    densely packed unless a leading dimension is given); nothing is allocated
 *    inside except where a workspace is passed in explicitly;
 *  - `stream` is a cudaStream_t / CUstream passed as void*; kernels are
 *    enqueued on it and the call returns without synchronising;
 *  - there is NO CPU fallback: without a CUDA device the calls fail with
 *    QRAG_ERR_CUDA.
 *
 * Ordering contract (bit-exact with the reference's `sorted(..., reverse=True)`,
 * quantum.py:70-72 / classical.py:302-304): best score first, ties broken by
 * the smaller input position / id.
 */
#ifndef QRAG_H_
#define QRAG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QRAG_OK               0
#define QRAG_ERR_INVALID     -1   /* bad argument (null pointer, size, alignment) */
#define QRAG_ERR_CUDA        -2   /* CUDA runtime / driver error, or no device    */
#define QRAG_ERR_UNSUPPORTED -3   /* shape outside what the kernels implement     */
#define QRAG_ERR_WORKSPACE   -4   /* workspace too small                           */
#define QRAG_ERR_INEXACT     -5   /* tensor-core search could not certify exactness (never silent) */

#define QRAG_METRIC_IP     0      /* inner product, descending                     */
#define QRAG_METRIC_L2     1      /* squared L2, ascending (faiss METRIC_L2 == 1)  */
#define QRAG_METRIC_COSINE 2      /* cosine similarity, descending                 */

#define QRAG_MAX_QUBITS      12   /* 2^12 complex128 amplitudes staged in shared memory */
#define QRAG_MAX_SORT_LEN  4096   /* longest per-query list the in-kernel sorts accept   */
#define QRAG_TC_HIST_BINS   256   /* per-query survivor-score histogram of the tensor-core search */

const char* qrag_last_error(void);
int         qrag_version(void);
/* sm count / compute capability of the current device */
int         qrag_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Measured FP64 FMA rate of the current device (thread-level fused multiply-adds per second; x2 = flop/s): the
 * roofline the statevector kernels are reported against (MEASURED_PEAKS.json carries HBM and bf16 only).
 * Runs a short register-only kernel on `stream` and SYNCHRONISES.  scratch: 8 bytes of device memory. */
int qrag_probe_fp64_fma_rate(double* fma_per_s, double* scratch, void* stream);

/* ---------------------------------------------------------------------------
 * Stream-overlap policy of the streaming kernels (process-wide, default SAFE).
 * The kernels are launched with programmatic stream serialization so that back-to-back
 * batches on one stream overlap instead of paying launch latency and tail drain per batch:
 *   NONE           plain stream order.
 *   SAFE           the next kernel's CTAs become resident early, but every global access
 *                  waits for the previous kernel to complete (always correct).
 *   INPUTS_STABLE  only the kernel's WRITES wait for the previous kernel; its input reads
 *                  start as soon as an SM is free.  The caller promises that Q / cand / X /
 *                  idx are not produced by the kernel immediately preceding it on the stream
 *                  (copies and events are unaffected: only kernel -> kernel edges relax).
 *   INTERLEAVED    INPUTS_STABLE with half-size CTAs (half an SM's shared memory each, still one CTA per SM
 *                  and launch): the free half of every SM is taken by the NEXT launch on the stream, so
 *                  back-to-back batches run two deep, staggered by half a batch, and each one's start-up and
 *                  drain are covered by the other's streaming.  For steady streams of equal batches; a single
 *                  isolated launch gains nothing.  Same promise about the inputs as INPUTS_STABLE.
 * ------------------------------------------------------------------------- */
#define QRAG_OVERLAP_NONE          0
#define QRAG_OVERLAP_SAFE          1
#define QRAG_OVERLAP_INPUTS_STABLE 2
#define QRAG_OVERLAP_INTERLEAVED   3
int qrag_set_overlap(int mode);
int qrag_get_overlap(void);

/* ---------------------------------------------------------------------------
 * (1a) Reference circuit, exact statevector, complex128.
 * Replaces QuantumReranker._quantum_similarity + _vector_to_circuit
 * (quantum.py:108-167) for a whole batch in one launch.
 *
 *   qvec [nq, vec_len], dvec [nd, vec_len]  fp64 embeddings (not necessarily
 *   normalised; the kernel renormalises like quantum.py:149-151).
 *   Document j is scored against query doc_query[j]; if doc_query is NULL the
 *   layout is dense: query = j / docs_per_query.
 *   layers == 1 is the reference circuit: RY(pi v_i), RZ(pi v_i / 2) on qubit
 *   i < min(vec_len, n_qubits), then CX(i, i+1) for i = 0..n-2.  layers > 1
 *   repeats the block, layer l reading v[(l*n + i) % vec_len].
 *   out_scores[j] = |<psi_d|psi_q>|^2  (qiskit state_fidelity); NaN where doc_query[j] is outside [0, nq)
 *   (n_qubits <= 10; never an out-of-bounds read).
 * ------------------------------------------------------------------------- */
int qrag_sv_fidelity_angle(const double* qvec, int nq,
                           const double* dvec, int64_t nd,
                           const int32_t* doc_query, int64_t docs_per_query,
                           int vec_len, int n_qubits, int layers,
                           double* out_scores, void* stream);

/* ---------------------------------------------------------------------------
 * (1b) Amplitude encoding (+ optional feature-map layers) on fp32 embeddings.
 * The encoding quantum.py:156 names as the real scheme; builder-defined.
 *
 *   Q [nq, D] fp32.  Candidates are either dense  cand [nq, C, D]  (X, idx NULL)
 *   or gathered rows  X[idx[q, c]]  of a corpus X [N, D] (cand NULL; N is ignored for the dense form).
 *   An idx outside [0, N) is padding: it scores -inf and sorts last -- never an out-of-bounds read.
 *   State = zero-pad(x) / |x| on n_qubits (needs D <= 2^n_qubits), followed by
 *   `layers` blocks of the reference circuit with angles x^[(l*n+i) % D].
 *   layers == 0: F = (q.d)^2 / (|q|^2 |d|^2), evaluated with fp64 accumulation.
 *   out64 [nq, C] (required), out32 [nq, C] (optional, rounded from out64).
 * ------------------------------------------------------------------------- */
int qrag_amp_fidelity(const float* Q, int nq,
                      const float* cand, const float* X, int64_t N, const int64_t* idx,
                      int64_t C, int D, int n_qubits, int layers,
                      double* out64, float* out32, void* stream);

/* Kernel choice for layers >= 1 (process-wide, default AUTO).  AUTO runs the warp-per-state
 * register kernel where it applies (n_qubits == 10) and the shared-memory kernel elsewhere;
 * GENERIC forces the shared-memory kernel (A/B timing and cross-checks; results agree to
 * rounding, ~1e-15 relative). */
#define QRAG_FMAP_AUTO    0
#define QRAG_FMAP_GENERIC 1
int qrag_set_fmap_kernel(int mode);

/* ---------------------------------------------------------------------------
 * (1c) Fused amplitude-encoded rerank: score + stable descending sort + top_k
 * in one launch.  Replaces QuantumReranker.rerank (quantum.py:44-78) at tensor
 * level.  C <= QRAG_MAX_SORT_LEN, 0 < top_k <= C, layers == 0.
 *   out_scores [nq, top_k] fp64, out_pos [nq, top_k] position in the candidate
 *   list, out_ids [nq, top_k] = idx[q, pos] (optional; needs idx).
 * ------------------------------------------------------------------------- */
int qrag_amp_rerank(const float* Q, int nq,
                    const float* cand, const float* X, int64_t N, const int64_t* idx,
                    int64_t C, int D, int n_qubits, int top_k,
                    double* out_scores, int32_t* out_pos, int64_t* out_ids,
                    void* stream);

/* (1c, host form) The same rerank called with HOST buffers -- what a serving process behind a retrieval step holds
 * (the string API's documents, quantum.py:44-78, resolved to row ids): queries hQ [nq, D] fp32 and candidate ids
 * hIdx [nq, C] int64 in (pinned) host memory, the corpus X [N, D] resident on the device.  One call enqueues on
 * `stream`: the two host->device copies into `workspace` (device memory, 256-byte aligned, at least
 * qrag_amp_rerank_host_workspace bytes), the fused kernel of qrag_amp_rerank, and the device->host copies of
 * hScores [nq, top_k] fp64 and hIds [nq, top_k] int64.  Nothing is synchronised: the outputs are valid once the
 * stream has reached this point (record an event after the call).  Several calls on different streams with
 * different workspaces overlap copy and kernel. */
int qrag_amp_rerank_host_workspace(int nq, int64_t C, int D, int top_k, size_t* bytes);
int qrag_amp_rerank_host(const float* hQ, int nq, const int64_t* hIdx, int64_t C,
                         const float* X, int64_t N, int D, int n_qubits, int top_k,
                         void* workspace, size_t workspace_bytes,
                         double* hScores, int64_t* hIds, void* stream);

/* ---------------------------------------------------------------------------
 * (1c') Feature-map rerank: amplitude state + `layers` reference blocks at n_qubits == 10 (D <= 1024), top_k of
 * C <= QRAG_MAX_SORT_LEN candidates per query, by FILTER-THEN-CERTIFY: every candidate evolved in complex64
 * (fp64 normalisation, gate parameters and overlap), |F32 - F| <= delta = qrag_fmap_filter_error_bound(layers);
 * the candidates within 2 delta of the k-th best approximate score (a superset of the exact top_k and of everything
 * tied with its last member, typically k + 1..2 of them) re-evolved in complex128 by the kernel qrag_amp_fidelity
 * runs; sorted by (exact score desc, position asc).  Rankings and returned scores are those of the all-complex128
 * path, bit for bit.  status[q] != 0: the query's candidate list overflowed (more than max(64, 2 top_k + 32) within
 * the margin) -- never silent: rerun that query with qrag_amp_fidelity + qrag_sort_scores_stable.
 *   out_scores [nq, top_k] fp64, out_pos [nq, top_k] (-1 padding), out_ids [nq, top_k] = idx[q, pos] (optional).
 * qrag_fmap_filter_scores is the filter pass alone (diagnostic: tests measure its error against delta).
 * ------------------------------------------------------------------------- */
int qrag_fmap_rerank_workspace(int nq, int64_t C, int top_k, size_t* bytes);
int qrag_fmap_filter_error_bound(int layers, double* delta);
int qrag_fmap_filter_scores(const float* Q, int nq, const float* cand, const float* X, int64_t N, const int64_t* idx,
                            int64_t C, int D, int n_qubits, int layers, double* out64, void* stream);
int qrag_fmap_rerank(const float* Q, int nq,
                     const float* cand, const float* X, int64_t N, const int64_t* idx,
                     int64_t C, int D, int n_qubits, int layers, int top_k,
                     double* out_scores, int32_t* out_pos, int64_t* out_ids, int32_t* status,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * (1d) Stable segmented sort: the `sorted(scored, key=score, reverse=True)[:top_k]`
 * of quantum.py:70-76 / classical.py:302-308 for nq lists of C scores.
 *   descending != 0: (score desc, position asc); else (score asc, position asc).
 *   out_perm [nq, top_k] int32 positions; out_sorted [nq, top_k] optional.
 * Lists of up to QRAG_MAX_SORT_LEN scores are sorted in shared memory and need no workspace (NULL, 0);
 * longer lists (any C < 2^31; the reference sorts any length) are sorted in blocks and merged level by
 * level through a workspace of qrag_sort_scores_workspace bytes.
 * ------------------------------------------------------------------------- */
int qrag_sort_scores_workspace(int nq, int64_t C, size_t* bytes);
int qrag_sort_scores_stable(const double* scores, int nq, int64_t C, int top_k, int descending,
                            int32_t* out_perm, double* out_sorted,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * (1e) The reference's text-hash embedding (quantum.py:169-185) for a batch of
 * seeds: legacy MT19937 seeded with sum(ord(c)), 2*n_qubits doubles in [0,1),
 * L2-normalised.  out [n, 2*n_qubits] fp64.
 * ------------------------------------------------------------------------- */
int qrag_mock_embedding(const uint32_t* seeds, int64_t n, int n_qubits, double* out, void* stream);

/* ---------------------------------------------------------------------------
 * (2) Brute-force search over a flat index (the IndexFlatL2 the ingest tool
 * writes at store_in_faiss.py:99-109; `.search` is never called by the
 * reference, semantics follow faiss's API).  Exact: fp32 inputs, fp64
 * accumulation, order (best, id asc), ids = id_base + row, -1 padding if k > N.
 *
 *   qrag_search_topk       CUDA-core exact path (any nq; the checker for the tensor-core path).
 *   qrag_search_topk_tc    tcgen05 path: 16-bit (bf16 / fp16) GEMM used as a filter with a proven error bound
 *                          (bucket-maximum pass -> per-query threshold -> filter pass; the score
 *                          matrix never leaves TMEM), exact fp64 rescoring of the survivors with the
 *                          same code as the CUDA-core path, so scores and ids are bit-identical to
 *                          qrag_search_topk.  status[q] != 0 flags a query whose candidate list
 *                          overflowed (never silent): rerun those with qrag_search_topk.
 * ------------------------------------------------------------------------- */
int qrag_search_workspace(int nq, int64_t N, int D, int k, size_t* bytes);
int qrag_search_topk(const float* Q, int nq, const float* X, int64_t N, int D, int k,
                     int metric, int64_t id_base,
                     double* out_scores, int64_t* out_ids,
                     void* workspace, size_t workspace_bytes, void* stream);

/* 16-bit shadow of the corpus for one metric, built once per index (shard):
 *   Xb [N, Kp] 16-bit words with Kp from qrag_index_prepared_dims (IP: x as bf16; cosine: x/|x| as fp16 --
 *   normalised rows fit fp16's range and keep 3 more bits; L2: [x, hi(|x|^2), lo(|x|^2)] as bf16),
 *   aux [4] floats (aux[0] = max |x|, aux[1] = max over rows of the MEASURED rounding-error norm
 *   |b - round16(b)|: the filter's error bound uses both). */
int qrag_index_prepared_dims(int D, int metric, int* Kp);
int qrag_index_prepare(const float* X, int64_t N, int D, int metric,
                       uint16_t* Xb, float* aux, void* stream);
int qrag_search_tc_workspace(int nq, int64_t N, int D, int k, int metric, int shards, size_t* bytes);
int qrag_search_topk_tc(const float* Q, int nq, const float* X, const uint16_t* Xb, const float* aux,
                        int64_t N, int D, int k, int metric, int64_t id_base,
                        double* out_scores, int64_t* out_ids, int32_t* status,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Diagnostic: out [nq, N] = every approximate score of the filter GEMM (16-bit operands, fp32 accumulation in
 * tensor memory), in the units the filter thresholds use (IP / cosine: the similarity; L2: 2 q.x - |x|^2).  This is
 * the quantity the filter's error bound is a bound on; tests measure the bound's terms with it.  Workspace as for
 * qrag_search_tc_workspace(nq, N, D, 1, metric, 1). */
int qrag_search_tc_scores(const float* Q, int nq, const uint16_t* Xb, int64_t N, int D, int metric, float* out,
                          void* workspace, size_t workspace_bytes, void* stream);

/* The same search split at its two exchange points, for a corpus sharded over G GPUs (the
 * all-gathers in between are the caller's: quantum_rag_b200/sharded.py issues them with NCCL).
 * Thresholds come from ALL shards, so every shard filters and rescores only ~ (k + margin) / G
 * candidates and the work scales with 1/G.  The workspace carries state between the phases.
 *   (`shards` = G sizes the sample and the workspace: the same value in the workspace query and every phase;
 *    qrag_search_topk_tc is the G = 1 composition)
 * Two small exchanges, both the caller's:
 *   (1) all-gather of bm_top [nq, kt], this shard's kt largest sampled bucket maxima, kt =
 *       qrag_search_tc_exchange_len(k, G) (k for G = 1, else min(k, 2k/G + 64 rounded up to 32)): the k-th
 *       largest of the union of the G cut lists is a valid filter threshold, and the exact one unless a
 *       single shard holds more than kt of the global top k;
 *   (2) all-reduce(SUM) of hist [nq, QRAG_TC_HIST_BINS] int32: the filter pass counts every survivor into a
 *       per-query histogram of its approximate score (uniform bins from the threshold up, laid out identically
 *       on every shard because the threshold is global); the highest bin edge with k survivors at or above it
 *       bounds the k-th best approximate score from below, which is all the candidate cut needs.
 *   begin   bm_top [nq, kt]
 *   filter  bm_top_all [G, nq, kt] (all-gathered) -> tau, filter GEMM -> hist [nq, QRAG_TC_HIST_BINS] of THIS
 *           shard.  aux[0] and aux[1] must hold their maxima over ALL shards.
 *   finish  hist_all = hist summed over the shards -> exact, sorted list of this shard's members of the
 *           global top-k (ids -1 padded); merge the G lists with qrag_topk_merge.
 *   (G = 1: bm_top, bm_top_all, hist and hist_all may be NULL; the workspace carries them.) */
int qrag_search_tc_exchange_len(int k, int G, int* len);
int qrag_search_tc_begin(const float* Q, int nq, const uint16_t* Xb, int64_t N, int D, int k, int metric,
                         int shards /* = G of the later phases */,
                         float* bm_top, void* workspace, size_t workspace_bytes, void* stream);
int qrag_search_tc_filter(int nq, const uint16_t* Xb, const float* aux, int64_t N, int D, int k, int metric,
                          const float* bm_top_all, int G, int32_t* hist,
                          void* workspace, size_t workspace_bytes, void* stream);
int qrag_search_tc_finish(const float* Q, int nq, const float* X, int64_t N, int D, int k, int metric,
                          int64_t id_base, const int32_t* hist_all, int G,
                          double* out_scores, int64_t* out_ids, int32_t* status,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Search + quantum rerank (BASELINE config 4) without a second pass over the rows: `finish_packed` is `finish`
 * with the shard's list written as one record per query, cut to its kk best entries, carrying the amplitude-
 * encoded fidelity (q.d)^2 / (|q|^2 |d|^2) of every entry -- computed by the exact rescoring from the same
 * read of the row, bit-identical to qrag_amp_fidelity.  Record of query q at pack + q * (3 kk + 1), int64 words:
 *   [0]            entries in the record | bad << 32   (bad: the shard could not certify the query, or had more
 *                                                       than kk entries; the caller must rerun such a query)
 *   [1, kk]        search scores (fp64 bits), best first       [kk+1, 2kk]  ids (-1 padding)
 *   [2kk+1, 3kk]   fidelities (fp64 bits)
 * The records of query q travel to the rank that owns q (one all-to-all; sharded.py), where
 * `qrag_owner_finalize` merges the G lists by rank, keeps the global top-k1 and orders it by
 * (fidelity desc, position in the merged list asc) -- QuantumReranker.rerank's stable sort, quantum.py:70-76.
 *   recv [G, per, 3 kk + 1]  record of owned query j from every shard;  q_base + j = the query's global number
 *   out  [per, 2 k2 + 1]     k2 fidelities (fp64 bits), k2 ids, status (!= 0: some shard flagged the query) */
int qrag_search_tc_finish_packed(const float* Q, int nq, const float* X, int64_t N, int D, int k, int metric,
                                 int64_t id_base, const int32_t* hist_all, int G, int kk, int64_t* pack,
                                 void* workspace, size_t workspace_bytes, void* stream);
int qrag_owner_finalize(const int64_t* recv, int G, int per, int kk, int k1, int k2, int metric,
                        int q_base, int nq, int64_t* out, void* stream);

/* ---------------------------------------------------------------------------
 * (3) Merge of per-shard top-k lists after the all-gather: scores/ids
 * [G, nq, k] -> [nq, k_out] in the canonical order; id < 0 is padding.
 * Up to G * k = 8192 entries per query are merged by one kernel and need no workspace (NULL, 0);
 * beyond that the lists are merged in groups over several levels through a workspace of
 * qrag_topk_merge_workspace bytes (any G; k <= 4096).
 * ------------------------------------------------------------------------- */
int qrag_topk_merge_workspace(int G, int nq, int k, int k_out, size_t* bytes);
int qrag_topk_merge(const double* scores, const int64_t* ids, int G, int nq, int k, int k_out,
                    int metric, double* out_scores, int64_t* out_ids,
                    void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QRAG_H_ */
