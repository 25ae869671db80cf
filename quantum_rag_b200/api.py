"""Tensor-level API over libqrag (what the reranker classes and the benchmarks call).

torch is used for device memory, streams and (elsewhere) torch.distributed only;
every computation is a hand-written sm_100a kernel reached through the C ABI.
Inputs on the host are copied to the current CUDA device; outputs stay on the
device unless stated otherwise.  No CPU fallback: without CUDA these raise.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _lib

METRIC_IP = 0
METRIC_L2 = 1
METRIC_COSINE = 2
_METRICS = {"ip": METRIC_IP, "inner_product": METRIC_IP, "l2": METRIC_L2, "cosine": METRIC_COSINE,
            "cos": METRIC_COSINE, METRIC_IP: METRIC_IP, METRIC_L2: METRIC_L2, METRIC_COSINE: METRIC_COSINE}

MAX_QUBITS = 12
MAX_SORT_LEN = 4096
HIST_BINS = 256          # QRAG_TC_HIST_BINS

ArrayLike = Union[torch.Tensor, np.ndarray]


def metric_id(metric) -> int:
    try:
        return _METRICS[metric.lower() if isinstance(metric, str) else metric]
    except KeyError:
        raise ValueError(f"unknown metric {metric!r}; expected one of ip, l2, cosine") from None


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("quantum_rag_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _dev(x: Optional[ArrayLike], dtype: torch.dtype) -> Optional[torch.Tensor]:
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    return x.to(device=_device(), dtype=dtype, non_blocking=True).contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _nrows(t: Optional[torch.Tensor]) -> int:
    return int(t.shape[0]) if t is not None else 0


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """The caller's current CUDA stream as a ``cudaStream_t``.  The raw accessor skips building a ``torch.cuda.Stream``
    object (several microseconds per call, which shows in the 20-document string API); same stream either way."""
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


OVERLAP_NONE, OVERLAP_SAFE, OVERLAP_INPUTS_STABLE, OVERLAP_INTERLEAVED = 0, 1, 2, 3


def set_overlap(mode: int) -> int:
    """Stream-overlap policy of the streaming kernels (include/qrag.h); returns the previous mode."""
    lib = _lib.load()
    old = lib.qrag_get_overlap()
    _lib.check(lib.qrag_set_overlap(int(mode)))
    return old


FMAP_AUTO, FMAP_GENERIC = 0, 1


def set_fmap_kernel(mode: int) -> None:
    """Feature-map kernel choice (include/qrag.h): FMAP_AUTO or FMAP_GENERIC (shared-memory kernel everywhere)."""
    _lib.check(_lib.load().qrag_set_fmap_kernel(int(mode)))


def probe_fp64_fma_rate() -> float:
    """Measured FP64 fused multiply-adds per second of the current device (include/qrag.h); synchronises."""
    scratch = torch.zeros(1, dtype=torch.float64, device=_device())
    rate = ctypes.c_double(0.0)
    _lib.check(_lib.load().qrag_probe_fp64_fma_rate(ctypes.byref(rate), _ptr(scratch), _stream()))
    return rate.value


def qubits_for(dim: int) -> int:
    """Smallest n with 2**n >= dim (amplitude encoding zero-pads to 2**n)."""
    return max(1, int(dim - 1).bit_length())


# ---------------------------------------------------------------------------
def sv_fidelity_angle(qvec: ArrayLike, dvec: ArrayLike, doc_query: Optional[ArrayLike] = None,
                      docs_per_query: Optional[int] = None, n_qubits: int = 4, layers: int = 1) -> torch.Tensor:
    """Fidelity of the reference circuit for every document (quantum.py:108-167).

    qvec [nq, L] / dvec [nd, L] fp64.  ``doc_query[j]`` names the query of document j;
    without it the layout is dense (``docs_per_query`` consecutive documents per query,
    default nd // nq).  Returns fp64 [nd] on the device.
    """
    q = _dev(qvec, torch.float64)
    d = _dev(dvec, torch.float64)
    if q.dim() == 1:
        q = q[None, :]
    if d.dim() == 1:
        d = d[None, :]
    if q.shape[1] != d.shape[1]:
        raise ValueError("query and document vectors must have the same length")
    nq, nd = q.shape[0], d.shape[0]
    dq = _dev(doc_query, torch.int32) if doc_query is not None else None
    if dq is not None and dq.numel() != nd:
        raise ValueError("doc_query must have one entry per document")
    if dq is None and docs_per_query is None:
        docs_per_query = max(1, -(-nd // max(nq, 1)))
    out = torch.empty(nd, dtype=torch.float64, device=q.device)
    lib = _lib.load()
    _lib.check(lib.qrag_sv_fidelity_angle(_ptr(q), nq, _ptr(d), nd, _ptr(dq), int(docs_per_query or 0),
                                          q.shape[1], n_qubits, layers, _ptr(out), _stream()))
    return out


def _cand_args(Q, cand, X, idx):
    Qd = _dev(Q, torch.float32)
    if Qd.dim() != 2:
        raise ValueError("Q must be [nq, D]")
    nq, D = Qd.shape
    if cand is not None:
        if X is not None or idx is not None:
            raise ValueError("pass either cand, or X together with idx")
        c = _dev(cand, torch.float32)
        if c.dim() != 3 or c.shape[0] != nq or c.shape[2] != D:
            raise ValueError("cand must be [nq, C, D]")
        return Qd, c, None, None, c.shape[1]
    if X is None or idx is None:
        raise ValueError("pass either cand, or X together with idx")
    Xd = _dev(X, torch.float32)
    i = _dev(idx, torch.int64)
    if Xd.dim() != 2 or Xd.shape[1] != D or i.dim() != 2 or i.shape[0] != nq:
        raise ValueError("X must be [N, D] and idx [nq, C]")
    return Qd, None, Xd, i, i.shape[1]


def amp_fidelity(Q: ArrayLike, cand: Optional[ArrayLike] = None, X: Optional[ArrayLike] = None,
                 idx: Optional[ArrayLike] = None, n_qubits: Optional[int] = None, layers: int = 0,
                 want_fp32: bool = False):
    """Amplitude-encoded fidelity |<q^|d^>|^2 (optionally after ``layers`` feature-map blocks).

    Returns fp64 [nq, C] (and the fp32 rounding if ``want_fp32``).
    """
    Qd, c, Xd, i, C = _cand_args(Q, cand, X, idx)
    nq, D = Qd.shape
    n = qubits_for(D) if n_qubits is None else n_qubits
    out = torch.empty((nq, C), dtype=torch.float64, device=Qd.device)
    out32 = torch.empty((nq, C), dtype=torch.float32, device=Qd.device) if want_fp32 else None
    lib = _lib.load()
    _lib.check(lib.qrag_amp_fidelity(_ptr(Qd), nq, _ptr(c), _ptr(Xd), _nrows(Xd), _ptr(i), C, D, n, layers, _ptr(out),
                                     _ptr(out32), _stream()))
    return (out, out32) if want_fp32 else out


def sort_scores(scores: ArrayLike, top_k: Optional[int] = None, descending: bool = True):
    """Stable segmented sort: the order Python's ``sorted(key=score, reverse=True)`` gives.

    scores [nq, C] fp64 -> (perm int32 [nq, k], sorted fp64 [nq, k]).
    """
    return _sort_scores(scores, top_k, descending)[:2]


def _sort_scores(scores: ArrayLike, top_k: Optional[int], descending: bool):
    """(perm, sorted, the one uint8 buffer both are views of)."""
    s = _dev(scores, torch.float64)
    if s.dim() == 1:
        s = s[None, :]
    nq, C = s.shape
    k = C if top_k is None else max(0, min(int(top_k), C))
    # one allocation for both outputs (sorted scores first: 8-byte aligned), so that a caller who wants them on the host
    # can bring both back with one copy (sort_scores_host)
    both = torch.empty(nq * k * 12, dtype=torch.uint8, device=s.device)
    srt = both[: nq * k * 8].view(torch.float64).view(nq, k)
    perm = both[nq * k * 8:].view(torch.int32).view(nq, k)
    lib = _lib.load()
    # lists longer than the shared-memory sort (the reference sorts any length, quantum.py:70-72) go through the
    # library's block-sort + global-memory merge, which needs a workspace; lists of 65535+ queries go in slices
    nbytes = ctypes.c_size_t(0)
    step = nq if C <= MAX_SORT_LEN else min(nq, 65535)
    _lib.check(lib.qrag_sort_scores_workspace(step, C, ctypes.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=s.device) if nbytes.value else None
    for a in range(0, nq, max(step, 1)):
        b = min(nq, a + step)
        _lib.check(lib.qrag_sort_scores_stable(_ptr(s[a:b]), b - a, C, k, 1 if descending else 0, _ptr(perm[a:b]),
                                               _ptr(srt[a:b]), _ptr(ws), nbytes.value, _stream()))
    return perm, srt, both


def sort_scores_host(scores: ArrayLike, top_k: Optional[int] = None, descending: bool = True):
    """``sort_scores`` with the result on the host as two NumPy arrays (perm int32 [nq, k], sorted fp64 [nq, k]):
    one device-to-host copy for both (what the string API needs for its list of tuples)."""
    perm, _, both = _sort_scores(scores, top_k, descending)
    nq, k = perm.shape
    host = both.cpu().numpy()
    return host[nq * k * 8:].view(np.int32).reshape(nq, k), host[: nq * k * 8].view(np.float64).reshape(nq, k)


def quantum_rerank_batch(Q: ArrayLike, cand: Optional[ArrayLike] = None, X: Optional[ArrayLike] = None,
                         idx: Optional[ArrayLike] = None, top_k: Optional[int] = None,
                         n_qubits: Optional[int] = None, layers: int = 0, certify: bool = True):
    """Tensor-level ``QuantumReranker.rerank`` with amplitude encoding.

    Returns ``(scores fp64 [nq, k], pos int32 [nq, k], ids int64 [nq, k] | None)`` ordered
    by (score desc, candidate position asc); ``ids`` is ``idx[q, pos]`` in the gathered form.
    """
    Qd, c, Xd, i, C = _cand_args(Q, cand, X, idx)
    nq, D = Qd.shape
    n = qubits_for(D) if n_qubits is None else n_qubits
    k = C if top_k is None else max(0, min(int(top_k), C))
    lib = _lib.load()
    if layers == 0 and 1 <= C <= MAX_SORT_LEN and k >= 1:
        scores = torch.empty((nq, k), dtype=torch.float64, device=Qd.device)
        pos = torch.empty((nq, k), dtype=torch.int32, device=Qd.device)
        ids = torch.empty((nq, k), dtype=torch.int64, device=Qd.device) if i is not None else None
        _lib.check(lib.qrag_amp_rerank(_ptr(Qd), nq, _ptr(c), _ptr(Xd), _nrows(Xd), _ptr(i), C, D, n, k, _ptr(scores),
                                       _ptr(pos), _ptr(ids), _stream()))
        return scores, pos, ids
    if layers >= 1 and n == 10 and D <= 1024 and 1 <= C <= MAX_SORT_LEN and k >= 1 and certify:
        # filter-then-certify (include/qrag.h: qrag_fmap_rerank): complex64 evolution of every candidate, complex128
        # re-evolution of the few within the error margin of the top-k boundary; same bits as the all-complex128 path
        scores = torch.empty((nq, k), dtype=torch.float64, device=Qd.device)
        pos = torch.empty((nq, k), dtype=torch.int32, device=Qd.device)
        ids = torch.empty((nq, k), dtype=torch.int64, device=Qd.device) if i is not None else None
        status = torch.empty(nq, dtype=torch.int32, device=Qd.device)
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.qrag_fmap_rerank_workspace(nq, C, k, ctypes.byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=Qd.device)
        _lib.check(lib.qrag_fmap_rerank(_ptr(Qd), nq, _ptr(c), _ptr(Xd), _nrows(Xd), _ptr(i), C, D, n, layers, k, _ptr(scores),
                                        _ptr(pos), _ptr(ids), _ptr(status), _ptr(ws), nbytes.value, _stream()))
        flagged = torch.nonzero(status).flatten()                # synchronises; the certificate is never skipped
        if flagged.numel():                                      # a margin too crowded for the list: those queries exactly
            fs, fp, fi = quantum_rerank_batch(Qd[flagged], cand=c[flagged] if c is not None else None, X=Xd,
                                              idx=i[flagged] if i is not None else None, top_k=k, n_qubits=n, layers=layers,
                                              certify=False)
            scores[flagged], pos[flagged] = fs, fp
            if ids is not None:
                ids[flagged] = fi
        return scores, pos, ids
    full = torch.empty((nq, C), dtype=torch.float64, device=Qd.device)
    _lib.check(lib.qrag_amp_fidelity(_ptr(Qd), nq, _ptr(c), _ptr(Xd), _nrows(Xd), _ptr(i), C, D, n, layers, _ptr(full), None,
                                     _stream()))
    pos, scores = sort_scores(full, k)
    ids = torch.gather(i, 1, pos.long()) if i is not None else None
    return scores, pos, ids


def fmap_filter_scores(Q: ArrayLike, cand: Optional[ArrayLike] = None, X: Optional[ArrayLike] = None,
                       idx: Optional[ArrayLike] = None, layers: int = 4) -> torch.Tensor:
    """Diagnostic (include/qrag.h): the complex64 filter pass of the feature-map rerank alone, fp64 [nq, C]."""
    Qd, c, Xd, i, C = _cand_args(Q, cand, X, idx)
    nq, D = Qd.shape
    out = torch.empty((nq, C), dtype=torch.float64, device=Qd.device)
    _lib.check(_lib.load().qrag_fmap_filter_scores(_ptr(Qd), nq, _ptr(c), _ptr(Xd), _nrows(Xd), _ptr(i), C, D, 10, layers,
                                                   _ptr(out), _stream()))
    return out


def fmap_filter_error_bound(layers: int) -> float:
    d = ctypes.c_double(0.0)
    _lib.check(_lib.load().qrag_fmap_filter_error_bound(int(layers), ctypes.byref(d)))
    return d.value


def mock_embedding(seeds: ArrayLike, n_qubits: int = 4) -> torch.Tensor:
    """quantum.py:169-185 for a batch of seeds (= sum(ord(c))), fp64 [n, 2*n_qubits]."""
    host = np.asarray(seeds.cpu() if isinstance(seeds, torch.Tensor) else seeds, dtype=np.int64).reshape(-1)
    if host.size and (host.min() < 0 or host.max() > 0xFFFFFFFF):
        raise ValueError("Seed must be between 0 and 2**32 - 1")      # numpy's own error text
    s32 = _dev(host.astype(np.uint32).view(np.int32), torch.int32)     # same 32 bits
    out = torch.empty((host.size, 2 * n_qubits), dtype=torch.float64, device=s32.device)
    lib = _lib.load()
    _lib.check(lib.qrag_mock_embedding(_ptr(s32), host.size, n_qubits, _ptr(out), _stream()))
    return out


# ---------------------------------------------------------------------------
def search_topk(Q: ArrayLike, X: ArrayLike, k: int, metric="l2", id_base: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact brute-force top-k (CUDA-core fp64 path).  Returns (scores fp64 [nq,k], ids int64 [nq,k])."""
    Qd = _dev(Q, torch.float32)
    Xd = _dev(X, torch.float32)
    if Qd.dim() == 1:
        Qd = Qd[None, :]
    if Xd.dim() != 2 or Qd.shape[1] != Xd.shape[1]:
        raise ValueError("Q [nq, D] and X [N, D] must share D")
    nq, D = Qd.shape
    N = Xd.shape[0]
    m = metric_id(metric)
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.qrag_search_workspace(nq, N, D, k, ctypes.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=Qd.device)
    scores = torch.empty((nq, k), dtype=torch.float64, device=Qd.device)
    ids = torch.empty((nq, k), dtype=torch.int64, device=Qd.device)
    _lib.check(lib.qrag_search_topk(_ptr(Qd), nq, _ptr(Xd), N, D, k, m, id_base, _ptr(scores), _ptr(ids), _ptr(ws),
                                    nbytes.value, _stream()))
    return scores, ids


class FlatIndexTC:
    """A flat index (shard) prepared for the tcgen05 search: fp32 rows plus their 16-bit shadow (bf16; fp16 for cosine).

    ``search(Q, k)`` returns exactly what ``search_topk`` returns (same fp64 scores, same ids, same
    order): the tensor cores only filter, survivors are rescored by the exact code.  Queries whose
    candidate list overflowed are flagged by the kernel and rerun on the exact path here, so the
    result never silently degrades.
    """

    def __init__(self, X: ArrayLike, metric="l2", id_base: int = 0):
        self.X = _dev(X, torch.float32)
        if self.X.dim() != 2:
            raise ValueError("X must be [N, D]")
        self.N, self.D = self.X.shape
        self.metric = metric_id(metric)
        self.id_base = int(id_base)
        lib = _lib.load()
        kp = ctypes.c_int(0)
        _lib.check(lib.qrag_index_prepared_dims(self.D, self.metric, ctypes.byref(kp)))
        self.Kp = kp.value
        # 16-bit shadow: bf16 (inner product, L2) or fp16 (cosine) -- opaque to the host, typed only for debugging views
        self.shadow_dtype = torch.float16 if self.metric == METRIC_COSINE else torch.bfloat16
        self.Xb = torch.empty((max(self.N, 1), self.Kp), dtype=self.shadow_dtype, device=self.X.device)
        self.aux = torch.empty(4, dtype=torch.float32, device=self.X.device)
        _lib.check(lib.qrag_index_prepare(_ptr(self.X), self.N, self.D, self.metric, _ptr(self.Xb), _ptr(self.aux),
                                          _stream()))
        self._ws = None
        self.last_fallback = 0          # queries the last search() had to rerun exactly

    def fork(self) -> "FlatIndexTC":
        """A second handle on the same rows and shadow with its own search workspace and phase state (nothing is
        copied): two searches can then be between their exchange points at once (sharded.py lanes)."""
        import copy
        f = copy.copy(self)
        f._ws = None
        f._phase = None
        return f

    def _workspace(self, nq: int, k: int, shards: int = 1) -> torch.Tensor:
        lib = _lib.load()
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.qrag_search_tc_workspace(nq, self.N, self.D, k, self.metric, shards, ctypes.byref(nbytes)))
        if self._ws is None or self._ws.numel() < nbytes.value:
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.X.device)
        return self._ws

    def search_async(self, Q: ArrayLike, k: int):
        """Enqueue the search; returns (Q, scores, ids, status) device tensors without synchronising."""
        Qd = _dev(Q, torch.float32)
        if Qd.dim() == 1:
            Qd = Qd[None, :]
        if Qd.shape[1] != self.D:
            raise ValueError("Q [nq, D] must match the index dimension")
        nq = Qd.shape[0]
        lib = _lib.load()
        ws = self._workspace(nq, k)
        scores = torch.empty((nq, k), dtype=torch.float64, device=Qd.device)
        ids = torch.empty((nq, k), dtype=torch.int64, device=Qd.device)
        status = torch.zeros(nq, dtype=torch.int32, device=Qd.device)
        _lib.check(lib.qrag_search_topk_tc(_ptr(Qd), nq, _ptr(self.X), _ptr(self.Xb), _ptr(self.aux), self.N, self.D, k,
                                           self.metric, self.id_base, _ptr(scores), _ptr(ids), _ptr(status), _ptr(ws),
                                           ws.numel(), _stream()))
        return Qd, scores, ids, status

    def approx_scores(self, Q: ArrayLike) -> torch.Tensor:
        """Diagnostic (include/qrag.h: qrag_search_tc_scores): fp32 [nq, N], every approximate score of the filter GEMM."""
        Qd = _dev(Q, torch.float32)
        if Qd.dim() == 1:
            Qd = Qd[None, :]
        nq = Qd.shape[0]
        ws = self._workspace(nq, 1)
        out = torch.empty((nq, self.N), dtype=torch.float32, device=Qd.device)
        _lib.check(_lib.load().qrag_search_tc_scores(_ptr(Qd), nq, _ptr(self.Xb), self.N, self.D, self.metric, _ptr(out),
                                                     _ptr(ws), ws.numel(), _stream()))
        return out

    # ---- the search split at its two exchange points (corpus sharded over G GPUs) ----
    def tc_begin(self, Q: ArrayLike, k: int, shards: int) -> Optional[torch.Tensor]:
        """Phase 1 of a search over ``shards`` shards: bm_top [nq, kt] fp32, this shard's kt largest sampled bucket
        maxima (kt = exchange_len(k, shards)); a single shard keeps them in the workspace and returns None."""
        Qd = _dev(Q, torch.float32)
        if Qd.dim() == 1:
            Qd = Qd[None, :]
        if Qd.shape[1] != self.D:
            raise ValueError("Q [nq, D] must match the index dimension")
        nq = Qd.shape[0]
        ws = self._workspace(nq, k, shards)
        kt = exchange_len(k, shards)
        bm_top = torch.empty((nq, kt), dtype=torch.float32, device=Qd.device) if shards > 1 else None
        _lib.check(_lib.load().qrag_search_tc_begin(_ptr(Qd), nq, _ptr(self.Xb), self.N, self.D, k, self.metric, shards,
                                                    _ptr(bm_top), _ptr(ws), ws.numel(), _stream()))
        self._phase = (Qd, k, ws, shards, kt)
        return bm_top

    def tc_filter(self, bm_top_all: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """Phase 2: bm_top_all [G, nq, kt] from every shard -> hist [nq, HIST_BINS] int32, this shard's survivor
        histogram (to be summed over the shards); a single shard keeps it in the workspace and returns None."""
        Qd, k, ws, shards, kt = self._phase
        nq = Qd.shape[0]
        bm = hist = None
        if shards > 1:
            bm = bm_top_all.contiguous()
            if tuple(bm.shape) != (shards, nq, kt):
                raise ValueError(f"bm_top_all must be [{shards}, {nq}, {kt}]")
            hist = torch.empty((nq, HIST_BINS), dtype=torch.int32, device=Qd.device)
        _lib.check(_lib.load().qrag_search_tc_filter(nq, _ptr(self.Xb), _ptr(self.aux), self.N, self.D, k, self.metric,
                                                     _ptr(bm), shards, _ptr(hist), _ptr(ws), ws.numel(), _stream()))
        return hist

    def _hist(self, hist_all):
        Qd, k, ws, shards, kt = self._phase
        if shards == 1:
            return None
        h = hist_all.contiguous()
        if tuple(h.shape) != (Qd.shape[0], HIST_BINS) or h.dtype != torch.int32:
            raise ValueError(f"hist_all must be int32 [{Qd.shape[0]}, {HIST_BINS}]")
        return h

    def tc_finish(self, hist_all: Optional[torch.Tensor]):
        """Phase 3: hist_all = the survivor histograms summed over the shards -> (scores, ids, status): this shard's
        exact, sorted members of the global top-k (ids -1 padded); status != 0 flags a query this shard could not certify."""
        Qd, k, ws, shards, kt = self._phase
        nq = Qd.shape[0]
        h = self._hist(hist_all)
        scores = torch.empty((nq, k), dtype=torch.float64, device=Qd.device)
        ids = torch.empty((nq, k), dtype=torch.int64, device=Qd.device)
        status = torch.zeros(nq, dtype=torch.int32, device=Qd.device)
        _lib.check(_lib.load().qrag_search_tc_finish(_ptr(Qd), nq, _ptr(self.X), self.N, self.D, k, self.metric,
                                                     self.id_base, _ptr(h), shards, _ptr(scores), _ptr(ids),
                                                     _ptr(status), _ptr(ws), ws.numel(), _stream()))
        self._phase = None
        return scores, ids, status

    def tc_finish_packed(self, hist_all: Optional[torch.Tensor], kk: int, pack: torch.Tensor) -> torch.Tensor:
        """Phase 3, packed (include/qrag.h): the shard's list of every query as a record [3 kk + 1] of int64 words --
        header, kk search scores, kk ids, kk amplitude fidelities -- written into ``pack[:nq]``."""
        Qd, k, ws, shards, kt = self._phase
        nq = Qd.shape[0]
        ap = self._hist(hist_all)
        if pack.dtype != torch.int64 or pack.dim() != 2 or pack.shape[0] < nq or pack.shape[1] != 3 * kk + 1 \
                or not pack.is_contiguous():
            raise ValueError(f"pack must be a contiguous int64 [>= {nq}, {3 * kk + 1}] tensor")
        _lib.check(_lib.load().qrag_search_tc_finish_packed(_ptr(Qd), nq, _ptr(self.X), self.N, self.D, k, self.metric,
                                                            self.id_base, _ptr(ap), shards, kk, _ptr(pack), _ptr(ws),
                                                            ws.numel(), _stream()))
        self._phase = None
        return pack

    def search_sharded(self, Q: ArrayLike, k: int, all_gather, all_reduce_sum, shards: int):
        """This shard's part of a search over G shards: ``all_gather(t)`` must return ``[G, *t.shape]`` and
        ``all_reduce_sum(t)`` the sum of ``t`` over the shards.

        Thresholds are global, so the shard filters and rescores only ~ (k + margin) / G rows.
        ``self.aux[0]`` must already hold the maximum |x| over all shards (see sharded.py).
        """
        bm = self.tc_begin(Q, k, shards)
        bm_all = all_gather(bm) if shards > 1 else None
        hist = self.tc_filter(bm_all)
        return self.tc_finish(all_reduce_sum(hist) if shards > 1 else None)

    def search(self, Q: ArrayLike, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.N == 0:
            return search_topk(Q, self.X, k, self.metric, self.id_base)
        Qd, scores, ids, status = self.search_async(Q, k)
        flagged = torch.nonzero(status).flatten()          # synchronises; the certificate is never skipped
        self.last_fallback = int(flagged.numel())
        if self.last_fallback:
            s2, i2 = search_topk(Qd[flagged], self.X, k, self.metric, self.id_base)
            scores[flagged] = s2
            ids[flagged] = i2
        return scores, ids


def exchange_len(k: int, shards: int) -> int:
    """Entries of the lists the shards of a G-way search exchange (and of the packed records): include/qrag.h."""
    n = ctypes.c_int(0)
    _lib.check(_lib.load().qrag_search_tc_exchange_len(int(k), int(shards), ctypes.byref(n)))
    return n.value


def owner_finalize(recv: torch.Tensor, kk: int, k1: int, k2: int, metric, q_base: int, nq: int,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Owner side of the search + rerank exchange (include/qrag.h): recv [G, per, 3 kk + 1] int64 records ->
    out [per, 2 k2 + 1] int64 (k2 fidelity bits, k2 ids, status)."""
    if recv.dtype != torch.int64 or recv.dim() != 3 or recv.shape[2] != 3 * kk + 1 or not recv.is_contiguous():
        raise ValueError("recv must be a contiguous int64 [G, per, 3 kk + 1] tensor")
    G, per = recv.shape[0], recv.shape[1]
    if out is None:
        out = torch.empty((per, 2 * k2 + 1), dtype=torch.int64, device=recv.device)
    _lib.check(_lib.load().qrag_owner_finalize(_ptr(recv), G, per, kk, k1, k2, metric_id(metric), int(q_base), int(nq),
                                               _ptr(out), _stream()))
    return out


def topk_merge(scores: ArrayLike, ids: ArrayLike, k_out: int, metric="l2") -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge per-shard lists [G, nq, k] into [nq, k_out] (id < 0 = padding)."""
    s = _dev(scores, torch.float64)
    i = _dev(ids, torch.int64)
    if s.dim() != 3 or s.shape != i.shape:
        raise ValueError("scores and ids must both be [G, nq, k]")
    G, nq, k = s.shape
    out_s = torch.empty((nq, k_out), dtype=torch.float64, device=s.device)
    out_i = torch.empty((nq, k_out), dtype=torch.int64, device=s.device)
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.qrag_topk_merge_workspace(G, nq, k, k_out, ctypes.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=s.device) if nbytes.value else None
    _lib.check(lib.qrag_topk_merge(_ptr(s), _ptr(i), G, nq, k, k_out, metric_id(metric), _ptr(out_s), _ptr(out_i),
                                   _ptr(ws), nbytes.value, _stream()))
    return out_s, out_i


# ---------------------------------------------------------------------------
class HostRerankPipeline:
    """End-to-end amplitude-encoded rerank from HOST buffers (what a serving process calls).

    Queries are processed in ``chunks`` slices on two CUDA streams so the host->device
    copy of slice i+1 overlaps the kernel of slice i; results come back in pinned host
    tensors.  ``__call__`` returns after everything has landed on the host.
    """

    def __init__(self, nq: int, C: int, D: int, top_k: int, n_qubits: Optional[int] = None, chunks: int = 4):
        dev = _device()
        self.nq, self.C, self.D, self.k = nq, C, D, top_k
        self.n = qubits_for(D) if n_qubits is None else n_qubits
        self.chunks = max(1, min(chunks, nq))
        self.step = -(-nq // self.chunks)
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        self.dQ = [torch.empty((self.step, D), dtype=torch.float32, device=dev) for _ in range(2)]
        self.dC = [torch.empty((self.step, C, D), dtype=torch.float32, device=dev) for _ in range(2)]
        self.dS = torch.empty((nq, top_k), dtype=torch.float64, device=dev)
        self.dP = torch.empty((nq, top_k), dtype=torch.int32, device=dev)
        self.hS = torch.empty((nq, top_k), dtype=torch.float64).pin_memory()
        self.hP = torch.empty((nq, top_k), dtype=torch.int32).pin_memory()
        self.h2d_bytes = nq * D * 4 + nq * C * D * 4
        self.d2h_bytes = nq * top_k * (8 + 4)
        self.launches_per_call = self.chunks

    def __call__(self, Q_host: torch.Tensor, cand_host: torch.Tensor):
        lib = _lib.load()
        cur = torch.cuda.current_stream()
        for s in self.streams:
            s.wait_stream(cur)
        for c in range(self.chunks):
            a, b = c * self.step, min(self.nq, (c + 1) * self.step)
            if a >= b:
                break
            s = self.streams[c & 1]
            dq, dc = self.dQ[c & 1][: b - a], self.dC[c & 1][: b - a]
            with torch.cuda.stream(s):
                dq.copy_(Q_host[a:b], non_blocking=True)
                dc.copy_(cand_host[a:b], non_blocking=True)
                _lib.check(lib.qrag_amp_rerank(_ptr(dq), b - a, _ptr(dc), None, 0, None, self.C, self.D, self.n, self.k,
                                               _ptr(self.dS[a:b]), _ptr(self.dP[a:b]), None,
                                               ctypes.c_void_p(s.cuda_stream)))
                self.hS[a:b].copy_(self.dS[a:b], non_blocking=True)
                self.hP[a:b].copy_(self.dP[a:b], non_blocking=True)
        for s in self.streams:
            s.synchronize()
        return self.hS, self.hP


class HostIdRerankPipeline:
    """End-to-end rerank of candidates named by ID against a corpus resident on the GPU.

    This is the serving shape behind a retrieval step (faiss-style ids in, reranked ids out): per batch the host
    sends the queries [nq, D] fp32 and the candidate ids [nq, C] int64 from pinned memory, the kernel gathers the rows
    from the resident corpus, and (score, id) top-k come back to pinned host memory.  One batch is ONE library call
    (include/qrag.h: qrag_amp_rerank_host enqueues both copies in, the fused kernel and both copies out).

    ``submit(Q_host, idx_host)`` queues a batch on the next of ``depth`` slots (each with its own stream, device
    workspace and pinned outputs) and returns a ticket; ``result(ticket)`` waits for that batch only and returns its
    pinned (scores, ids).  With two or more batches in flight the host->device copy of one overlaps the kernel of
    another: for batches the size of BASELINE config 2 (2.3 MB in, ~45 us of copy, ~40 us of kernel) the steady state is
    bound by the copy.  ``__call__`` is submit + result: one batch, host waits (the latency form).
    """

    def __init__(self, X: ArrayLike, nq: int, C: int, top_k: int, n_qubits: Optional[int] = None, depth: int = 3):
        self.X = _dev(X, torch.float32)
        dev = self.X.device
        self.nq, self.C, self.D, self.k = nq, C, self.X.shape[1], top_k
        self.n = qubits_for(self.D) if n_qubits is None else n_qubits
        self.depth = max(1, int(depth))
        nbytes = ctypes.c_size_t(0)
        _lib.check(_lib.load().qrag_amp_rerank_host_workspace(nq, C, self.D, top_k, ctypes.byref(nbytes)))
        self._slots = []
        for _ in range(self.depth):
            self._slots.append({"stream": torch.cuda.Stream(device=dev), "event": torch.cuda.Event(),
                                "ws": torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=dev),
                                "hS": torch.empty((nq, top_k), dtype=torch.float64).pin_memory(),
                                "hO": torch.empty((nq, top_k), dtype=torch.int64).pin_memory(), "busy": False})
        self._next = 0
        self.h2d_bytes = nq * self.D * 4 + nq * C * 8
        self.d2h_bytes = nq * top_k * 16
        self.launches_per_call = 1

    def submit(self, Q_host: torch.Tensor, idx_host: torch.Tensor) -> int:
        """Queue one batch; at most ``depth`` may be outstanding (collect the oldest with ``result`` first)."""
        if Q_host.is_cuda or idx_host.is_cuda:
            raise ValueError("HostIdRerankPipeline takes HOST tensors (use amp_rerank for device tensors)")
        if Q_host.dtype != torch.float32 or tuple(Q_host.shape) != (self.nq, self.D) or not Q_host.is_contiguous():
            raise ValueError(f"Q_host must be a contiguous float32 [{self.nq}, {self.D}] tensor")
        if idx_host.dtype != torch.int64 or tuple(idx_host.shape) != (self.nq, self.C) or not idx_host.is_contiguous():
            raise ValueError(f"idx_host must be a contiguous int64 [{self.nq}, {self.C}] tensor")
        t = self._next
        slot = self._slots[t]
        if slot["busy"]:
            raise RuntimeError(f"{self.depth} batches already in flight: call result() on the oldest ticket first")
        self._next = (t + 1) % self.depth
        ws = slot["ws"]
        _lib.check(_lib.load().qrag_amp_rerank_host(
            Q_host.data_ptr(), self.nq, idx_host.data_ptr(), self.C, _ptr(self.X), self.X.shape[0], self.D, self.n, self.k,
            _ptr(ws), ws.numel(), slot["hS"].data_ptr(), slot["hO"].data_ptr(),
            ctypes.c_void_p(slot["stream"].cuda_stream)))
        slot["event"].record(slot["stream"])
        slot["busy"] = True
        slot["inputs"] = (Q_host, idx_host)          # keep the host buffers alive until the copies have run
        return t

    def result(self, ticket: int):
        """Pinned (scores [nq, k] fp64, ids [nq, k] int64) of that batch; valid until the slot is submitted to again."""
        slot = self._slots[ticket]
        if not slot["busy"]:
            raise RuntimeError("no batch outstanding on this ticket")
        slot["event"].synchronize()
        slot["busy"] = False
        slot["inputs"] = None
        return slot["hS"], slot["hO"]

    def __call__(self, Q_host: torch.Tensor, idx_host: torch.Tensor):
        return self.result(self.submit(Q_host, idx_host))
