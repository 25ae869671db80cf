"""NumPy restatement of flat-index search, top-k merge and the ``IxF2`` file format.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  **Parity unpinned**: the
reference only *writes* a ``faiss.IndexFlatL2`` (``mcp/server/tools/store_in_faiss.py:99-109``)
and never calls ``.search``; faiss-cpu (pinned 1.11.0.post1, ``poetry.lock:699-700``)
is not installed here.  The canonical definition used by the whole build is:

    score = metric evaluated from the fp32 inputs with fp64 arithmetic;
    order = (similarity descending | L2^2 distance ascending, id ascending).

FAISS's published contract (``IndexFlatL2.search`` returns squared L2 distances
in ascending order with int64 labels, ``-1`` padding when k > ntotal) is kept.
"""
from __future__ import annotations

import struct
from typing import Tuple

import numpy as np

METRIC_IP = 0       # inner product, larger is better
METRIC_L2 = 1       # squared L2, smaller is better (faiss METRIC_L2 == 1, the fixture's metric)
METRIC_COSINE = 2   # cosine similarity, larger is better


def exact_scores(Q: np.ndarray, X: np.ndarray, metric: int) -> np.ndarray:
    """[nq, N] fp64 scores from fp32 inputs."""
    Q64 = np.asarray(Q, dtype=np.float32).astype(np.float64)
    X64 = np.asarray(X, dtype=np.float32).astype(np.float64)
    if metric == METRIC_IP:
        return Q64 @ X64.T
    if metric == METRIC_L2:
        out = np.empty((Q64.shape[0], X64.shape[0]), dtype=np.float64)
        for i in range(Q64.shape[0]):
            diff = X64 - Q64[i]
            out[i] = np.einsum("nd,nd->n", diff, diff)
        return out
    if metric == METRIC_COSINE:
        ip = Q64 @ X64.T
        den = np.einsum("qd,qd->q", Q64, Q64)[:, None] * np.einsum("nd,nd->n", X64, X64)[None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(den > 0, ip / np.sqrt(np.where(den > 0, den, 1.0)), 0.0)
    raise ValueError(f"unknown metric {metric}")


def topk_from_scores(scores: np.ndarray, k: int, metric: int, id_base: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Canonical top-k: (best first, id ascending on ties); pads with id -1."""
    nq, n = scores.shape
    ids = np.broadcast_to(np.arange(n, dtype=np.int64), (nq, n))
    key = scores if metric == METRIC_L2 else -scores
    order = np.lexsort((ids, key), axis=1)[:, :k]
    out_s = np.take_along_axis(scores, order, axis=1)
    out_i = order.astype(np.int64) + id_base
    if k > n:
        pad_s = np.full((nq, k - n), np.inf if metric == METRIC_L2 else -np.inf)
        out_s = np.concatenate([out_s, pad_s], axis=1)
        out_i = np.concatenate([out_i, np.full((nq, k - n), -1, dtype=np.int64)], axis=1)
    return out_s, out_i


def exact_search(Q: np.ndarray, X: np.ndarray, k: int, metric: int = METRIC_L2, id_base: int = 0):
    return topk_from_scores(exact_scores(Q, X, metric), k, metric, id_base)


def exact_search_chunked(Q: np.ndarray, X: np.ndarray, k: int, metric: int = METRIC_L2, id_base: int = 0,
                         chunk: int = 1 << 16):
    """``exact_search`` for a corpus too large to score at once: chunks of rows, running canonical top-k
    (each chunk's list merged into the running one with ``merge_topk``; same result as one pass)."""
    run_s = run_i = None
    for lo in range(0, X.shape[0], chunk):
        s, i = exact_search(Q, X[lo:lo + chunk], k, metric, id_base + lo)
        if run_s is None:
            run_s, run_i = s, i
        else:
            run_s, run_i = merge_topk(np.stack([run_s, s]), np.stack([run_i, i]), k, metric)
    if run_s is None:
        return exact_search(Q, X, k, metric, id_base)
    return run_s, run_i


def merge_topk(scores: np.ndarray, ids: np.ndarray, k_out: int, metric: int) -> Tuple[np.ndarray, np.ndarray]:
    """Merge per-shard lists ``[G, nq, k]`` into the global ``[nq, k_out]`` list.

    Entries with ``id < 0`` are padding and always sort last.
    """
    g, nq, k = scores.shape
    s = np.transpose(scores, (1, 0, 2)).reshape(nq, g * k)
    i = np.transpose(ids, (1, 0, 2)).reshape(nq, g * k)
    key = s if metric == METRIC_L2 else -s
    key = np.where(i < 0, np.inf, key)
    tie = np.where(i < 0, np.iinfo(np.int64).max, i)
    order = np.lexsort((tie, key), axis=1)[:, :k_out]
    out_s = np.take_along_axis(s, order, axis=1)
    out_i = np.take_along_axis(i, order, axis=1)
    if k_out > g * k:
        pad = k_out - g * k
        out_s = np.concatenate([out_s, np.full((nq, pad), np.inf if metric == METRIC_L2 else -np.inf)], axis=1)
        out_i = np.concatenate([out_i, np.full((nq, pad), -1, dtype=np.int64)], axis=1)
    return out_s, out_i


# --------------------------------------------------------------------------
# The packed search + rerank exchange (include/qrag.h: qrag_search_tc_finish_packed /
# qrag_owner_finalize; builder-defined, the reference is single-process).  A record is
# 3 kk + 1 int64 words: header (entries | bad << 32), kk score bits, kk ids, kk fidelity bits.
# --------------------------------------------------------------------------
def pack_records(scores: np.ndarray, ids: np.ndarray, fid: np.ndarray, kk: int, metric: int,
                 bad: np.ndarray = None) -> np.ndarray:
    """Per-shard sorted lists [nq, k] (ids -1 padded) -> records [nq, 3 kk + 1], cut to kk entries."""
    nq, k = ids.shape
    valid = (ids >= 0).sum(axis=1)
    rec = np.zeros((nq, 3 * kk + 1), dtype=np.int64)
    pad_s = np.inf if metric == METRIC_L2 else -np.inf
    s = np.full((nq, kk), pad_s, dtype=np.float64)
    i = np.full((nq, kk), -1, dtype=np.int64)
    f = np.full((nq, kk), -np.inf, dtype=np.float64)
    w = min(k, kk)
    keep = np.arange(w)[None, :] < np.minimum(valid, kk)[:, None]
    s[:, :w] = np.where(keep, scores[:, :w], pad_s)
    i[:, :w] = np.where(keep, ids[:, :w], -1)
    f[:, :w] = np.where(keep, fid[:, :w], -np.inf)
    flag = (valid > kk).astype(np.int64)
    if bad is not None:
        flag |= (np.asarray(bad) != 0).astype(np.int64)
    rec[:, 0] = np.minimum(valid, kk) | (flag << 32)
    rec[:, 1:1 + kk] = s.view(np.int64)
    rec[:, 1 + kk:1 + 2 * kk] = i
    rec[:, 1 + 2 * kk:] = f.view(np.int64)
    return rec


def owner_finalize(recv: np.ndarray, kk: int, k1: int, k2: int, metric: int, q_base: int = 0,
                   nq: int = None) -> np.ndarray:
    """recv [G, per, 3 kk + 1] -> out [per, 2 k2 + 1]: merge the G lists in (score, id) order, keep the global
    top-k1, order by (fidelity desc, merged position asc) -- Python's stable sorted(reverse=True),
    quantum.py:70-76 -- and keep k2; last word = status (some shard flagged the query)."""
    G, per, _ = recv.shape
    out = np.empty((per, 2 * k2 + 1), dtype=np.int64)
    for j in range(per):
        if nq is not None and q_base + j >= nq:
            out[j, :k2] = np.full(k2, -np.inf).view(np.int64)
            out[j, k2:2 * k2] = -1
            out[j, 2 * k2] = 0
            continue
        entries, bad = [], 0
        for g in range(G):
            r = recv[g, j]
            n = min(max(int(r[0] & 0xFFFFFFFF), 0), kk)
            bad |= int(r[0] >> 32 != 0)
            s = r[1:1 + kk].view(np.float64)
            f = r[1 + 2 * kk:].view(np.float64)
            for e in range(n):
                key = s[e] if metric == METRIC_L2 else -s[e]
                entries.append((key, int(r[1 + kk + e]), f[e]))
        entries.sort(key=lambda t: (t[0], t[1]))
        members = entries[:k1]
        order = sorted(range(len(members)), key=lambda p: members[p][2], reverse=True)[:k2]     # stable
        fo = np.full(k2, -np.inf)
        io = np.full(k2, -1, dtype=np.int64)
        for t, p_ in enumerate(order):
            fo[t], io[t] = members[p_][2], members[p_][1]
        out[j, :k2] = fo.view(np.int64)
        out[j, k2:2 * k2] = io
        out[j, 2 * k2] = bad
    return out


# --------------------------------------------------------------------------
# IxF2 (faiss IndexFlatL2 / IndexFlat) on-disk layout, as written by
# faiss.write_index at store_in_faiss.py:109 and observed in the fixture:
#   "IxF2" | int32 d | int64 ntotal | int64 dummy | int64 dummy | uint8 is_trained
#   | int32 metric_type | uint64 count (= d * ntotal) | count * fp32 (row major, LE)
# "IxFI" is the same layout with metric_type 0 (inner product).
# --------------------------------------------------------------------------
_IXF_HEADER = struct.Struct("<4siqqqBiQ")   # 45 bytes


def read_ixf(buf: bytes):
    magic, d, ntotal, _d1, _d2, is_trained, metric_type, count = _IXF_HEADER.unpack_from(buf, 0)
    if magic not in (b"IxF2", b"IxFI"):
        raise ValueError(f"not a flat faiss index: {magic!r}")
    if count != d * ntotal:
        raise ValueError("corrupt flat index: count != d * ntotal")
    x = np.frombuffer(buf, dtype="<f4", count=count, offset=_IXF_HEADER.size).reshape(ntotal, d)
    return {"magic": magic, "d": d, "ntotal": ntotal, "is_trained": bool(is_trained),
            "metric_type": metric_type, "vectors": x}


def write_ixf(x: np.ndarray, metric_type: int = 1) -> bytes:
    x = np.ascontiguousarray(x, dtype="<f4")
    ntotal, d = x.shape
    magic = b"IxF2" if metric_type == 1 else b"IxFI"
    head = _IXF_HEADER.pack(magic, d, ntotal, 1 << 20, 1 << 20, 1, metric_type, d * ntotal)
    return head + x.tobytes()
