"""Drop-in for ``src/reranker/quantum.py`` of jon-fox/quantum-rag, on a B200.

``QuantumReranker(config).rerank(query, documents, top_k)`` keeps the reference's
behaviour (quantum.py:44-78): empty input -> ``[]``; every document scored by the
state fidelity of the reference circuit (or the constant 0.5 for any other
``method``, quantum.py:134-136); stable descending sort; ``[:top_k]`` whenever
``top_k is not None``.  What runs underneath is one batched launch of the
hand-written statevector kernel instead of two Qiskit ``execute`` calls per
document.

Extensions (all default to the reference behaviour):
    config["encoding"]  "angle" (reference circuit on the text-hash embedding) or
                        "amplitude" (real embeddings from config["embedder"] /
                        Document.metadata["embedding"], as quantum.py:93,99,156 anticipate)
    config["layers"]    repetitions of the circuit block (1 = reference; amplitude: 0)
    config["embedding_backend"]  "host" (NumPy legacy MT19937, bit-identical to the
                        reference's np.random.seed stream) or "device" (MT19937 kernel)
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .classical import ClassicalReranker, Document

# Kept for import compatibility (quantum.py:11-17).  The statevector engine is built
# into libqrag, so the quantum path is always "available"; when the library or a CUDA
# device is missing, rerank() raises instead of silently going classical.
QISKIT_AVAILABLE = True


def _char_sum(text: str) -> int:
    """sum(ord(c)) -- quantum.py:182.  ASCII text: the byte sum (one C loop); anything else: vectorised over the
    UTF-32 code units."""
    if not text:
        return 0
    if text.isascii():
        return sum(text.encode("ascii"))
    return int(np.frombuffer(text.encode("utf-32-le", "surrogatepass"), dtype="<u4").sum(dtype=np.uint64))


class QuantumReranker:
    """Quantum-circuit similarity reranker backed by batched statevector kernels."""

    def __init__(self, config: Dict[str, Any] = None):
        self.config = config or {}
        self.method = self.config.get("method", "state_fidelity")
        self.n_qubits = self.config.get("n_qubits", 4)
        self.encoding = self.config.get("encoding", "angle")
        self.layers = self.config.get("layers", 1 if self.encoding == "angle" else 0)
        self.embedding_backend = self.config.get("embedding_backend", "host")
        self.embedder = self.config.get("embedder")
        self.embedding_key = self.config.get("embedding_key", "embedding")
        # the reference builds a ClassicalReranker from the same dict (quantum.py:37); kept as an attribute
        self.classical_fallback = ClassicalReranker(config)
        # memo of the text-hash embeddings, keyed by sum(ord(c)); bounded (LRU) so that a long-running /rerank service
        # does not grow without limit -- an evicted seed is simply recomputed (same bits: the stream is seeded)
        self._embedding_memo: "OrderedDict[int, np.ndarray]" = OrderedDict()
        self._memo_cap = int(self.config.get("embedding_memo_size", 65536))
        if self.encoding == "angle" and not 1 <= int(self.n_qubits) <= 12:
            # the reference accepts any n its simulator can hold; the statevector kernels stage 2^n complex128
            # amplitudes in shared memory, which ends at n = 12 (QRAG_MAX_QUBITS).  Say so up front, not at rerank().
            raise ValueError(f"n_qubits={self.n_qubits}: the B200 statevector kernels support 1 <= n_qubits <= 12")

    # ------------------------------------------------------------------ API
    def rerank(self, query: str, documents: List[Document], top_k: int = None) -> List[Tuple[Document, float]]:
        if not documents:                                           # quantum.py:63-64
            return []
        if (self.method == "state_fidelity" and self.encoding == "amplitude" and isinstance(top_k, int)
                and 0 < top_k < len(documents) <= 4096):
            # only the top_k are wanted: the fused kernels score, rank and cut on the device (layers == 0: one
            # streaming launch; layers >= 1 at 10 qubits: complex64 filter + complex128 certification) -- the same
            # documents, order and score bits as ranking everything and slicing
            from .. import api
            q, x = self._real_embeddings(query, documents)
            n = max(self.n_qubits, api.qubits_for(q.shape[1])) if "n_qubits" not in self.config else self.n_qubits
            scores, pos, _ = api.quantum_rerank_batch(q, cand=x[None, :, :], top_k=top_k, n_qubits=n, layers=self.layers)
            return [(documents[i], s) for i, s in zip(pos[0].cpu().tolist(), scores[0].cpu().tolist())]
        order, scores = self._rank(query, documents)
        ranked = [(documents[i], s) for i, s in zip(order, scores)]
        if top_k is not None:                                       # quantum.py:75-76 (plain slice)
            ranked = ranked[:top_k]
        return ranked

    def _quantum_score_documents(self, query: str, documents: List[Document]) -> List[Tuple[Document, float]]:
        """Scores in input order (quantum.py:80-106)."""
        scores = self._scores(query, documents).cpu().tolist()
        return list(zip(documents, scores))

    # ------------------------------------------------------------ internals
    def _mock_embedding(self, text: str) -> np.ndarray:
        """quantum.py:169-185 without touching NumPy's global RNG; memoised by seed."""
        seed = _char_sum(text)
        vec = self._embedding_memo.get(seed)
        if vec is None:
            raw = np.random.RandomState(seed).random_sample(self.n_qubits * 2)
            vec = raw / np.linalg.norm(raw)
            self._embedding_memo[seed] = vec
            if len(self._embedding_memo) > self._memo_cap:
                self._embedding_memo.popitem(last=False)
        else:
            self._embedding_memo.move_to_end(seed)
        return vec

    def _text_embeddings(self, query: str, documents: List[Document]):
        import torch
        from .. import api
        if self.embedding_backend == "device":
            seeds = np.array([_char_sum(query)] + [_char_sum(d.content) for d in documents], dtype=np.int64)
            emb = api.mock_embedding(seeds, self.n_qubits)
            return emb[:1], emb[1:]
        # query and documents in ONE array: one host-to-device copy, sliced on the device
        both = np.empty((1 + len(documents), 2 * self.n_qubits), dtype=np.float64)
        both[0] = self._mock_embedding(query)
        for i, doc in enumerate(documents):
            both[i + 1] = self._mock_embedding(doc.content)
        dev = api._dev(both, torch.float64)
        return dev[:1], dev[1:]

    def _real_embeddings(self, query: str, documents: List[Document]):
        docs_e = [d.metadata.get(self.embedding_key) if isinstance(d.metadata, dict) else None for d in documents]
        if any(e is None for e in docs_e):
            if self.embedder is None:
                raise ValueError("encoding='amplitude' needs config['embedder'] or Document.metadata['embedding']")
            docs_e = self.embedder([d.content for d in documents])
        q_e = self.config.get("query_embedding")
        if q_e is None:
            if self.embedder is None:
                raise ValueError("encoding='amplitude' needs config['embedder'] or config['query_embedding']")
            q_e = self.embedder([query])
        q = np.asarray(q_e, dtype=np.float32).reshape(1, -1)
        x = np.asarray(docs_e, dtype=np.float32).reshape(len(documents), -1)
        return q, x

    def _scores(self, query: str, documents: List[Document]):
        """fp64 device tensor [len(documents)] of similarity scores, input order."""
        import torch
        from .. import api
        if self.method != "state_fidelity":                         # quantum.py:134-136
            return torch.full((len(documents),), 0.5, dtype=torch.float64, device=api._device())
        if self.encoding == "angle":
            q, d = self._text_embeddings(query, documents)
            return api.sv_fidelity_angle(q, d, docs_per_query=len(documents), n_qubits=self.n_qubits,
                                         layers=self.layers)
        if self.encoding == "amplitude":
            q, x = self._real_embeddings(query, documents)
            n = max(self.n_qubits, api.qubits_for(q.shape[1])) if "n_qubits" not in self.config else self.n_qubits
            return api.amp_fidelity(q, cand=x[None, :, :], n_qubits=n, layers=self.layers)[0]
        raise ValueError(f"unknown encoding {self.encoding!r}")

    def _rank(self, query: str, documents: List[Document]):
        """(input positions, scores) in final order: score desc, input position asc."""
        from .. import api
        scores = self._scores(query, documents)
        perm, srt = api.sort_scores_host(scores[None, :], None, descending=True)     # one device-to-host copy for both
        return perm[0].tolist(), srt[0].tolist()
