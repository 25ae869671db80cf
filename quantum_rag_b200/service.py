"""Search -> select -> rerank over a resident index, and the HTTP routes around it.

The reference serves one route, ``POST /rerank`` (app.py:56-77): strings in, the controller's dict out.
It has no retrieval step; its ingest tool only writes a flat faiss index (store_in_faiss.py:99-109).
This module wires the pieces of the hot path together the way the reference's own comments anticipate
(quantum.py:93,99,156: "a real implementation would use embeddings"):

    embeddings [nq, d] --K3/K4 search--> top-k1 ids --controller--> quantum | classical
        quantum    amplitude-encoded fidelity of the k1 rows, gathered from the resident corpus (K2),
                   stable order (fidelity desc, search position asc), top-k2
        classical  the search order itself (the index metric is the classical score), top-k2

``create_app`` returns a FastAPI app with the reference's ``/rerank`` (same request / response /
error shape) and the new ``/search_rerank``.  Nothing here computes on the CPU: without libqrag and a
CUDA device the calls raise (the routes report ``{"error": ...}`` like app.py:75-77).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence

import numpy as np

try:                                                      # the request schemas live at module level so that FastAPI
    from pydantic import BaseModel                        # can resolve them from the (postponed) annotations
except ImportError:                                       # pragma: no cover - the service class works without pydantic
    BaseModel = object


class DocumentRequest(BaseModel):                         # app.py:23-26
    id: str
    content: str
    source: Optional[str] = None


class RerankRequest(BaseModel):                           # app.py:29-33
    query: str
    documents: List[DocumentRequest]
    reranker_type: Optional[str] = "auto"
    top_k: Optional[int] = 5


class SearchRerankRequest(BaseModel):
    embeddings: List[List[float]]
    queries: Optional[List[str]] = None
    k1: int = 20
    top_k: Optional[int] = 5
    reranker_type: Optional[str] = "quantum"


class SearchRerankService:
    """Brute-force search over a ``FlatIndex`` followed by the reranker the controller selects."""

    def __init__(self, index, controller=None, layers: int = 0, n_qubits: Optional[int] = None):
        self.index = index
        self.controller = controller
        self.layers = int(layers)
        self.n_qubits = n_qubits

    def _choices(self, nq: int, reranker_type: str, queries: Optional[Sequence[str]]) -> List[str]:
        if reranker_type == "auto":                       # controller.py:88-90
            if queries is None:
                raise ValueError('reranker_type="auto" needs the query texts (the heuristic reads words)')
            if self.controller is None:
                raise ValueError('reranker_type="auto" needs a RerankerController')
            return self.controller.select_rerankers(list(queries))
        # controller.py:92-99: exactly "quantum" -> quantum, anything else -> classical
        return ["quantum" if reranker_type == "quantum" else "classical"] * nq

    def search_rerank(self, embeddings, k1: int = 20, k2: Optional[int] = 5, reranker_type: str = "quantum",
                      queries: Optional[Sequence[str]] = None) -> List[Dict[str, Any]]:
        """One result dict per query: ``{"query", "reranker_used", "documents": [{"id", "label", "score",
        "search_score", "search_rank"}]}``, best first.  ``k2=None`` keeps all k1 (quantum.py:74 semantics)."""
        import torch
        from . import api
        emb = np.asarray(embeddings, dtype=np.float32) if not isinstance(embeddings, torch.Tensor) else embeddings
        if emb.ndim == 1:
            emb = emb[None, :]
        nq = int(emb.shape[0])
        if queries is not None and len(queries) != nq:
            raise ValueError("one query text per embedding expected")
        if emb.shape[1] != self.index.d:
            raise ValueError(f"embeddings have dimension {emb.shape[1]}, the index {self.index.d}")
        k1 = max(1, min(int(k1), max(self.index.ntotal, 1)))
        k_out = k1 if k2 is None else max(0, min(int(k2), k1))
        choices = self._choices(nq, reranker_type, queries)
        dist, ids = self.index.search(emb, k1)                                   # [nq, k1] fp64 / int64, -1 padded
        Q = emb if isinstance(emb, torch.Tensor) else torch.from_numpy(emb)
        Q = Q.to(ids.device)
        sign = -1.0 if self.index.metric == "l2" else 1.0                        # higher is better in the response
        ids_h, dist_h = ids.cpu().numpy(), dist.cpu().numpy()
        qsel = [i for i, c in enumerate(choices) if c == "quantum"]
        q_scores = q_pos = None
        if qsel and k_out > 0:
            sel = torch.as_tensor(qsel, device=ids.device)
            s, pos, _ = api.quantum_rerank_batch(Q[sel], X=self.index._tc.X, idx=ids[sel], top_k=k_out,
                                                 n_qubits=self.n_qubits, layers=self.layers)
            q_scores, q_pos = s.cpu().numpy(), pos.cpu().numpy()
        labels = self.index.labels
        out = []
        for qi in range(nq):
            docs = []
            if choices[qi] == "quantum" and k_out > 0:
                row = qsel.index(qi)
                order, scores = q_pos[row].tolist(), q_scores[row].tolist()
            else:
                order = list(range(k_out))
                scores = [sign * float(dist_h[qi, p]) for p in order]
            for p, sc in zip(order, scores):
                doc_id = int(ids_h[qi, p])
                if doc_id < 0:                                                   # fewer than k1 rows in the index
                    continue
                docs.append({"id": doc_id, "label": labels[doc_id] if labels is not None else None,
                             "score": float(sc), "search_score": float(dist_h[qi, p]), "search_rank": int(p)})
            out.append({"query": queries[qi] if queries is not None else None, "reranker_used": choices[qi],
                        "documents": docs})
        return out


def create_app(controller=None, service: Optional[SearchRerankService] = None):
    """FastAPI app: the reference's ``POST /rerank`` and ``GET /`` (app.py:56-96) plus ``POST /search_rerank``."""
    from fastapi import FastAPI

    from .reranker import Document, RerankerController

    if BaseModel is object:
        raise RuntimeError("pydantic is required for the HTTP routes")
    ctl = controller if controller is not None else RerankerController()
    app = FastAPI(title="Quantum RAG Reranker", version="0.1.0",
                  description="B200-native reranking hot path behind the reference's API")

    @app.post("/rerank")
    async def rerank_documents(request: RerankRequest):
        try:
            documents = [Document(d.id, d.content, d.source) for d in request.documents]
            return ctl.rerank(query=request.query, documents=documents, top_k=request.top_k,
                              reranker_type=request.reranker_type)
        except Exception as exc:                          # app.py:75-77
            return {"error": str(exc)}

    @app.post("/search_rerank")
    async def search_rerank(request: SearchRerankRequest):
        try:
            if service is None:
                raise RuntimeError("no index loaded: create_app(service=SearchRerankService(FlatIndex.read(...)))")
            return {"results": service.search_rerank(request.embeddings, k1=request.k1, k2=request.top_k,
                                                     reranker_type=request.reranker_type, queries=request.queries)}
        except Exception as exc:
            return {"error": str(exc)}

    @app.get("/")
    async def root():
        return {"message": "Quantum RAG Reranker API", "docs_url": "/docs", "version": "0.1.0",
                "endpoints": {"rerank": "POST /rerank - rerank documents (reference route)",
                              "search_rerank": "POST /search_rerank - embeddings in, searched and reranked ids out"}}

    return app
