"""Drop-in for ``src/reranker/controller.py`` of jon-fox/quantum-rag.

Pure host logic, kept behaviourally identical to the reference
(controller.py:11-104): a keyword / length heuristic picks the reranker, and the
result is wrapped in ``{"documents", "reranker_used", "query"}``.
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence

from .classical import ClassicalReranker, Document
from .quantum import QuantumReranker

_DEFAULT_KEYWORDS = (
    "advertisement", "ad", "sponsor", "commercial", "promotion",
    "product", "brand", "discount", "offer", "deal",
)


class RerankerController:
    """Decides between the quantum and the classical reranker and runs it."""

    def __init__(self, config: Dict[str, Any] = None):
        self.config = config or {}
        self.classical_reranker = ClassicalReranker(self.config.get("classical_config", {}))
        self.quantum_reranker = QuantumReranker(self.config.get("quantum_config", {}))
        self.quantum_keywords = list(_DEFAULT_KEYWORDS)             # controller.py:25-36
        self.complexity_threshold = self.config.get("complexity_threshold", 8)

    def select_reranker(self, query: str) -> str:
        """"quantum" if the query is long or any word *contains* a keyword (controller.py:42-67)."""
        words = query.lower().split()
        keyword_words = 0
        for word in words:
            if any(kw in word for kw in self.quantum_keywords):
                keyword_words += 1
        if len(words) > self.complexity_threshold or keyword_words > 0:
            return "quantum"
        return "classical"

    def select_rerankers(self, queries: Sequence[str]) -> List[str]:
        """Batched ``select_reranker`` (convenience for batch callers)."""
        return [self.select_reranker(q) for q in queries]

    def rerank(self, query: str, documents: List[Document], top_k: int = None,
               reranker_type: str = "auto") -> Dict[str, Any]:
        """controller.py:69-104: "auto" -> heuristic, exactly "quantum" -> quantum, anything else -> classical."""
        choice = self.select_reranker(query) if reranker_type == "auto" else reranker_type
        if choice == "quantum":
            ranked = self.quantum_reranker.rerank(query, documents, top_k)
            used = "quantum"
        else:
            ranked = self.classical_reranker.rerank(query, documents, top_k)
            used = "classical"
        return {"documents": ranked, "reranker_used": used, "query": query}
