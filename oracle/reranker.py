"""Restatement of the reranker API semantics (selection heuristic, validation,
ordering, top_k rules).  TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Cites ``/root/reference/src/reranker/controller.py`` and ``classical.py``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

QUANTUM_KEYWORDS = (  # controller.py:25-36
    "advertisement", "ad", "sponsor", "commercial", "promotion",
    "product", "brand", "discount", "offer", "deal",
)


def select_reranker(query: str, complexity_threshold: int = 8) -> str:
    """controller.py:42-67.  Keywords match as *substrings* of lower-cased words."""
    words = query.lower().split()
    hits = 0
    for w in words:
        for kw in QUANTUM_KEYWORDS:
            if kw in w:
                hits += 1
                break
    if len(words) > complexity_threshold or hits > 0:
        return "quantum"
    return "classical"


def dispatch(reranker_type: str, query: str, complexity_threshold: int = 8) -> str:
    """controller.py:88-98: "auto" -> heuristic; exactly "quantum" -> quantum; anything else -> classical."""
    chosen = select_reranker(query, complexity_threshold) if reranker_type == "auto" else reranker_type
    return "quantum" if chosen == "quantum" else "classical"


def classical_inputs_valid(query, contents: Sequence, is_document: Sequence[bool]) -> bool:
    """classical.py:169-187 (documents given as their ``content`` plus an isinstance flag)."""
    if not isinstance(query, str) or not query.strip():
        return False
    if not isinstance(contents, list) or not contents:
        return False
    for ok, c in zip(is_document, contents):
        if not ok or not c:
            return False
    return True


def classical_order(scores: Sequence[float], top_k: Optional[int]) -> List[int]:
    """classical.py:301-308: stable descending sort; slice only if top_k is not None and > 0."""
    order = sorted(range(len(scores)), key=lambda i: scores[i], reverse=True)
    if top_k is not None and top_k > 0:
        order = order[:top_k]
    return order
