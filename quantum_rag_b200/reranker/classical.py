"""Drop-in for ``src/reranker/classical.py`` of jon-fox/quantum-rag.

Same public surface (``Document``, ``ClassicalReranker(config).rerank(query,
documents, top_k)``) and the same validation / ordering / top_k rules
(reference classical.py:29-42, 169-187, 231-320).

What differs, on purpose:

* ``method`` is honoured.  The reference reads ``config["method"]``
  (classical.py:56) but only ever runs the HuggingFace cross-encoder.  Here
  ``"cosine"``, ``"ip"`` and ``"l2"`` score on the GPU (libqrag brute-force kernels)
  from embeddings supplied by ``config["embedder"]`` (``List[str] -> [n, D]``) or
  ``Document.metadata["embedding"]``.
* ``"cross-encoder"`` (the default) is the one transformer forward pass on the
  path and is out of scope for the kernels.  If ``sentence_transformers`` is
  importable it is used exactly like the reference; otherwise the model counts as
  "failed to load" and ``rerank`` takes the reference's own failure route: every
  document scored 0.5 in the original order (classical.py:218-229, 258-260).
"""
from __future__ import annotations

import logging
import os
import re
import time
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

logger = logging.getLogger(__name__)

EMBEDDING_METHODS = ("cosine", "ip", "l2")
NEUTRAL_SCORE = 0.5


class Document:
    """Value object handed to the rerankers (reference classical.py:29-42)."""

    def __init__(self, id: str, content: str, source: Optional[str] = None,
                 metadata: Optional[Dict[str, Any]] = None):
        self.id = id
        self.content = content
        self.source = source
        self.metadata = metadata or {}

    def __repr__(self) -> str:  # convenience only; the reference has none
        return f"Document(id={self.id!r}, source={self.source!r})"


class ClassicalReranker:
    """Classical reranker: cross-encoder (reference) or GPU cosine / inner-product / L2."""

    def __init__(self, config: Optional[Dict[str, Any]] = None):
        cfg = self.config = config or {}
        self.method = cfg.get("method", "cross-encoder")
        self.model_name = cfg.get("model_name", "cross-encoder/ms-marco-MiniLM-L-6-v2")
        self.batch_size = cfg.get("batch_size", 32)
        self.max_sequence_length = cfg.get("max_sequence_length", 512)
        self.max_retries = cfg.get("max_retries", 3)
        self.timeout = cfg.get("timeout", 30)
        self.device = cfg.get("device", _default_device())
        self.embedder: Optional[Callable[[List[str]], Any]] = cfg.get("embedder")
        self.embedding_key = cfg.get("embedding_key", "embedding")
        self.model = None
        self.model_loaded = False
        if self.method not in EMBEDDING_METHODS:
            self._initialize_model()
        self.score_cache: Dict[str, float] = {}
        self.enable_cache = cfg.get("enable_cache", True)

    # ------------------------------------------------------------------ model
    def _initialize_model(self) -> None:
        """Cross-encoder load with local cache and public fall-back models (classical.py:79-153)."""
        try:
            from sentence_transformers import CrossEncoder  # type: ignore
        except Exception as exc:  # not installed in this image
            logger.warning("sentence_transformers unavailable (%s); cross-encoder reranking disabled", exc)
            return
        cache_root = self.config.get("model_cache_dir", "cross_encoder")
        local_dir = os.path.join(cache_root, self.model_name.replace("/", "_"))
        candidates = []
        if os.path.isdir(local_dir):
            candidates.append((local_dir, self.model_name, False))
        candidates.append((self.model_name, self.model_name, True))
        for fb in ("cross-encoder/ms-marco-TinyBERT-L-2-v2", "cross-encoder/ms-marco-MiniLM-L-2-v2"):
            candidates.append((fb, fb, False))
        for path, name, save in candidates:
            try:
                self.model = CrossEncoder(path, device=self.device)
                self.model_name, self.model_loaded = name, True
                if save:
                    try:
                        os.makedirs(cache_root, exist_ok=True)
                        self.model.save(local_dir)
                    except Exception as exc:
                        logger.warning("could not cache model under %s: %s", local_dir, exc)
                logger.info("loaded cross-encoder %s on %s", name, self.device)
                return
            except Exception as exc:
                logger.warning("loading %s failed: %s", path, exc)

    # ---------------------------------------------------------------- helpers
    def _sanitize_text(self, text: str) -> str:
        """Collapse whitespace, truncate to ~4 chars/token (classical.py:155-167)."""
        if not isinstance(text, str):
            text = str(text)
        text = re.sub(r"\s+", " ", text).strip()
        limit = self.max_sequence_length * 4
        return text[:limit] if len(text) > limit else text

    def _validate_inputs(self, query: str, documents: List[Document]) -> bool:
        """classical.py:169-187."""
        if not isinstance(query, str) or not query.strip():
            logger.error("Query must be a non-empty string")
            return False
        if not isinstance(documents, list) or not documents:
            logger.error("Documents must be a non-empty list")
            return False
        for i, doc in enumerate(documents):
            if not isinstance(doc, Document):
                logger.error("Document at index %d is not a Document instance", i)
                return False
            if not getattr(doc, "content", None):
                logger.error("Document at index %d has empty content", i)
                return False
        return True

    def _get_cache_key(self, query: str, doc_content: str) -> str:
        return f"{hash(query)}_{hash(doc_content)}"

    def _predict_with_retries(self, inputs: List[Tuple[str, str]]) -> np.ndarray:
        """classical.py:193-216."""
        if self.model is None:
            raise RuntimeError("Model is not initialized")
        for attempt in range(self.max_retries):
            try:
                return self.model.predict(inputs, show_progress_bar=False)
            except Exception as exc:
                logger.warning("Prediction attempt %d failed: %s", attempt + 1, exc)
                if attempt == self.max_retries - 1:
                    raise
                time.sleep(0.5 * (attempt + 1))
        raise RuntimeError("All retry attempts failed")

    def _handle_reranker_failure(self, query: str, documents: List[Document]) -> List[Tuple[Document, float]]:
        """Neutral 0.5 scores, original order (classical.py:218-229)."""
        logger.error("reranking failed for %d documents (query %.100s...) - returning neutral scores",
                     len(documents), query)
        return [(doc, NEUTRAL_SCORE) for doc in documents]

    # ----------------------------------------------------------- GPU methods
    def _embeddings(self, query: str, documents: Sequence[Document]):
        docs_e = [d.metadata.get(self.embedding_key) if isinstance(d.metadata, dict) else None for d in documents]
        if any(e is None for e in docs_e):
            if self.embedder is None:
                raise ValueError(
                    f"method={self.method!r} needs config['embedder'] or Document.metadata[{self.embedding_key!r}]")
            docs_e = self.embedder([d.content for d in documents])
        q_e = self.config.get("query_embedding")
        if q_e is None:
            if self.embedder is None:
                raise ValueError(f"method={self.method!r} needs config['embedder'] or config['query_embedding']")
            q_e = self.embedder([query])
        q = np.asarray(q_e, dtype=np.float32).reshape(1, -1)
        x = np.asarray(docs_e, dtype=np.float32).reshape(len(documents), -1)
        if q.shape[1] != x.shape[1]:
            raise ValueError("query and document embeddings differ in dimension")
        return q, x

    def _score_embeddings(self, query: str, documents: List[Document]) -> List[Tuple[Document, float]]:
        """Whole list scored and ordered on the GPU: (score desc, input position asc)."""
        from .. import api
        import torch
        q, x = self._embeddings(query, documents)
        sign = -1.0 if self.method == "l2" else 1.0        # higher is better for every method
        n, kmax = len(documents), 2048                     # the exact search returns at most 2048 per call
        if n <= kmax:
            scores, ids = api.search_topk(q, x, k=n, metric=self.method)
            scores, ids = scores[0], ids[0]
        else:
            # longer lists: every chunk scored completely, scores scattered back to input order, one stable sort
            full = torch.empty(n, dtype=torch.float64, device=api._device())
            for lo in range(0, n, kmax):
                hi = min(n, lo + kmax)
                s, i = api.search_topk(q, x[lo:hi], k=hi - lo, metric=self.method)
                full[lo + i[0]] = sign * s[0]
            perm, srt = api.sort_scores(full[None, :], None, descending=True)
            scores, ids = sign * srt[0], perm[0]
        scores = scores.cpu().tolist()
        ids = ids.cpu().tolist()
        return [(documents[i], sign * s) for s, i in zip(scores, ids)]

    # ----------------------------------------------------------------- rerank
    def rerank(self, query: str, documents: List[Document], top_k: Optional[int] = None
               ) -> List[Tuple[Document, float]]:
        """Order ``documents`` by relevance to ``query`` (reference classical.py:231-320)."""
        started = time.time()
        if not self._validate_inputs(query, documents):
            # reference: neutral scores in the given order; tolerate non-iterables like the reference would not
            return [(doc, NEUTRAL_SCORE) for doc in documents]
        query = self._sanitize_text(query)

        if self.method in EMBEDDING_METHODS:
            ranked = self._score_embeddings(query, documents)      # already in final order; errors propagate
        else:
            if not self.model_loaded or self.model is None:
                return self._handle_reranker_failure(query, documents)
            try:
                scored = self._cross_encoder_scores(query, documents)
            except Exception as exc:
                logger.error("Cross-Encoder prediction failed: %s", exc)
                return self._handle_reranker_failure(query, documents)
            ranked = sorted(scored, key=lambda pair: pair[1], reverse=True)

        if top_k is not None and top_k > 0:                         # classical.py:307
            ranked = ranked[:top_k]
        logger.info("Reranking completed in %.2fs for %d documents", time.time() - started, len(documents))
        return ranked

    def _cross_encoder_scores(self, query: str, documents: List[Document]) -> List[Tuple[Document, float]]:
        """Cached documents first, then the uncached ones in input order (classical.py:263-295)."""
        hits: List[Tuple[Document, float]] = []
        todo_pairs: List[Tuple[str, str]] = []
        todo_docs: List[Document] = []
        for doc in documents:
            text = self._sanitize_text(doc.content)
            key = self._get_cache_key(query, text)
            if self.enable_cache and key in self.score_cache:
                hits.append((doc, self.score_cache[key]))
            else:
                todo_pairs.append((query, text))
                todo_docs.append(doc)
        fresh: List[float] = []
        for start in range(0, len(todo_pairs), self.batch_size):
            fresh.extend(self._predict_with_retries(todo_pairs[start:start + self.batch_size]))
        for doc, (_, text), score in zip(todo_docs, todo_pairs, fresh):
            if self.enable_cache:
                self.score_cache[self._get_cache_key(query, text)] = float(score)
            hits.append((doc, float(score)))
        return hits


def _default_device() -> str:
    try:
        import torch
        return "cuda" if torch.cuda.is_available() else "cpu"
    except Exception:
        return "cpu"
