"""Import-path shim: ``from src.reranker.classical import Document`` (reference app.py:13)."""
from quantum_rag_b200.reranker.classical import ClassicalReranker, Document  # noqa: F401
