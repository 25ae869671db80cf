"""Import-path shim: ``from src.reranker.controller import RerankerController`` (reference app.py:12)."""
from quantum_rag_b200.reranker.controller import RerankerController  # noqa: F401
from quantum_rag_b200.reranker.classical import ClassicalReranker, Document  # noqa: F401
from quantum_rag_b200.reranker.quantum import QuantumReranker  # noqa: F401
