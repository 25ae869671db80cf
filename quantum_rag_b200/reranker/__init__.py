"""Host-side mirror of the reference's ``src/reranker`` package."""
from .classical import ClassicalReranker, Document
from .controller import RerankerController
from .quantum import QISKIT_AVAILABLE, QuantumReranker

__all__ = ["ClassicalReranker", "Document", "QuantumReranker", "RerankerController", "QISKIT_AVAILABLE"]
