"""Import-path shim for the reference's ``src.reranker.quantum`` module."""
from quantum_rag_b200.reranker.quantum import QISKIT_AVAILABLE, QuantumReranker  # noqa: F401
from quantum_rag_b200.reranker.classical import ClassicalReranker, Document  # noqa: F401
