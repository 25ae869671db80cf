"""quantum_rag_b200 -- B200-native reranking hot path behind the quantum-rag reranker API.

Layout
    csrc/        hand-written sm_100a CUDA kernels and the C ABI (include/qrag.h)
    _lib.py      in-tree build + ctypes binding of libqrag.so
    api.py       tensor-level API (torch tensors in / out, no torch in the ABI)
    reranker/    drop-in for the reference's src/reranker classes
    index.py     faiss IndexFlat (IxF2 / IxFI) reader, writer and searchable wrapper
    sharded.py   row-sharded search + rerank over torch.distributed
    service.py   search -> select -> rerank service and the HTTP routes (/rerank, /search_rerank)
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
