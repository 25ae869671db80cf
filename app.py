"""Drop-in for the reference's ``app.py``: same ``POST /rerank`` route, served by the B200 reranker classes.

    uvicorn app:app --host 0.0.0.0 --port 8000

``QRAG_INDEX`` (a faiss flat index file as written by mcp/server/tools/store_in_faiss.py:99-109) and optionally
``QRAG_INDEX_METADATA`` (its label side-car) additionally enable ``POST /search_rerank``.
"""
import os

from quantum_rag_b200.service import SearchRerankService, create_app
from src.reranker.controller import RerankerController

reranker_controller = RerankerController()
_service = None
if os.environ.get("QRAG_INDEX"):
    from quantum_rag_b200.index import FlatIndex
    _service = SearchRerankService(FlatIndex.read(os.environ["QRAG_INDEX"], os.environ.get("QRAG_INDEX_METADATA")),
                                   reranker_controller)
app = create_app(controller=reranker_controller, service=_service)

if __name__ == "__main__":
    import uvicorn
    uvicorn.run("app:app", host="0.0.0.0", port=8000)
