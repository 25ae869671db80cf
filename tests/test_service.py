"""Routes of quantum_rag_b200.service: the reference's POST /rerank contract (app.py:56-77) and /search_rerank.

The CPU tests drive the HTTP layer with stub compute (no GPU, no oracle in the product path); the GPU test runs the
real service over the reference's own index fixture and checks it against the oracle.
"""
import numpy as np
import pytest

from quantum_rag_b200.reranker import Document
from quantum_rag_b200.service import SearchRerankService, create_app


class _StubController:
    """Same surface as RerankerController; scores by document length (no CUDA)."""

    def __init__(self):
        self.calls = []

    def select_rerankers(self, queries):
        return ["quantum" if len(q.split()) > 8 else "classical" for q in queries]

    def rerank(self, query, documents, top_k=None, reranker_type="auto"):
        self.calls.append((query, [d.id for d in documents], top_k, reranker_type))
        if query == "boom":
            raise RuntimeError("QRAG_ERR_CUDA: no CUDA device available (libqrag has no CPU fallback)")
        ranked = sorted(((d, float(len(d.content))) for d in documents), key=lambda p: p[1], reverse=True)
        return {"documents": ranked[:top_k], "reranker_used": "classical", "query": query}


def _client(app):
    from starlette.testclient import TestClient
    return TestClient(app)


def test_rerank_route_matches_reference_contract():
    ctl = _StubController()
    client = _client(create_app(controller=ctl))
    body = {"query": "find the ad", "documents": [{"id": "a", "content": "xx"}, {"id": "b", "content": "xxxx", "source": "s"}]}
    r = client.post("/rerank", json=body)
    assert r.status_code == 200
    out = r.json()
    # app.py returns the controller's dict: tuples of (Document, score) serialise as [object, float]
    assert out["reranker_used"] == "classical" and out["query"] == "find the ad"
    assert [d[0]["id"] for d in out["documents"]] == ["b", "a"]
    assert out["documents"][0][0] == {"id": "b", "content": "xxxx", "source": "s", "metadata": {}}
    assert out["documents"][0][1] == 4.0
    assert ctl.calls[-1] == ("find the ad", ["a", "b"], 5, "auto")          # defaults of app.py:32-33
    # errors are reported in the body with status 200 (app.py:75-77)
    r = client.post("/rerank", json={"query": "boom", "documents": []})
    assert r.status_code == 200 and "no CPU fallback" in r.json()["error"]
    # malformed requests are rejected by the schema like in the reference
    assert client.post("/rerank", json={"documents": []}).status_code == 422
    assert "search_rerank" in client.get("/").json()["endpoints"]


def test_search_rerank_route_without_index_reports_error():
    client = _client(create_app(controller=_StubController()))
    r = client.post("/search_rerank", json={"embeddings": [[0.0, 1.0]]})
    assert r.status_code == 200 and "no index loaded" in r.json()["error"]


def test_service_choices_follow_the_controller_rules():
    svc = SearchRerankService(index=None, controller=_StubController())
    assert svc._choices(2, "quantum", None) == ["quantum", "quantum"]
    assert svc._choices(2, "classical", None) == ["classical", "classical"]
    assert svc._choices(1, "anything-else", None) == ["classical"]          # controller.py:97-99
    long_q = "one two three four five six seven eight nine"
    assert svc._choices(2, "auto", [long_q, "short"]) == ["quantum", "classical"]
    with pytest.raises(ValueError, match="query texts"):
        svc._choices(1, "auto", None)


@pytest.mark.gpu
def test_search_rerank_over_reference_fixture(cuda, piers):
    from oracle import quantum as oq
    from quantum_rag_b200.index import FlatIndex
    from quantum_rag_b200.reranker import RerankerController
    x, labels = piers["vectors"], [str(s) for s in piers["labels"]]
    svc = SearchRerankService(FlatIndex(x, labels=labels), RerankerController())
    queries = ["which segments contain a sponsor advertisement", "hello"]
    emb = x[[0, 5]]
    res = svc.search_rerank(emb, k1=20, k2=5, reranker_type="auto", queries=queries)
    assert [r["reranker_used"] for r in res] == ["quantum", "classical"]
    # query 0: search = the golden top-20 of row 0, rerank = oracle amplitude fidelity, stable
    top20 = piers["top20_ids"][0]
    f = oq.amplitude_fidelity_batch(emb[:1], x[top20][None])[0]
    order = oq.rank_rows(f[None], 5)[0]
    got = res[0]["documents"]
    assert [d["id"] for d in got] == top20[order].tolist()
    assert [d["search_rank"] for d in got] == order.tolist()
    assert np.allclose([d["score"] for d in got], f[order], rtol=1e-12)
    assert got[0]["label"] == labels[got[0]["id"]]
    # query 1: classical = search order, score = -squared L2 (higher is better)
    got = res[1]["documents"]
    assert [d["search_rank"] for d in got] == [0, 1, 2, 3, 4] and got[0]["id"] == 5
    assert all(got[i]["score"] >= got[i + 1]["score"] for i in range(4)) and got[0]["score"] == pytest.approx(0.0, abs=1e-12)
    # through HTTP, k1 larger than the index, top_k = None keeps the whole list
    client = _client(create_app(controller=svc.controller, service=svc))
    r = client.post("/search_rerank", json={"embeddings": emb[:1].tolist(), "k1": 500, "top_k": None})
    docs = r.json()["results"][0]["documents"]
    assert len(docs) == 119 and docs[0]["id"] == 0 and docs[0]["score"] == pytest.approx(1.0, abs=1e-12)
    assert all(docs[i]["score"] >= docs[i + 1]["score"] for i in range(118))
