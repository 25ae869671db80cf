"""Drop-in reranker classes on the GPU vs the restated reference (string API, config 1)."""
import random
import string

import numpy as np
import pytest

from oracle import quantum as oq
from oracle import reranker as orr
from oracle import search as osr
from src.reranker.classical import ClassicalReranker, Document
from src.reranker.controller import RerankerController
from src.reranker.quantum import QuantumReranker

pytestmark = pytest.mark.gpu


def _texts(rnd, n):
    out = []
    for _ in range(n):
        out.append("".join(rnd.choice(string.ascii_letters + "  .,") for _ in range(rnd.randint(0, 60))))
    return out


@pytest.mark.parametrize("n_qubits", [4, 5, 9])
@pytest.mark.parametrize("backend", ["host", "device"])
def test_quantum_reranker_matches_reference_semantics(cuda, n_qubits, backend):
    rnd = random.Random(n_qubits)
    rr = QuantumReranker({"n_qubits": n_qubits, "embedding_backend": backend})
    for trial in range(4):
        texts = _texts(rnd, rnd.randint(1, 40)) + ["ab", "ba", "ab"]           # anagrams collide (same char sum)
        rnd.shuffle(texts)
        docs = [Document(str(i), t) for i, t in enumerate(texts)]
        query = "".join(rnd.choice(string.ascii_lowercase + " ") for _ in range(30))
        for top_k in (None, 0, 1, 5, len(docs) + 3, -2):
            got = rr.rerank(query, docs, top_k)
            want = oq.quantum_rerank_strings(query, texts, top_k, n_qubits)
            assert [docs.index(d) for d, _ in got] == [i for i, _ in want], (trial, top_k)
            assert all(d is docs[i] for (d, _), (i, _) in zip(got, want))      # identity preserved
            assert np.allclose([s for _, s in got], [s for _, s in want], rtol=1e-12, atol=0)
            assert all(isinstance(s, float) for _, s in got)


def test_known_answers_through_the_class(cuda, kat):
    s = kat["survey"]
    rr = QuantumReranker()
    docs = [Document(str(i), t) for i, (t, _, _) in enumerate(s["docs_n4"])]
    scored = rr._quantum_score_documents(s["query"], docs)
    assert np.allclose([v for _, v in scored], [f for _, _, f in s["docs_n4"]], rtol=1e-12)
    ranked = rr.rerank(s["query"], docs)
    ab = [d.content for d, _ in ranked if d.content in ("ab", "ba")]
    assert ab == ["ab", "ba"]                          # exact tie keeps input order


def test_lists_longer_than_the_kernel_sort(cuda):
    """5 000 documents (> QRAG_MAX_SORT_LEN): the reference sorts any length, so must the drop-in; ties stay stable."""
    rnd = random.Random(5)
    texts = _texts(rnd, 4997) + ["ab", "ba", "ab"]
    rnd.shuffle(texts)
    docs = [Document(str(i), t) for i, t in enumerate(texts)]
    rr = QuantumReranker()
    for top_k in (None, 7):
        got = rr.rerank("which ad is this", docs, top_k)
        want = oq.quantum_rerank_strings("which ad is this", texts, top_k, 4)
        assert [int(d.id) for d, _ in got] == [i for i, _ in want]
        assert np.allclose([s for _, s in got], [s for _, s in want], rtol=1e-12, atol=0)


def test_other_method_gives_half_in_input_order(cuda):
    rr = QuantumReranker({"method": "swap_test"})
    docs = [Document(str(i), t) for i, t in enumerate(["x", "y", "z"])]
    assert rr.rerank("q", docs, 2) == [(docs[0], 0.5), (docs[1], 0.5)]


def test_controller_dispatch(cuda):
    ctl = RerankerController()
    docs = [Document(str(i), t) for i, t in enumerate(["alpha", "beta", "gamma delta"])]
    out = ctl.rerank("find the sponsor read", docs, top_k=2)
    assert out["reranker_used"] == "quantum" and out["query"] == "find the sponsor read"
    want = oq.quantum_rerank_strings("find the sponsor read", [d.content for d in docs], 2)
    assert [docs.index(d) for d, _ in out["documents"]] == [i for i, _ in want]
    assert ctl.rerank("hello", docs)["reranker_used"] == orr.dispatch("auto", "hello") == "classical"
    assert ctl.rerank("hello", docs, reranker_type="quantum")["reranker_used"] == "quantum"
    assert ctl.rerank("sponsor", docs, reranker_type="weird")["reranker_used"] == "classical"


def test_config1_fixture_search_then_quantum_rerank(cuda, piers):
    """FAISS top-20 over the reference's index, then quantum rerank of those 20 (both encodings)."""
    from quantum_rag_b200 import api
    x, labels = piers["vectors"], [str(s) for s in piers["labels"]]
    q = x[:1]
    _, ids = api.search_topk(q, x, 20, "l2")
    ids = ids[0].cpu().tolist()
    assert ids == piers["top20_ids"][0].tolist()
    # (i) reference circuit on Document(content=label), n_qubits = 4
    docs = [Document(str(i), labels[i]) for i in ids]
    got = QuantumReranker().rerank(labels[0], docs, top_k=10)
    want = oq.quantum_rerank_strings(labels[0], [labels[i] for i in ids], 10)
    assert [docs.index(d) for d, _ in got] == [i for i, _ in want]
    # (ii) amplitude encoding of the stored 1536-d embeddings, 11 qubits
    docs = [Document(str(i), labels[i], metadata={"embedding": x[i]}) for i in ids]
    rr = QuantumReranker({"encoding": "amplitude", "n_qubits": 11, "query_embedding": q[0]})
    got = rr.rerank(labels[0], docs)
    f = oq.amplitude_fidelity_batch(q, x[ids][None])[0]
    assert [docs.index(d) for d, _ in got] == oq.stable_rank(f.tolist())
    assert np.allclose([s for _, s in got], sorted(f.tolist(), reverse=True), rtol=1e-12)


@pytest.mark.parametrize("method,mid", [("cosine", osr.METRIC_COSINE), ("ip", osr.METRIC_IP), ("l2", osr.METRIC_L2)])
def test_classical_embedding_methods(cuda, method, mid):
    rng = np.random.RandomState(3)
    emb = rng.standard_normal((30, 64)).astype(np.float32)
    emb[7] = emb[2]
    table = {f"doc {i}": emb[i] for i in range(30)}
    qv = rng.standard_normal(64).astype(np.float32)
    table["the query"] = qv
    rr = ClassicalReranker({"method": method, "embedder": lambda texts: np.stack([table[t] for t in texts])})
    docs = [Document(str(i), f"doc {i}") for i in range(30)]
    got = rr.rerank("the query", docs, top_k=12)
    s, i = osr.exact_search(qv[None], emb, 12, mid)
    assert [docs.index(d) for d, _ in got] == i[0].tolist()
    sign = -1.0 if method == "l2" else 1.0
    assert np.allclose([v for _, v in got], sign * s[0], rtol=1e-12, atol=1e-13)
    assert len(rr.rerank("the query", docs, top_k=0)) == 30                      # classical.py:307: only if > 0
    # embeddings carried by the documents instead of an embedder
    docs_m = [Document(str(i), f"doc {i}", metadata={"embedding": emb[i]}) for i in range(30)]
    rr2 = ClassicalReranker({"method": method, "query_embedding": qv})
    assert [docs_m.index(d) for d, _ in rr2.rerank("the query", docs_m, 12)] == i[0].tolist()


@pytest.mark.parametrize("method,mid", [("cosine", osr.METRIC_COSINE), ("l2", osr.METRIC_L2)])
def test_classical_embedding_methods_long_list(cuda, method, mid):
    """More documents than one exact-search call returns (2048): chunked scoring, one stable sort."""
    rng = np.random.RandomState(8)
    n = 5000
    emb = rng.standard_normal((n, 32)).astype(np.float32)
    emb[4100] = emb[17]                                                          # tie across chunks
    qv = rng.standard_normal(32).astype(np.float32)
    docs = [Document(str(i), f"d{i}", metadata={"embedding": emb[i]}) for i in range(n)]
    rr = ClassicalReranker({"method": method, "query_embedding": qv})
    got = rr.rerank("q", docs, top_k=None)
    assert len(got) == n
    order = [int(d.id) for d, _ in got]
    sc = np.array([v for _, v in got])
    assert np.all(sc[:-1] >= sc[1:]) and sorted(order) == list(range(n))
    assert order.index(17) + 1 == order.index(4100)                              # equal scores keep input order
    s, i = osr.exact_search(qv[None], emb, 100, mid)
    sign = -1.0 if method == "l2" else 1.0
    assert order[:100] == i[0].tolist() and np.allclose(sc[:100], sign * s[0], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("dim,layers", [(384, 0), (1024, 3), (1000, 1)])
def test_quantum_reranker_amplitude_topk_uses_fused_kernels_same_answer(cuda, dim, layers):
    """``QuantumReranker.rerank(..., top_k=k)`` with amplitude encoding hands top-k to the fused kernels (streaming rerank;
    at 10 qubits with layers: complex64 filter + complex128 certification).  Same Document objects, order and score bits as
    ranking everything and slicing (quantum.py:70-76), exact ties in input order."""
    rng = np.random.RandomState(dim + layers)
    emb = rng.standard_normal((60, dim)).astype(np.float32)
    emb[31] = emb[4]
    emb[50] = emb[4]
    qv = rng.standard_normal(dim).astype(np.float32)
    docs = [Document(str(i), f"doc {i}", metadata={"embedding": emb[i]}) for i in range(60)]
    rr = QuantumReranker({"encoding": "amplitude", "layers": layers, "query_embedding": qv})
    full = rr.rerank("q", docs)                                   # top_k None: everything, the generic path
    assert len(full) == 60
    for k in (1, 7, 59):
        got = rr.rerank("q", docs, top_k=k)
        assert [d is w for (d, _), (w, _) in zip(got, full[:k])] == [True] * k
        assert [s for _, s in got] == [s for _, s in full[:k]]
    assert rr.rerank("q", docs, top_k=0) == [] and len(rr.rerank("q", docs, top_k=60)) == 60
    assert rr.rerank("q", docs, top_k=-3) == full[:-3]            # plain Python slice semantics
