"""Drop-in API surface and host-side logic, runnable without a GPU."""
import random
import string

import numpy as np
import pytest

from oracle import reranker as orr
from src.reranker.classical import ClassicalReranker, Document
from src.reranker.controller import RerankerController
from src.reranker.quantum import QISKIT_AVAILABLE, QuantumReranker
from quantum_rag_b200.reranker.quantum import _char_sum


def test_import_paths_match_reference_app():
    # reference app.py:12-13
    import src.reranker.classical as c
    import src.reranker.controller as k
    import src.reranker.quantum as q
    assert k.RerankerController is RerankerController and c.Document is Document
    assert q.QuantumReranker is QuantumReranker and isinstance(QISKIT_AVAILABLE, bool)


def test_document_defaults():
    d = Document("1", "text")
    assert (d.id, d.content, d.source, d.metadata) == ("1", "text", None, {})
    d2 = Document("2", "t", "src", {"a": 1})
    assert d2.source == "src" and d2.metadata == {"a": 1}
    assert Document("3", "t", metadata=None).metadata == {}


def test_controller_defaults_and_keywords():
    ctl = RerankerController()
    assert ctl.complexity_threshold == 8
    assert ctl.quantum_keywords == list(orr.QUANTUM_KEYWORDS)
    assert isinstance(ctl.classical_reranker, ClassicalReranker)
    assert isinstance(ctl.quantum_reranker, QuantumReranker)
    assert ctl.quantum_reranker.n_qubits == 4 and ctl.quantum_reranker.method == "state_fidelity"
    assert isinstance(ctl.quantum_reranker.classical_fallback, ClassicalReranker)
    assert RerankerController({"complexity_threshold": 2}).complexity_threshold == 2


@pytest.mark.parametrize("query,want", [
    ("had a good time", "quantum"),            # "ad" is a substring of "had"
    ("what is the weather", "classical"),
    ("one two three four five six seven eight", "classical"),
    ("one two three four five six seven eight nine", "quantum"),
    ("SPONSOR segment", "quantum"),
    ("", "classical"),
    ("ideal", "quantum"),                       # contains "deal"
])
def test_select_reranker_examples(query, want):
    assert RerankerController().select_reranker(query) == want
    assert orr.select_reranker(query) == want


def test_select_reranker_random_against_oracle():
    rnd = random.Random(5)
    vocab = ["ad", "bad", "offer", "coffee", "the", "a", "brand", "xyz", "Deal", "pro", "motion", "q"]
    ctl = RerankerController({"complexity_threshold": 5})
    for _ in range(300):
        words = [rnd.choice(vocab) if rnd.random() < 0.5 else
                 "".join(rnd.choice(string.ascii_letters) for _ in range(rnd.randint(1, 6)))
                 for _ in range(rnd.randint(0, 9))]
        q = rnd.choice([" ", "  ", "\t"]).join(words)
        assert ctl.select_reranker(q) == orr.select_reranker(q, 5)
    assert ctl.select_rerankers(["ad", "x"]) == ["quantum", "classical"]


def test_classical_validation_returns_neutral_scores_in_order():
    r = ClassicalReranker()
    docs = [Document("a", "x"), Document("b", "y")]
    assert r.rerank("", docs) == [(docs[0], 0.5), (docs[1], 0.5)]
    assert r.rerank("   ", docs, top_k=1) == [(docs[0], 0.5), (docs[1], 0.5)]     # no slicing on this route
    assert r.rerank("q", []) == []
    bad = [Document("a", "x"), Document("b", "")]
    assert r.rerank("q", bad) == [(bad[0], 0.5), (bad[1], 0.5)]
    mixed = [Document("a", "x"), "not a document"]
    assert r.rerank("q", mixed) == [(mixed[0], 0.5), (mixed[1], 0.5)]
    assert not r._validate_inputs(123, docs)
    assert orr.classical_inputs_valid("q", ["x", ""], [True, True]) is False


def test_classical_cross_encoder_unavailable_is_the_reference_failure_route():
    r = ClassicalReranker()
    assert r.method == "cross-encoder" and r.batch_size == 32 and r.max_sequence_length == 512
    if r.model_loaded:
        pytest.skip("sentence_transformers present")
    docs = [Document(str(i), f"doc {i}") for i in range(4)]
    out = r.rerank("query", docs, top_k=2)
    assert out == [(d, 0.5) for d in docs]           # classical.py:258-260: not sliced, original order


def test_sanitize_and_cache_key():
    r = ClassicalReranker({"max_sequence_length": 2})
    assert r._sanitize_text("  a \n\t b  ") == "a b"
    assert r._sanitize_text("x" * 100) == "x" * 8
    assert r._sanitize_text(12) == "12"
    assert r._get_cache_key("q", "d") == f"{hash('q')}_{hash('d')}"


def test_embedding_method_requires_embeddings():
    r = ClassicalReranker({"method": "cosine"})
    with pytest.raises(ValueError):
        r.rerank("q", [Document("a", "x")])


def test_quantum_empty_and_config():
    q = QuantumReranker()
    assert q.rerank("anything", []) == []
    assert q.rerank("anything", None) == []
    q9 = QuantumReranker({"n_qubits": 9, "method": "other"})
    assert q9.n_qubits == 9 and q9.method == "other"
    assert QuantumReranker({"encoding": "amplitude"}).layers == 0


def test_mock_embedding_matches_numpy_global_stream():
    q = QuantumReranker()
    for text in ["hello", "", "naïve café", "ab", "ba", "\U0001F600 emoji"]:
        seed = sum(ord(c) for c in text)
        assert _char_sum(text) == seed
        np.random.seed(seed)                       # what the reference does (quantum.py:183)
        v = np.random.random(8)
        want = v / np.linalg.norm(v)
        assert np.array_equal(q._mock_embedding(text), want)
    state = np.random.get_state()[1].copy()
    q._mock_embedding("does not touch the global generator")
    assert np.array_equal(np.random.get_state()[1], state)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        QuantumReranker().rerank("query", [Document("a", "x")])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        RerankerController().rerank("sponsor ad", [Document("a", "x")], reranker_type="quantum")


def test_ixf_parse_dump_roundtrip_and_fixture_header(piers):
    from oracle import search as osr
    from quantum_rag_b200 import index as qidx
    x = piers["vectors"]
    buf = qidx.dump_ixf(x, qidx.FAISS_METRIC_L2)
    assert buf == osr.write_ixf(x, 1)                      # product writer == oracle writer
    assert len(buf) == 731181                              # the size of the reference's fixture file
    y, metric = qidx.parse_ixf(buf)
    assert metric == 1 and np.array_equal(y, x)
    y2, metric2 = qidx.parse_ixf(qidx.dump_ixf(x[:3], qidx.FAISS_METRIC_IP))
    assert metric2 == 0 and np.array_equal(y2, x[:3])
    with pytest.raises(ValueError):
        qidx.parse_ixf(b"IxF2" + buf[4:40])
    with pytest.raises(ValueError):
        qidx.parse_ixf(b"IVFx" + buf[4:])


def test_metadata_sidecar_only_plain_lists(tmp_path):
    import pickle
    from quantum_rag_b200 import index as qidx
    p = tmp_path / "meta.pkl"
    p.write_bytes(pickle.dumps(["show/a", "show/b"]))
    assert qidx.load_metadata(str(p)) == ["show/a", "show/b"]
    p.write_bytes(pickle.dumps(np.arange(3)))              # needs a global -> refused, never executed
    with pytest.raises(pickle.UnpicklingError):
        qidx.load_metadata(str(p))


def test_exchange_len_matches_the_library(libqrag):
    """sharded.exchange_len (host logic) == qrag_search_tc_exchange_len (what the kernels size their lists with)."""
    import ctypes
    from quantum_rag_b200.sharded import exchange_len
    for k in (1, 7, 10, 100, 999, 1000, 2048):
        for world in (1, 2, 3, 4, 8, 16, 64):
            n = ctypes.c_int(0)
            assert libqrag.qrag_search_tc_exchange_len(k, world, ctypes.byref(n)) == 0
            assert n.value == exchange_len(k, world), (k, world)
            assert 1 <= n.value <= k and (world == 1) <= (n.value == k)
            assert world * n.value >= k                      # the union of the cut lists still holds k entries


def test_json_sidecar_roundtrip_and_pickle_conversion(tmp_path):
    """(f)4: the JSON id -> label side-car replaces the pickle of store_in_faiss.py:111-122."""
    import json
    import pickle
    from quantum_rag_b200 import index as qidx
    labels = ["Piers_Morgan_Uncensored/02c7ef14", "show/ünïcode ✓", 'quote " and \\ backslash', "", "show/a"]
    jp = tmp_path / "idx_labels.json"
    qidx.dump_labels(labels, str(jp))
    assert qidx.load_labels(str(jp)) == labels
    doc = json.loads(jp.read_text(encoding="utf-8"))
    assert doc["format"] == "qrag-labels" and doc["version"] == 1 and doc["count"] == len(labels)
    # the reference's pickle (a plain list of str) converts in one call, through the restricted unpickler
    pk = tmp_path / "piers_morgan_faiss_index_metadata.pkl"
    pk.write_bytes(pickle.dumps(labels))
    out = qidx.convert_metadata_pickle(str(pk))
    assert out == str(tmp_path / "piers_morgan_faiss_index_labels.json") and qidx.load_labels(out) == labels
    assert qidx.sidecar_paths(str(tmp_path / "piers_morgan_faiss_index.faiss")) == (out, str(pk))
    # malformed side-cars are refused: wrong format tag, unknown version, count mismatch
    for bad in ({"format": "other", "version": 1, "count": 0, "labels": []},
                {"format": "qrag-labels", "version": 2, "count": 0, "labels": []},
                {"format": "qrag-labels", "version": 1, "count": 3, "labels": ["a"]},
                ["a", "b"]):
        jp.write_text(json.dumps(bad))
        with pytest.raises(ValueError):
            qidx.load_labels(str(jp))
    # a pickle that needs a global is still refused by the converter (never executed)
    pk.write_bytes(pickle.dumps(np.arange(3)))
    with pytest.raises(pickle.UnpicklingError):
        qidx.convert_metadata_pickle(str(pk))


def test_numa_binding_helpers(tmp_path):
    """hostmem: cpulist parsing and the sysfs lookups behind bind_to_gpu_numa_node (a fake sysfs tree)."""
    from quantum_rag_b200 import hostmem
    assert hostmem.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert hostmem.parse_cpulist("") == [] and hostmem.parse_cpulist("5") == [5]
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    (dev / "numa_node").write_text("1\n")
    node = tmp_path / "devices/system/node/node1"
    node.mkdir(parents=True)
    (node / "cpulist").write_text("56-111,168-223\n")
    assert hostmem.gpu_numa_node("00000000:1B:00.0", sysfs=str(tmp_path)) == 1      # CUDA's 8-digit domain, upper case
    assert hostmem.gpu_numa_node("0000:1b:00.0", sysfs=str(tmp_path)) == 1
    assert hostmem.gpu_numa_node("0000:ff:00.0", sysfs=str(tmp_path)) is None
    (dev / "numa_node").write_text("-1\n")
    assert hostmem.gpu_numa_node("0000:1b:00.0", sysfs=str(tmp_path)) is None       # platform does not say
    cpus = hostmem.node_cpus(1, sysfs=str(tmp_path))
    assert len(cpus) == 112 and cpus[0] == 56 and cpus[-1] == 223
    assert hostmem.node_cpus(7, sysfs=str(tmp_path)) == []


def test_embedding_memo_is_bounded_and_qubit_range_is_checked_up_front():
    """the text-hash memo of a long-running service must not grow without bound; n_qubits beyond what the
    kernels stage (12) is refused at construction, not at the first rerank."""
    from src.reranker.quantum import QuantumReranker
    q = QuantumReranker({"embedding_memo_size": 8})
    first = q._mock_embedding("a").copy()
    for i in range(100):
        q._mock_embedding("x" * (i + 2))
    assert len(q._embedding_memo) <= 8
    assert np.array_equal(q._mock_embedding("a"), first)             # evicted and recomputed: same bits (seeded stream)
    with pytest.raises(ValueError, match="n_qubits"):
        QuantumReranker({"n_qubits": 13})
    assert QuantumReranker({"n_qubits": 12}).n_qubits == 12


def test_char_sum_fast_paths_agree_with_the_definition():
    """sum(ord(c)) (quantum.py:182): the ASCII byte-sum path and the UTF-32 path against the definition, on random
    strings that mix ASCII, Latin-1, BMP and astral code points (and the lone-surrogate case Python strings allow)."""
    rng = np.random.RandomState(5)
    pools = [(32, 127), (128, 256), (0x400, 0x500), (0x4E00, 0x4F00), (0x1F600, 0x1F650)]
    for trial in range(300):
        n = int(rng.randint(0, 60))
        chars = []
        for _ in range(n):
            lo, hi = pools[int(rng.randint(0, len(pools) if trial % 3 else 1))]
            chars.append(chr(int(rng.randint(lo, hi))))
        text = "".join(chars)
        assert _char_sum(text) == sum(ord(c) for c in text), repr(text)
    assert _char_sum("\ud800x") == 0xD800 + ord("x")
