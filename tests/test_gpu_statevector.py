"""K1 parity: batched statevector kernels (through the C ABI) against the NumPy oracle."""
import numpy as np
import pytest

from oracle import quantum as oq

pytestmark = pytest.mark.gpu

REL = 1e-12      # fp64 kernel vs complex128 oracle (differences: summation order, sincos ulp)


def _oracle_scores(qv, dv, doc_query, n, layers):
    return np.array([oq.quantum_similarity(qv[doc_query[j]], dv[j], n, layers=layers) for j in range(len(dv))])


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12])
@pytest.mark.parametrize("layers", [1, 3])
def test_angle_circuit_matches_oracle(cuda, n, layers):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(100 * n + layers)
    nq, per = 3, (5 if n >= 11 else 11)
    for vec_len in (2 * n, max(1, n - 2), 1):
        qv = rng.random_sample((nq, vec_len))
        dv = rng.random_sample((nq * per, vec_len))
        got = api.sv_fidelity_angle(qv, dv, docs_per_query=per, n_qubits=n, layers=layers).cpu().numpy()
        want = _oracle_scores(qv, dv, np.arange(nq * per) // per, n, layers)
        assert np.allclose(got, want, rtol=REL, atol=1e-15), (n, layers, vec_len, np.abs(got - want).max())
        if layers == 1:
            cf = np.array([oq.closed_form_fidelity(qv[j // per], dv[j], n) for j in range(nq * per)])
            assert np.allclose(got, cf, rtol=REL, atol=1e-15)


@pytest.mark.parametrize("n", [4, 9])
def test_ragged_doc_query_map_and_ties(cuda, n):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(n)
    qv = rng.random_sample((4, 2 * n))
    dv = rng.random_sample((37, 2 * n))
    dv[5] = dv[20] = dv[36]                         # duplicates
    dq = rng.randint(0, 4, size=37).astype(np.int32)
    dq[5] = dq[20] = dq[36] = 2
    got = api.sv_fidelity_angle(qv, dv, doc_query=dq, n_qubits=n).cpu().numpy()
    want = _oracle_scores(qv, dv, dq, n, 1)
    assert np.allclose(got, want, rtol=REL)
    assert got[5] == got[20] == got[36]             # bit-identical, so the stable sort ties exactly


@pytest.mark.parametrize("n", [4, 9])
def test_doc_query_entry_outside_the_queries_gives_nan_not_a_wild_read(cuda, n):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(50 + n)
    qv, dv = rng.random_sample((3, n)), rng.random_sample((9, n))
    dq = np.array([0, 1, 2, 7, -1, 2, 1, 3, 0], dtype=np.int32)           # 7, -1 and 3 name no query
    got = api.sv_fidelity_angle(qv, dv, doc_query=dq, n_qubits=n).cpu().numpy()
    bad = (dq < 0) | (dq >= 3)
    assert np.isnan(got[bad]).all()
    want = _oracle_scores(qv, dv[~bad], dq[~bad], n, 1)
    assert np.allclose(got[~bad], want, rtol=REL)


def test_known_answers(cuda, kat):
    from quantum_rag_b200 import api
    s = kat["survey"]
    for n, docs in ((4, [(t, f) for t, _, f in s["docs_n4"]]),
                    (9, [(s["docs_n4"][0][0], s["doc6825_n9"])]), (10, [(s["docs_n4"][0][0], s["doc6825_n10"])])):
        qv = oq.mock_embedding(s["query"], n)[None]
        dv = np.stack([oq.mock_embedding(t, n) for t, _ in docs])
        got = api.sv_fidelity_angle(qv, dv, docs_per_query=len(docs), n_qubits=n).cpu().numpy()
        assert np.allclose(got, [f for _, f in docs], rtol=REL)
    for c in kat["oracle"]["angle"]:
        got = api.sv_fidelity_angle(np.array([c["a"]]), np.array([c["b"]]), n_qubits=c["n"], layers=c["layers"])
        assert float(got[0]) == pytest.approx(c["f"], rel=REL)


def test_self_fidelity_unnormalised_and_zero_vectors(cuda):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(0)
    for n in (4, 6, 10):
        v = rng.random_sample((6, 2 * n)) * 7.0      # kernel renormalises like quantum.py:149-151
        got = api.sv_fidelity_angle(v, v, docs_per_query=1, n_qubits=n).cpu().numpy()
        assert np.allclose(got, 1.0, atol=1e-13)
        z = np.zeros((1, 2 * n))
        got = api.sv_fidelity_angle(z, np.vstack([z, v[:1]]), docs_per_query=2, n_qubits=n).cpu().numpy()
        want = [oq.quantum_similarity(z[0], z[0], n), oq.quantum_similarity(z[0], v[0], n)]
        assert np.allclose(got, want, rtol=REL)


def test_mock_embedding_kernel_matches_numpy_legacy_stream(cuda):
    from quantum_rag_b200 import api
    seeds = np.array([0, 1, 97, 195, 4597, 6825, 2**31, 2**32 - 1, 123456789], dtype=np.int64)
    for n in (1, 4, 9, 12):
        got = api.mock_embedding(seeds, n).cpu().numpy()
        for row, seed in zip(got, seeds):
            raw = np.random.RandomState(int(seed)).random_sample(2 * n)
            assert np.allclose(row, raw / np.linalg.norm(raw), rtol=4e-16, atol=0)
    with pytest.raises(ValueError):
        api.mock_embedding(np.array([2**32]), 4)


@pytest.mark.parametrize("shape", [(1, 1), (3, 100), (2, 129), (2, 1000), (1, 4096), (5, 7)])
def test_stable_sort_matches_python_sorted(cuda, shape):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(shape[1])
    s = rng.random_sample(shape)
    s = np.round(s, 2)                                # many ties
    s[0, 0] = np.inf if shape[1] > 1 else s[0, 0]
    for top_k in (None, 1, min(5, shape[1])):
        perm, srt = api.sort_scores(s, top_k)
        for i in range(shape[0]):
            want = oq.stable_rank(s[i].tolist(), top_k)
            assert perm[i].cpu().tolist() == want
            assert srt[i].cpu().tolist() == [s[i, j] for j in want]
    perm, _ = api.sort_scores(s, None, descending=False)
    for i in range(shape[0]):
        assert perm[i].cpu().tolist() == sorted(range(shape[1]), key=lambda j: s[i, j])


def test_unsupported_shapes_fail_loudly(cuda):
    from quantum_rag_b200 import api
    from quantum_rag_b200._lib import QragError
    with pytest.raises(QragError):
        api.sv_fidelity_angle(np.ones((1, 4)), np.ones((1, 4)), n_qubits=13)
    # lists longer than QRAG_MAX_SORT_LEN are ordered by the device library's stable sort (the reference sorts any
    # length): all-equal scores must come back in input order
    perm, srt = api.sort_scores(np.zeros((1, 5000)))
    assert perm[0].cpu().tolist() == list(range(5000)) and float(srt.abs().max()) == 0.0
