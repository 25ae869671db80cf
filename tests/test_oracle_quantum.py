"""The oracle against the recorded known answers and against its own closed form (CPU)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import quantum as oq

REL = 2e-15 * 50     # two independent fp64 evaluations of ~100 flops each


def test_legacy_rng_stream(kat):
    rs = np.random.RandomState(0)
    assert np.allclose(rs.random_sample(3), kat["survey"]["legacy_rng_seed0"], atol=5e-9)


def test_query_embedding_and_state(kat):
    s = kat["survey"]
    assert oq.char_sum(s["query"]) == s["query_charsum"]
    emb = oq.mock_embedding(s["query"], 4)
    assert emb.shape == (8,)
    assert np.array_equal(emb[:4], np.array(s["query_embedding_head"]))        # bit-exact legacy stream
    psi = oq.circuit_statevector(emb, 4)
    assert abs(psi[0] - complex(*s["psi0"])) < 1e-15
    assert abs(psi[15] - complex(*s["psi15"])) < 1e-15
    assert abs(np.vdot(psi, psi) - 1) < 1e-14


def test_survey_fidelities(kat):
    s = kat["survey"]
    q = oq.mock_embedding(s["query"], 4)
    for text, csum, want in s["docs_n4"]:
        assert oq.char_sum(text) == csum
        got = oq.quantum_similarity(q, oq.mock_embedding(text, 4), 4)
        assert got == pytest.approx(want, rel=1e-14)
        assert oq.closed_form_fidelity(q, oq.mock_embedding(text, 4), 4) == pytest.approx(want, rel=1e-14)
    doc = s["docs_n4"][0][0]
    for n, key in ((9, "doc6825_n9"), (10, "doc6825_n10")):
        got = oq.quantum_similarity(oq.mock_embedding(s["query"], n), oq.mock_embedding(doc, n), n)
        assert got == pytest.approx(s[key], rel=1e-14)


def test_anagrams_tie_exactly_and_keep_input_order():
    q = "sponsor"
    ranked = oq.quantum_rerank_strings(q, ["zz", "ab", "ba", "ab"], None, 4)
    scores = {i: s for i, s in ranked}
    assert scores[1] == scores[2] == scores[3]
    tied = [i for i, _ in ranked if i in (1, 2, 3)]
    assert tied == [1, 2, 3]


def test_other_method_is_constant_half():
    assert oq.quantum_similarity([1, 2], [3, 4], 4, method="swap_test") == 0.5
    ranked = oq.quantum_rerank_strings("q", ["a", "b", "c"], 2, 4, method="other")
    assert ranked == [(0, 0.5), (1, 0.5)]


def test_top_k_slice_semantics():
    texts = ["a", "b", "c", "d"]
    full = oq.quantum_rerank_strings("q", texts, None)
    assert oq.quantum_rerank_strings("q", texts, 0) == []
    assert oq.quantum_rerank_strings("q", texts, -1) == full[:-1]
    assert oq.quantum_rerank_strings("q", texts, 10) == full
    assert oq.quantum_rerank_strings("q", [], 3) == []


def test_regression_lock(kat):
    for c in kat["oracle"]["angle"]:
        assert oq.quantum_similarity(c["a"], c["b"], c["n"], layers=c["layers"]) == pytest.approx(c["f"], rel=1e-13)
    for c in kat["oracle"]["amplitude"]:
        q, d = np.float32(c["q"]), np.float32(c["d"])
        assert oq.amplitude_fidelity(q, d) == pytest.approx(c["f"], rel=1e-13)
        f_state = oq.state_fidelity(oq.amplitude_state(q, c["n"]), oq.amplitude_state(d, c["n"]))
        assert f_state == pytest.approx(c["f"], rel=1e-12)
    for c in kat["oracle"]["feature_map"]:
        got = oq.feature_map_fidelity(np.float32(c["q"]), np.float32(c["d"]), c["n"], c["layers"])
        assert got == pytest.approx(c["f"], rel=1e-13)


vec = st.lists(st.floats(min_value=0.0, max_value=1.0, allow_nan=False), min_size=1, max_size=12)


@settings(max_examples=60, deadline=None)
@given(a=vec, b=vec, n=st.integers(1, 6), layers=st.integers(1, 3))
def test_fidelity_properties(a, b, n, layers):
    m = min(len(a), len(b))
    a, b = a[:m], b[:m]
    f_ab = oq.quantum_similarity(a, b, n, layers=layers)
    f_ba = oq.quantum_similarity(b, a, n, layers=layers)
    assert -1e-12 <= f_ab <= 1 + 1e-12
    assert f_ab == pytest.approx(f_ba, abs=1e-13)
    assert oq.quantum_similarity(a, a, n, layers=layers) == pytest.approx(1.0, abs=1e-12)
    if layers == 1:
        assert f_ab == pytest.approx(oq.closed_form_fidelity(a, b, n), abs=1e-13)


@settings(max_examples=30, deadline=None)
@given(a=vec, n=st.integers(2, 6))
def test_trailing_cx_chain_is_a_permutation(a, n):
    # the CX chain maps basis x to prefix-xor(x): check on the simulated state
    v = np.asarray(a)
    nv = np.linalg.norm(v)
    vn = v / nv if nv > 0 else v
    state = np.zeros(1 << n, dtype=np.complex128)
    state[0] = 1
    for i in range(min(len(vn), n)):
        state = oq._apply_1q(state, oq.ry(vn[i] * np.pi), i)
        state = oq._apply_1q(state, oq.rz(vn[i] * np.pi / 2), i)
    full = oq.circuit_statevector(a, n)
    for x in range(1 << n):
        y, acc = 0, 0
        for k in range(n):
            acc ^= (x >> k) & 1
            y |= acc << k
        assert abs(full[y] - state[x]) < 1e-15


def test_amplitude_batch_matches_scalar():
    rng = np.random.RandomState(1)
    Q = rng.standard_normal((3, 20)).astype(np.float32)
    C = rng.standard_normal((3, 7, 20)).astype(np.float32)
    C[1, 2] = 0
    out = oq.amplitude_fidelity_batch(Q, C)
    for i in range(3):
        for j in range(7):
            assert out[i, j] == pytest.approx(oq.amplitude_fidelity(Q[i], C[i, j]), rel=1e-13, abs=1e-300)
    assert out[1, 2] == 0.0
    order = oq.rank_rows(out)
    for i in range(3):
        assert order[i].tolist() == oq.stable_rank(out[i].tolist())
