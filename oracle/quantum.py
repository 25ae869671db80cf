"""NumPy (complex128) restatement of the reference's quantum reranker arithmetic.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Parity unpinned by reference
tests; every function cites the reference lines it restates
(``/root/reference/src/reranker/quantum.py`` unless noted).

Third-party semantics restated here because Qiskit (pinned 2.1.1 in
``poetry.lock:2122-2123``; ``Aer``/``execute`` not importable from it) and
qiskit-aer (not pinned at all) are absent from this image:

* qubits start in |0>, little-endian: qubit ``i`` is bit ``i`` of the basis index;
* ``RY(t) = [[cos t/2, -sin t/2], [sin t/2, cos t/2]]``;
* ``RZ(p) = diag(exp(-i p/2), exp(+i p/2))``;
* ``CX(c, t)`` flips bit ``t`` of every basis index whose bit ``c`` is 1;
* ``state_fidelity`` of two pure states is ``|<psi2|psi1>|^2``.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

PI = np.pi


# --------------------------------------------------------------------------
# quantum.py:169-185  _mock_embedding
# --------------------------------------------------------------------------
def char_sum(text: str) -> int:
    """``sum(ord(c) for c in text)`` -- quantum.py:182."""
    return sum(ord(c) for c in text)


def mock_embedding(text: str, n_qubits: int = 4) -> np.ndarray:
    """Deterministic text-hash embedding, quantum.py:169-185.

    The reference reseeds NumPy's *global* legacy MT19937 (``np.random.seed``,
    quantum.py:183) and draws ``2*n_qubits`` doubles.  A private ``RandomState``
    seeded with the same integer yields the identical stream (NumPy freezes the
    legacy generator), without the global side effect.
    """
    rs = np.random.RandomState(char_sum(text))
    v = rs.random_sample(n_qubits * 2)
    return v / np.linalg.norm(v)


# --------------------------------------------------------------------------
# quantum.py:138-167  _vector_to_circuit  (+ Aer statevector simulation, :125-129)
# --------------------------------------------------------------------------
def _apply_1q(state: np.ndarray, u: np.ndarray, q: int) -> np.ndarray:
    n_amp = state.shape[0]
    lo = 1 << q
    s = state.reshape(n_amp // (2 * lo), 2, lo)
    out = np.empty_like(s)
    out[:, 0, :] = u[0, 0] * s[:, 0, :] + u[0, 1] * s[:, 1, :]
    out[:, 1, :] = u[1, 0] * s[:, 0, :] + u[1, 1] * s[:, 1, :]
    return out.reshape(n_amp)


def _apply_cx(state: np.ndarray, c: int, t: int) -> np.ndarray:
    idx = np.arange(state.shape[0])
    src = np.where((idx >> c) & 1, idx ^ (1 << t), idx)
    # new[idx] = old[idx with bit t flipped when bit c set]  (CX is an involution)
    return state[src]


def ry(theta: float) -> np.ndarray:
    c, s = np.cos(theta / 2.0), np.sin(theta / 2.0)
    return np.array([[c, -s], [s, c]], dtype=np.complex128)


def rz(phi: float) -> np.ndarray:
    return np.array([[np.exp(-0.5j * phi), 0.0], [0.0, np.exp(0.5j * phi)]], dtype=np.complex128)


def apply_feature_layers(
    state: np.ndarray, angles: np.ndarray, n_qubits: int, layers: int, first_layer_limit: Optional[int] = None
) -> np.ndarray:
    """Apply ``layers`` repetitions of [RY(pi a), RZ(pi a / 2) per qubit; CX chain].

    Layer ``l`` qubit ``i`` takes ``a = angles[(l * n_qubits + i) % len(angles)]``.
    With ``layers == 1`` this is exactly quantum.py:158-165, where only the first
    ``min(len(vector), n_qubits)`` qubits are rotated (``first_layer_limit``).
    """
    m = len(angles)
    for layer in range(layers):
        for i in range(n_qubits):
            k = layer * n_qubits + i
            if layers == 1 and first_layer_limit is not None and i >= first_layer_limit:
                continue
            a = float(angles[k % m])
            state = _apply_1q(state, ry(a * PI), i)          # quantum.py:160
            state = _apply_1q(state, rz(a * PI / 2.0), i)    # quantum.py:161
        for i in range(n_qubits - 1):                        # quantum.py:164-165
            state = _apply_cx(state, i, i + 1)
    return state


def circuit_statevector(vector: Sequence[float], n_qubits: int = 4, layers: int = 1) -> np.ndarray:
    """Statevector of ``_vector_to_circuit(vector)`` (quantum.py:138-167).

    ``layers > 1`` is a builder-defined extension: the reference block is
    repeated, layer ``l`` reading components ``l*n .. l*n+n-1`` (mod len).
    """
    v = np.asarray(vector, dtype=np.float64)
    norm = np.linalg.norm(v)                                  # quantum.py:149
    if norm > 0:
        v = v / norm                                          # quantum.py:150-151
    state = np.zeros(1 << n_qubits, dtype=np.complex128)
    state[0] = 1.0
    if len(v) == 0:
        # no rotations; only the CX chain, which fixes |0..0>
        return state
    limit = min(len(v), n_qubits)                             # quantum.py:158
    return apply_feature_layers(state, v, n_qubits, layers, first_layer_limit=limit)


def state_fidelity(psi1: np.ndarray, psi2: np.ndarray) -> float:
    """``qiskit.quantum_info.state_fidelity`` for two pure states (quantum.py:132)."""
    return float(np.abs(np.vdot(psi2, psi1)) ** 2)


def quantum_similarity(
    vec1: Sequence[float], vec2: Sequence[float], n_qubits: int = 4, method: str = "state_fidelity", layers: int = 1
) -> float:
    """quantum.py:108-136 -- fidelity, or the constant 0.5 for any other method."""
    if method == "state_fidelity":
        return state_fidelity(
            circuit_statevector(vec1, n_qubits, layers), circuit_statevector(vec2, n_qubits, layers)
        )
    return 0.5                                                # quantum.py:134-136


def closed_form_fidelity(vec1: Sequence[float], vec2: Sequence[float], n_qubits: int = 4) -> float:
    """Independent cross-check for ``layers == 1``.

    Both states share the trailing CX chain (a permutation, hence unitary and
    cancelling in the overlap) and are product states before it, so with
    ``t = pi v``, ``p = pi v / 2``:

        F = prod_i [ cos^2(dp_i/2) cos^2((ta_i - tb_i)/2) + sin^2(dp_i/2) cos^2((ta_i + tb_i)/2) ]
    """
    a = np.asarray(vec1, dtype=np.float64)
    b = np.asarray(vec2, dtype=np.float64)
    na, nb = np.linalg.norm(a), np.linalg.norm(b)
    if na > 0:
        a = a / na
    if nb > 0:
        b = b / nb
    f = 1.0
    for i in range(n_qubits):
        ai = a[i] if i < len(a) else 0.0
        bi = b[i] if i < len(b) else 0.0
        ta, tb = PI * ai, PI * bi
        dp = (PI / 2.0) * (ai - bi)
        f *= np.cos(dp / 2) ** 2 * np.cos((ta - tb) / 2) ** 2 + np.sin(dp / 2) ** 2 * np.cos((ta + tb) / 2) ** 2
    return float(f)


# --------------------------------------------------------------------------
# quantum.py:44-106  rerank / _quantum_score_documents
# --------------------------------------------------------------------------
def stable_rank(scores: Sequence[float], top_k: Optional[int] = None) -> List[int]:
    """Indices in the order ``sorted(..., key=score, reverse=True)[:top_k]`` yields.

    quantum.py:70-76: Python's sort is stable and ``reverse=True`` preserves the
    input order of equal keys, so the order is (score desc, input index asc);
    the slice is applied whenever ``top_k is not None`` (0 -> empty, negative ->
    Python slice semantics).
    """
    order = sorted(range(len(scores)), key=lambda i: scores[i], reverse=True)
    if top_k is not None:
        order = order[:top_k]
    return order


def quantum_rerank_strings(
    query: str, contents: Sequence[str], top_k: Optional[int] = None, n_qubits: int = 4,
    method: str = "state_fidelity",
) -> List[Tuple[int, float]]:
    """(input index, score) pairs in the order QuantumReranker.rerank returns them.

    Follows the reference's structure on purpose (quantum.py:94-104): the query
    embedding is computed once, but the query *state* is re-simulated for every
    document (quantum.py:121,126).
    """
    if len(contents) == 0:                                    # quantum.py:63-64
        return []
    qv = mock_embedding(query, n_qubits)                      # quantum.py:94
    scores = []
    for text in contents:                                     # quantum.py:98
        dv = mock_embedding(text, n_qubits)                   # quantum.py:100
        scores.append(quantum_similarity(qv, dv, n_qubits, method))
    return [(i, scores[i]) for i in stable_rank(scores, top_k)]


# --------------------------------------------------------------------------
# Builder-defined extensions (no reference counterpart; the reference's own
# comment at quantum.py:156 names amplitude encoding as the "real" scheme).
# --------------------------------------------------------------------------
def amplitude_state(x: Sequence[float], n_qubits: int) -> np.ndarray:
    """|x> = zero-pad(x) / ||x|| on ``n_qubits`` qubits.  Zero vector -> zeros."""
    x = np.asarray(x, dtype=np.float64)
    if x.shape[0] > (1 << n_qubits):
        raise ValueError("vector longer than 2**n_qubits")
    psi = np.zeros(1 << n_qubits, dtype=np.complex128)
    nrm2 = float(np.dot(x, x))
    if nrm2 > 0:
        psi[: x.shape[0]] = x / np.sqrt(nrm2)
    return psi


def amplitude_fidelity(q: Sequence[float], d: Sequence[float]) -> float:
    """``|<q^|d^>|^2 = (q.d)^2 / (|q|^2 |d|^2)`` in fp64; 0 if either is zero."""
    q = np.asarray(q, dtype=np.float64)
    d = np.asarray(d, dtype=np.float64)
    nq2, nd2 = float(np.dot(q, q)), float(np.dot(d, d))
    if nq2 == 0.0 or nd2 == 0.0:
        return 0.0
    qd = float(np.dot(q, d))
    return qd * qd / (nq2 * nd2)


def amplitude_fidelity_batch(Q: np.ndarray, cand: np.ndarray) -> np.ndarray:
    """Vectorised amplitude fidelity: Q [nq, D] fp32, cand [nq, C, D] fp32 -> [nq, C] fp64."""
    Q64 = Q.astype(np.float64)
    out = np.empty(cand.shape[:2], dtype=np.float64)
    nq2 = np.einsum("qd,qd->q", Q64, Q64)
    for i in range(cand.shape[0]):
        c64 = cand[i].astype(np.float64)
        # row-wise pairwise sums (not a BLAS gemv): every row is reduced in the same order, so
        # duplicate candidates get bit-identical scores and tie exactly, as they do in the reference
        qd = (c64 * Q64[i]).sum(axis=1)
        nd2 = (c64 * c64).sum(axis=1)
        den = nq2[i] * nd2
        with np.errstate(divide="ignore", invalid="ignore"):
            f = np.where(den > 0, qd * qd / np.where(den > 0, den, 1.0), 0.0)
        out[i] = f
    return out


def feature_map_state(x: Sequence[float], n_qubits: int, layers: int) -> np.ndarray:
    """Amplitude-encoded |x^> followed by ``layers`` reference-style blocks.

    Layer ``l`` qubit ``i`` is rotated by ``a = x^[(l*n + i) % D]`` (the
    L2-normalised input), RY(pi a) then RZ(pi a / 2), then the CX chain
    (builder-defined).
    """
    x = np.asarray(x, dtype=np.float64)
    psi = amplitude_state(x, n_qubits)
    nrm = np.sqrt(float(np.dot(x, x)))
    if nrm == 0.0 or layers == 0:
        return psi
    xh = x / nrm
    return apply_feature_layers(psi, xh, n_qubits, layers)


def feature_map_fidelity(q: Sequence[float], d: Sequence[float], n_qubits: int, layers: int) -> float:
    return state_fidelity(feature_map_state(q, n_qubits, layers), feature_map_state(d, n_qubits, layers))


def rank_rows(scores: np.ndarray, top_k: Optional[int] = None) -> np.ndarray:
    """Row-wise (score desc, index asc) order of a 2-D score array."""
    nq, c = scores.shape
    idx = np.broadcast_to(np.arange(c), (nq, c))
    order = np.lexsort((idx, -scores), axis=1)
    if top_k is not None:
        order = order[:, :top_k]
    return order
