"""Known-answer tests taken from the PUBLISHED definitions of the third-party code the reference calls, written
as literals -- not from the oracle's own output.

The reference's arithmetic lives in qiskit / qiskit-aer (quantum.py:12-13, 125-133, 158-165) and faiss
(store_in_faiss.py:99-109); neither runs in this image, so these literals are what pins ``oracle/`` to the
documented semantics:

* qiskit.circuit.library.RYGate:  RY(t) = [[cos t/2, -sin t/2], [sin t/2, cos t/2]]
* qiskit.circuit.library.RZGate:  RZ(l) = [[exp(-i l/2), 0], [0, exp(+i l/2)]]
* qiskit bit order: little-endian -- qubit 0 is the LEAST significant bit of the statevector index
* qiskit CXGate(control, target): flips the target where the control is 1
* qiskit.quantum_info.state_fidelity(psi, phi) for pure states = |<psi|phi>|^2
* faiss.IndexFlatL2.search: SQUARED L2 distances, ascending, int64 labels, label -1 where fewer than k rows exist
"""
import math

import numpy as np

from oracle import quantum as oq
from oracle import search as osr

R = 1.0 / math.sqrt(2.0)


def test_ry_rz_matrices_are_the_published_ones():
    assert np.allclose(oq.ry(math.pi / 2), np.array([[R, -R], [R, R]]), atol=1e-16)
    assert np.allclose(oq.ry(math.pi), np.array([[0.0, -1.0], [1.0, 0.0]]), atol=1e-15)          # = -iY: |0> -> |1>
    assert np.allclose(oq.rz(math.pi / 2), np.diag([complex(R, -R), complex(R, R)]), atol=1e-16)
    assert np.allclose(oq.rz(math.pi), np.diag([-1j, 1j]), atol=1e-15)
    assert np.allclose(oq.ry(0.0), np.eye(2)) and np.allclose(oq.rz(0.0), np.eye(2))


def test_little_endian_qubit_order():
    zero = np.zeros(4, dtype=np.complex128)
    zero[0] = 1.0
    # flipping qubit 0 of |00> lands on index 1 ("01"), flipping qubit 1 on index 2 ("10")
    assert np.allclose(oq._apply_1q(zero, oq.ry(math.pi), 0), [0, 1, 0, 0], atol=1e-15)
    assert np.allclose(oq._apply_1q(zero, oq.ry(math.pi), 1), [0, 0, 1, 0], atol=1e-15)
    zero3 = np.zeros(8, dtype=np.complex128)
    zero3[0] = 1.0
    assert np.allclose(oq._apply_1q(zero3, oq.ry(math.pi), 2), np.eye(8)[4], atol=1e-15)


def test_cx_control_and_target():
    basis = np.eye(4, dtype=np.complex128)
    # CX(0, 1): |01> (index 1: qubit 0 set) -> |11> (index 3); |10> (index 2: only qubit 1 set) is untouched
    assert np.array_equal(oq._apply_cx(basis[1], 0, 1), basis[3])
    assert np.array_equal(oq._apply_cx(basis[3], 0, 1), basis[1])
    assert np.array_equal(oq._apply_cx(basis[2], 0, 1), basis[2])
    assert np.array_equal(oq._apply_cx(basis[0], 0, 1), basis[0])
    # CX(1, 0): the roles swap
    assert np.array_equal(oq._apply_cx(basis[2], 1, 0), basis[3])
    assert np.array_equal(oq._apply_cx(basis[1], 1, 0), basis[1])


def test_bell_state_from_ry_and_cx():
    s = np.zeros(4, dtype=np.complex128)
    s[0] = 1.0
    s = oq._apply_cx(oq._apply_1q(s, oq.ry(math.pi / 2), 0), 0, 1)
    assert np.allclose(s, [R, 0, 0, R], atol=1e-16)


def test_state_fidelity_definition():
    zero, one = np.array([1, 0], dtype=np.complex128), np.array([0, 1], dtype=np.complex128)
    plus = np.array([R, R], dtype=np.complex128)
    assert oq.state_fidelity(zero, plus) == 0.5000000000000001 or abs(oq.state_fidelity(zero, plus) - 0.5) < 1e-15
    assert oq.state_fidelity(zero, one) == 0.0
    assert oq.state_fidelity(plus, plus) == 1.0000000000000002 or abs(oq.state_fidelity(plus, plus) - 1.0) < 1e-15
    psi = np.array([0.6, 0.8j], dtype=np.complex128)
    phi = np.array([0.8, -0.6j], dtype=np.complex128)
    # |<psi|phi>|^2 = |0.6*0.8 + conj(0.8j)*(-0.6j)|^2 = |0.48 - 0.48|^2 = 0
    assert abs(oq.state_fidelity(psi, phi)) < 1e-30
    assert abs(oq.state_fidelity(psi, np.exp(0.7j) * psi) - 1.0) < 1e-15            # a global phase does not matter
    assert oq.state_fidelity(psi, plus) == oq.state_fidelity(plus, psi)             # symmetric
    assert abs(oq.state_fidelity(psi, plus) - 0.5) < 1e-15                          # |0.6 R - 0.8i R|^2 = 0.5


def test_reference_circuit_literals():
    """quantum.py:158-165 on two qubits, amplitudes worked out from the definitions above by hand:
    RZ(p) RY(t)|0> = [exp(-ip/2) cos t/2, exp(+ip/2) sin t/2], product state, then CX(0, 1)."""
    # v = [1, 0]: qubit 0 gets RY(pi) RZ(pi/2) -> exp(i pi/4)|1>, qubit 1 stays |0>, CX(0,1) -> exp(i pi/4)|11>
    psi = oq.circuit_statevector([1.0, 0.0], 2)
    assert np.allclose(psi, [0, 0, 0, complex(R, R)], atol=1e-15)
    # v = [0.6, 0.8] (already unit length): theta = (0.6 pi, 0.8 pi), phi = (0.3 pi, 0.4 pi)
    psi = oq.circuit_statevector([0.6, 0.8], 2)
    want = [0.08246085134279689 - 0.16183853313827168j, 0.3493097717705924 + 0.685559027752571j,
            0.5521345675386733 + 0.08744952446344267j, 0.24692208514878447 - 0.03910861626005774j]
    assert np.allclose(psi, want, atol=2e-16 * 8)
    # the vector is renormalised first (quantum.py:149-151): [3, 4] is the same circuit as [0.6, 0.8]
    assert np.allclose(oq.circuit_statevector([3.0, 4.0], 2), want, atol=1e-15)
    # only the first min(len, n) qubits are rotated (quantum.py:158): one component on two qubits
    psi = oq.circuit_statevector([2.0], 2)                   # normalised to [1]: RY(pi) RZ(pi/2) on qubit 0 only
    assert np.allclose(psi, [0, 0, 0, complex(R, R)], atol=1e-15)
    # fidelity of the two literal states: F = |<a|b>|^2 with a = exp(i pi/4)|11>
    f = oq.quantum_similarity([1.0, 0.0], [0.6, 0.8], 2)
    assert abs(f - abs(complex(0.24692208514878447, -0.03910861626005774)) ** 2) < 1e-15


def test_faiss_flat_l2_contract():
    X = np.array([[0, 0], [3, 4], [1, 0], [0, 2]], dtype=np.float32)
    q = np.array([[0, 0]], dtype=np.float32)
    D, I = osr.exact_search(q, X, 4, osr.METRIC_L2)
    assert I.dtype == np.int64
    assert I[0].tolist() == [0, 2, 3, 1]                     # ascending distance
    assert D[0].tolist() == [0.0, 1.0, 4.0, 25.0]            # SQUARED L2: 25, not 5
    D, I = osr.exact_search(q, X, 6, osr.METRIC_L2)          # k > ntotal: label -1 in the unfilled slots
    assert I[0, 4:].tolist() == [-1, -1] and I[0, :4].tolist() == [0, 2, 3, 1]
    # faiss leaves its heap sentinel (FLT_MAX) as the distance of an unfilled slot; the canonical value here is +inf,
    # and either way it sorts after every real distance
    assert np.all(D[0, 4:] >= np.finfo(np.float32).max)
    # equal distances: faiss' order among ties is an implementation detail, the canonical rule is the smaller label
    X2 = np.array([[1, 0], [0, 1], [-1, 0], [0, -1]], dtype=np.float32)
    D, I = osr.exact_search(q, X2, 4, osr.METRIC_L2)
    assert I[0].tolist() == [0, 1, 2, 3] and np.all(D[0] == 1.0)
    # METRIC_INNER_PRODUCT: larger is better, descending
    D, I = osr.exact_search(np.array([[1, 1]], dtype=np.float32), X, 4, osr.METRIC_IP)
    assert I[0].tolist() == [1, 3, 2, 0] and D[0].tolist() == [7.0, 2.0, 1.0, 0.0]


def test_faiss_metric_type_constants_and_file_magic():
    # faiss MetricType: METRIC_INNER_PRODUCT = 0, METRIC_L2 = 1; write_index of an IndexFlatL2 starts with "IxF2"
    assert osr.METRIC_IP == 0 and osr.METRIC_L2 == 1
    raw = osr.write_ixf(np.zeros((2, 4), dtype=np.float32), 1)
    assert raw[:4] == b"IxF2" and len(raw) == 45 + 2 * 4 * 4
    assert int.from_bytes(raw[4:8], "little") == 4 and int.from_bytes(raw[8:16], "little") == 2
    assert osr.write_ixf(np.zeros((2, 4), dtype=np.float32), 0)[:4] == b"IxFI"
