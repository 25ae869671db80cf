"""CPU oracle for the quantum-rag reranking hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``quantum_rag_b200/`` or ``src/`` may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the
checker / reported CPU baseline, never as the thing shipped.

Parity status: **unpinned by reference tests** (the reference ships no tests,
goldens or benchmarks, and neither Qiskit nor FAISS can be imported in this
image).  The oracle is pinned instead by

* the known answers recorded in SURVEY.md section 8c (an independent NumPy
  restatement made during the survey), locked in ``tests/golden/kat.json``;
* a closed-form expression for the reference circuit's fidelity that is derived
  independently of the gate-by-gate simulation (``quantum.closed_form_fidelity``);
* the reference's own data fixture (``mcp/piers_morgan_faiss_index.faiss``),
  re-encoded as ``tests/golden/piers_index.npz`` by ``tests/golden/make_golden.py``.
"""
