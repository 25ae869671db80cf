#!/usr/bin/env python
"""Regenerate the committed fixtures under tests/golden/.

Run in the authoring container only (it reads /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

Outputs
* ``piers_index.npz``  -- the reference's only data fixture
  (``mcp/piers_morgan_faiss_index.faiss`` + ``_metadata.pkl``) re-encoded: the
  fp32 matrix exactly as stored, the 119 labels, the file's sha256 and the
  exact L2 top-20 of every row (fp64 arithmetic, ties by id).
* ``kat.json``         -- known answers.  ``survey`` holds the values recorded in
  SURVEY.md section 8c by an independent restatement made before this oracle
  existed; ``oracle`` holds this oracle's outputs on a wider set of inputs
  (regression lock; each one is cross-checked against the closed form).
"""
import hashlib
import io
import json
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import quantum as oq  # noqa: E402
from oracle import search as osr  # noqa: E402

REF = "/root/reference/mcp"


class _ListOfStrUnpickler(pickle.Unpickler):
    """Refuses every global: only builtin containers / str can be produced."""

    def find_class(self, module, name):  # pragma: no cover - must never trigger
        raise pickle.UnpicklingError(f"global {module}.{name} forbidden")


def main():
    raw = open(os.path.join(REF, "piers_morgan_faiss_index.faiss"), "rb").read()
    ix = osr.read_ixf(raw)
    x = np.array(ix["vectors"], dtype=np.float32)
    labels = _ListOfStrUnpickler(
        io.BytesIO(open(os.path.join(REF, "piers_morgan_faiss_index_metadata.pkl"), "rb").read())
    ).load()
    assert isinstance(labels, list) and all(isinstance(s, str) for s in labels)
    s20, i20 = osr.exact_search(x, x, 20, osr.METRIC_L2)
    np.savez_compressed(
        os.path.join(HERE, "piers_index.npz"),
        vectors=x,
        labels=np.array(labels),
        sha256=np.array(hashlib.sha256(raw).hexdigest()),
        metric_type=np.array(ix["metric_type"]),
        top20_ids=i20.astype(np.int64),
        top20_dist=s20,
    )

    query = "which segments contain a sponsor advertisement"
    survey = {
        "query": query,
        "query_charsum": 4597,
        "query_embedding_head": [0.5000056306886407, 0.08813468330103492, 0.606848294325272, 0.1690918857786345],
        "psi0": [0.18741385260416357, -0.3435076565355873],
        "psi15": [0.3754258165780682, -0.11037361922668742],
        "docs_n4": [
            ["This episode is brought to you by ExampleVPN, use code PIERS for a discount.", 6825, 0.40812770330140113],
            ["Today we discuss the election results with our panel.", 5073, 0.39463405284145719],
            ["Piers_Morgan_Uncensored/02c7ef143b0f1975e8100545e5a4045ef238e4a199c7ea8e39c21b790d3c781b", 6749,
             0.49948053547920807],
            ["a", 97, 0.39462324127141984],
            ["", 0, 0.51356955928427517],
            ["ab", 195, 0.5529677313846768],
            ["ba", 195, 0.5529677313846768],
        ],
        "doc6825_n9": 0.53347220304400933,
        "doc6825_n10": 0.50524499663079137,
        "legacy_rng_seed0": [0.5488135, 0.71518937, 0.60276338],
        "fixture_top20_row0": [0, 1, 88, 47, 31, 87, 43, 29, 105, 69, 72, 81, 108, 56, 63, 3, 46, 32, 17, 41],
        "fixture_duplicate_groups": [[4, 24, 49, 52, 83, 86, 90, 96, 112], [15, 45, 79], [26, 54], [28, 55], [66, 73]],
    }

    rng = np.random.RandomState(20261018)
    cases = []
    for n in (1, 2, 3, 4, 5, 6, 9, 10):
        for layers in (1, 2, 4):
            for vec_len in (2 * n, max(1, n - 1)):
                a = rng.random_sample(vec_len)
                b = rng.random_sample(vec_len)
                f = oq.quantum_similarity(a, b, n, layers=layers)
                if layers == 1:
                    assert abs(f - oq.closed_form_fidelity(a, b, n)) < 1e-13
                cases.append({"n": n, "layers": layers, "a": a.tolist(), "b": b.tolist(), "f": f})
    amp = []
    for d, n in ((3, 2), (8, 3), (384, 9), (1536, 11)):
        q = rng.standard_normal(d).astype(np.float32)
        c = rng.standard_normal(d).astype(np.float32)
        amp.append({"n": n, "q": q.tolist(), "d": c.tolist(), "f": oq.amplitude_fidelity(q, c)})
    fmap = []
    for d, n, layers in ((8, 3, 2), (40, 6, 4), (1024, 10, 4)):
        q = rng.standard_normal(d).astype(np.float32)
        c = rng.standard_normal(d).astype(np.float32)
        fmap.append({"n": n, "layers": layers, "q": q.tolist(), "d": c.tolist(),
                     "f": oq.feature_map_fidelity(q, c, n, layers)})
    with open(os.path.join(HERE, "kat.json"), "w") as fh:
        json.dump({"survey": survey, "oracle": {"angle": cases, "amplitude": amp, "feature_map": fmap}}, fh)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
