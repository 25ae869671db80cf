"""Flat-index oracle: the reference's fixture, IxF2 format, canonical ordering (CPU)."""
import hashlib

import numpy as np

from oracle import search as osr


def test_fixture_identity(piers, kat):
    x = piers["vectors"]
    assert x.shape == (119, 1536) and x.dtype == np.float32
    assert int(piers["metric_type"]) == 1
    raw = osr.write_ixf(x, 1)
    assert len(raw) == 731181
    assert hashlib.sha256(raw).hexdigest() == str(piers["sha256"])          # byte-exact re-encoding
    assert str(piers["sha256"]).startswith("a8f1ba4e") and str(piers["sha256"]).endswith("1373dd")
    back = osr.read_ixf(raw)
    assert back["d"] == 1536 and back["ntotal"] == 119 and back["metric_type"] == 1 and back["is_trained"]
    assert np.array_equal(back["vectors"], x)
    norms = np.linalg.norm(x.astype(np.float64), axis=1)
    assert norms.min() > 0.9999999 and norms.max() < 1.0000002


def test_fixture_top20_and_duplicates(piers, kat):
    x = piers["vectors"]
    s, i = osr.exact_search(x[:1], x, 20, osr.METRIC_L2)
    assert i[0].tolist() == kat["survey"]["fixture_top20_row0"]
    assert np.array_equal(osr.exact_search(x, x, 20, osr.METRIC_L2)[1], piers["top20_ids"])
    for grp in kat["survey"]["fixture_duplicate_groups"]:
        for g in grp[1:]:
            assert np.array_equal(x[g], x[grp[0]])
        # a duplicate group queried by its own member: all members tie at distance 0, ids ascending
        s, i = osr.exact_search(x[grp[0]:grp[0] + 1], x, len(grp), osr.METRIC_L2)
        assert i[0].tolist() == grp and np.all(s[0] == 0.0)
    assert len({r.tobytes() for r in x}) == 106
    # unit-norm rows: L2 and cosine / IP rankings coincide away from ties
    _, i_cos = osr.exact_search(x[:1], x, 5, osr.METRIC_COSINE)
    assert i_cos[0].tolist() == kat["survey"]["fixture_top20_row0"][:5]


def test_padding_and_merge():
    rng = np.random.RandomState(3)
    X = rng.standard_normal((50, 8)).astype(np.float32)
    X[10] = X[40]
    Q = rng.standard_normal((4, 8)).astype(np.float32)
    for metric in (osr.METRIC_IP, osr.METRIC_L2, osr.METRIC_COSINE):
        s, i = osr.exact_search(Q, X, 60, metric)
        assert np.all(i[:, 50:] == -1)
        assert np.all(np.isinf(s[:, 50:]))
        full_s, full_i = osr.exact_search(Q, X, 12, metric)
        parts = [osr.exact_search(Q, X[a:b], 12, metric, id_base=a) for a, b in ((0, 7), (7, 30), (30, 50))]
        ms, mi = osr.merge_topk(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), 12, metric)
        assert np.array_equal(mi, full_i)
        assert np.array_equal(ms, full_s)
