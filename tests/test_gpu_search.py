"""K3a/K5 parity: exact flat search and list merge vs the oracle, incl. the reference's own index."""
import numpy as np
import pytest

from oracle import search as osr

pytestmark = pytest.mark.gpu

METRICS = [("ip", osr.METRIC_IP), ("l2", osr.METRIC_L2), ("cosine", osr.METRIC_COSINE)]


def _check(api, Q, X, k, name, mid, id_base=0):
    s, i = api.search_topk(Q, X, k, name, id_base=id_base)
    rs, ri = osr.exact_search(Q, X, k, mid, id_base=id_base)
    assert np.array_equal(i.cpu().numpy(), ri), name
    got = s.cpu().numpy()
    fin = np.isfinite(rs)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.allclose(got[fin], rs[fin], rtol=1e-12, atol=1e-13)
    assert np.array_equal(got[~fin], rs[~fin])


@pytest.mark.parametrize("name,mid", METRICS)
@pytest.mark.parametrize("N,D,k", [(1, 8, 1), (119, 1536, 20), (50, 10, 60), (5000, 64, 100), (20000, 384, 1000),
                                   (9000, 33, 7), (4097, 128, 2048)])
def test_exact_search_matches_oracle(cuda, name, mid, N, D, k):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(N + D + k)
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((5, D)).astype(np.float32)
    if N > 40:
        X[17] = X[3]
        X[N - 1] = X[3]                               # duplicates across chunk boundaries
        Q[0] = X[3]
    _check(api, Q, X, k, name, mid, id_base=1000)


def test_reference_fixture_index(cuda, piers, kat):
    """Config 1: the reference's own FAISS index; every row as the query, exact L2 top-20."""
    from quantum_rag_b200 import api
    x = piers["vectors"]
    s, i = api.search_topk(x, x, 20, "l2")
    assert np.array_equal(i.cpu().numpy(), piers["top20_ids"])
    assert i[0].cpu().tolist() == kat["survey"]["fixture_top20_row0"]
    assert np.allclose(s.cpu().numpy(), piers["top20_dist"], rtol=1e-12, atol=1e-13)
    for grp in kat["survey"]["fixture_duplicate_groups"]:
        s, i = api.search_topk(x[grp[0]:grp[0] + 1], x, len(grp), "l2")
        assert i[0].cpu().tolist() == grp
        assert np.all(s[0].cpu().numpy() == 0.0)


def test_empty_corpus_and_unsupported_k(cuda):
    from quantum_rag_b200 import api
    from quantum_rag_b200._lib import QragError
    s, i = api.search_topk(np.ones((2, 8), np.float32), np.zeros((0, 8), np.float32), 3, "ip")
    assert i.cpu().tolist() == [[-1] * 3] * 2 and np.all(np.isneginf(s.cpu().numpy()))
    with pytest.raises(QragError):
        api.search_topk(np.ones((1, 8), np.float32), np.ones((10, 8), np.float32), 4096, "ip")


@pytest.mark.parametrize("name,mid", METRICS)
def test_merge_equals_single_shard_search(cuda, name, mid):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(5)
    X = rng.standard_normal((3000, 48)).astype(np.float32)
    X[2500] = X[10]
    Q = rng.standard_normal((7, 48)).astype(np.float32)
    Q[1] = X[10]
    k = 50
    bounds = [0, 40, 1000, 1700, 3000]                # first shard shorter than k -> padding ids
    parts = [api.search_topk(Q, X[a:b], k, name, id_base=a) for a, b in zip(bounds[:-1], bounds[1:])]
    import torch
    ms, mi = api.topk_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]), k, name)
    fs, fi = api.search_topk(Q, X, k, name)
    assert torch.equal(mi, fi)
    assert torch.equal(ms, fs)                        # bit-identical: shard scores are computed the same way
    rs, ri = osr.exact_search(Q, X, k, mid)
    assert np.array_equal(mi.cpu().numpy(), ri)
    # oracle merge agrees as well, including k_out > available
    os_, oi = osr.merge_topk(np.stack([p[0].cpu().numpy() for p in parts[:1]]),
                             np.stack([p[1].cpu().numpy() for p in parts[:1]]), 60, mid)
    gs, gi = api.topk_merge(parts[0][0][None], parts[0][1][None], 60, name)
    assert np.array_equal(gi.cpu().numpy(), oi)


@pytest.mark.parametrize("G,k,k_out,name,mid", [(20, 1000, 1000, "cosine", osr.METRIC_COSINE), (300, 100, 100, "l2", osr.METRIC_L2),
                                              (9, 1000, 10, "ip", osr.METRIC_IP), (40, 2048, 2048, "l2", osr.METRIC_L2),
                                              (8, 1000, 1000, "ip", osr.METRIC_IP)])
def test_merge_of_more_entries_than_one_kernel_sorts(cuda, G, k, k_out, name, mid):
    """G * k above 8192 entries per query (16 shards x top-1000, or many small shards): merged over several
    levels through a workspace, same canonical order as the oracle's one-shot merge; padding and ties included."""
    import torch
    from quantum_rag_b200 import api
    rng = np.random.RandomState(G + k)
    nq = 3
    scores = np.round(rng.standard_normal((G, nq, k)), 2)                     # coarse: many score ties across shards
    ids = np.stack([rng.permutation(G * k * 2)[:G * k].reshape(G, k) for _ in range(nq)], axis=1).astype(np.int64)
    key = scores if mid == osr.METRIC_L2 else -scores
    order = np.lexsort((ids, key), axis=2)                                    # every shard list sorted canonically
    scores, ids = np.take_along_axis(scores, order, 2), np.take_along_axis(ids, order, 2)
    short = rng.randint(0, G, size=3)                                         # some shards hold fewer than k entries
    for g in short:
        n = rng.randint(0, k)
        ids[g, :, n:] = -1
        scores[g, :, n:] = np.inf if mid == osr.METRIC_L2 else -np.inf
    ws, wi = osr.merge_topk(scores, ids, k_out, mid)
    gs, gi = api.topk_merge(torch.from_numpy(scores).cuda(), torch.from_numpy(ids).cuda(), k_out, name)
    assert np.array_equal(gi.cpu().numpy(), wi)
    assert np.array_equal(gs.cpu().numpy(), ws)
