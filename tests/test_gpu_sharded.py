"""Config 4 path on the GPU: search -> merge -> owner-computes quantum rerank, for 1 and for G shards.

One GPU is enough to check what matters for the multi-GPU claim: with the CUDA engine the G-shard
result (shards searched one after another on the same device, exchange steps emulated with
stack/max) is bit-identical to the 1-shard result.  The real NCCL exchange is covered by
tools/sharded_check.py under torchrun and by tests/test_sharded_gloo.py on CPU.
"""
import numpy as np
import pytest

from oracle import quantum as oq
from oracle import search as osr

pytestmark = pytest.mark.gpu


def _emulated(X, Q, G, k1, k2, metric, phased):
    import torch
    from quantum_rag_b200.sharded import CudaEngine, shard_bounds
    n = X.shape[0]
    engines = []
    for r in range(G):
        lo, hi = shard_bounds(n, G, r)
        engines.append((lo, hi, CudaEngine(X[lo:hi], metric, lo)))
    if phased:
        # the three phases of every shard with the exchanges emulated by stack (what NCCL does under torchrun)
        bound = torch.stack([e.index.aux[:2] for _, _, e in engines]).max(dim=0).values
        for _, _, e in engines:
            e.index.aux[:2] = bound
        bm_all = torch.stack([e.index.tc_begin(Q, k1, G) for _, _, e in engines])
        hist_all = torch.stack([e.index.tc_filter(bm_all) for _, _, e in engines]).sum(dim=0, dtype=torch.int32)
        lists = []
        for _, _, e in engines:
            s, i, status = e.index.tc_finish(hist_all)
            assert int(status.count_nonzero()) == 0
            lists.append((s, i))
    else:
        lists = [e.search(Q, k1) for _, _, e in engines]
    gs = torch.stack([s for s, _ in lists])
    gi = torch.stack([i for _, i in lists])
    ss, si = engines[0][2].merge(gs, gi, k1)
    f = None
    for lo, hi, e in engines:
        own = (si >= lo) & (si < hi)
        local = torch.where(own, si - lo, torch.full_like(si, -1))
        fr = e.fidelity_rows(Q, local)
        f = fr if f is None else torch.maximum(f, fr)
    pos, top = engines[0][2].sort_scores(f, k2)
    return top, torch.gather(si, 1, pos.long()), ss, si


def _emulated_packed(X, Q, G, k1, k2, metric, check_owner_vs_oracle=True):
    """The packed form for G shards on one device: phases of every shard in lockstep, the two threshold
    exchanges emulated by stack / sum, the all-to-all by slicing every shard's send buffer per owner."""
    import torch
    from quantum_rag_b200 import api
    from quantum_rag_b200.sharded import CudaEngine, exchange_len, shard_bounds
    n, nq = X.shape[0], Q.shape[0]
    engines = [CudaEngine(X[lo:hi], metric, lo) for lo, hi in (shard_bounds(n, G, r) for r in range(G))]
    bound = torch.stack([e.index.aux[:2] for e in engines]).max(dim=0).values
    for e in engines:
        e.index.aux[:2] = bound
    kk, per = exchange_len(k1, G), -(-nq // G)
    assert kk == api.exchange_len(k1, G)
    bm = [e.packed_begin(Q, k1, G) for e in engines]
    bm_all = torch.stack(bm) if G > 1 else None
    hists = [e.packed_filter(bm_all) for e in engines]
    ap_all = torch.stack(hists).sum(dim=0, dtype=torch.int32) if G > 1 else None       # the all-reduce(SUM)
    sends = []
    for e in engines:
        send = torch.zeros((per * G, 3 * kk + 1), dtype=torch.int64, device=X.device)
        e.packed_finish(ap_all, kk, send)
        sends.append(send.view(G, per, -1))
    outs = []
    mid = {"cosine": osr.METRIC_COSINE, "l2": osr.METRIC_L2, "ip": osr.METRIC_IP}[metric]
    for r in range(G):
        recv = torch.stack([snd[r] for snd in sends]).contiguous()                # [G, per, rec]: what rank r receives
        out = engines[r].owner_finalize(recv, kk, k1, min(k2, k1), r * per, nq, None)
        if check_owner_vs_oracle:
            want = osr.owner_finalize(recv.cpu().numpy(), kk, k1, min(k2, k1), mid, r * per, nq)
            assert np.array_equal(out.cpu().numpy(), want), f"owner kernel != oracle (rank {r} of {G})"
        outs.append(out)
    res = torch.cat(outs)[:nq]
    k2 = min(k2, k1)
    assert int(res[:, 2 * k2].max()) == 0, "a shard flagged a query"
    return res[:, :k2].contiguous().view(torch.float64), res[:, k2:2 * k2].contiguous()


@pytest.mark.parametrize("metric", ["cosine", "l2", "ip"])
def test_sharded_path_matches_oracle_and_single_shard(cuda, metric):
    import torch
    from quantum_rag_b200.sharded import ShardedSearchRerank
    rng = np.random.RandomState(4)
    n, d, nq, k1, k2 = 30000, 384, 12, 200, 10
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    X[n - 2] = X[1]
    X[n // 2 + 5] = X[1]
    Q[0] = X[1]
    Xd, Qd = torch.from_numpy(X).cuda(), torch.from_numpy(Q).cuda()
    path = ShardedSearchRerank(Xd, n, metric)                            # world size 1, CUDA engine
    res = path(Qd, k1, k2, return_search_lists=True)                     # the all-gather form (materialises the lists)
    own = path(Qd, k1, k2)                                               # the packed form
    assert own.search_ids is None and path.last_rerun == 0
    assert torch.equal(own.ids, res.ids) and torch.equal(own.scores, res.scores)
    # oracle: exact search, then amplitude fidelity of the found rows, stable order
    mid = {"cosine": osr.METRIC_COSINE, "l2": osr.METRIC_L2, "ip": osr.METRIC_IP}[metric]
    rs, ri = osr.exact_search(Q, X, k1, mid)
    assert np.array_equal(res.search_ids.cpu().numpy(), ri)
    f = oq.amplitude_fidelity_batch(Q, X[ri])
    order = oq.rank_rows(f, k2)
    assert np.array_equal(res.ids.cpu().numpy(), np.take_along_axis(ri, order, 1))
    assert np.allclose(res.scores.cpu().numpy(), np.take_along_axis(f, order, 1), rtol=1e-12, atol=1e-16)
    for G, phased in ((2, False), (2, True), (3, True), (8, True)):
        top, ids, ss, si = _emulated(Xd, Qd, G, k1, k2, metric, phased)
        assert torch.equal(si, res.search_ids) and torch.equal(ss, res.search_scores), f"G={G} phased={phased}"
        assert torch.equal(ids, res.ids) and torch.equal(top, res.scores), f"G={G} phased={phased}"
    for G in (1, 2, 3, 5, 8):
        top, ids = _emulated_packed(Xd, Qd, G, k1, k2, metric)
        assert torch.equal(ids, res.ids) and torch.equal(top, res.scores), f"packed G={G}"


@pytest.mark.parametrize("metric", ["cosine", "l2"])
def test_sharded_path_many_queries(cuda, metric):
    """Enough queries that one CTA rescoring a query's whole list also sorts and packs it (the fused tail of
    tc_rescore_bulk_kernel, lists of several 128-candidate chunks); fewer queries take the separate sort kernels."""
    import torch
    from quantum_rag_b200.sharded import ShardedSearchRerank
    rng = np.random.RandomState(14)
    n, d, nq, k1, k2 = 24000, 384, 640, 200, 10
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    X[n - 3] = X[2]
    Q[5] = X[2]
    Xd, Qd = torch.from_numpy(X).cuda(), torch.from_numpy(Q).cuda()
    path = ShardedSearchRerank(Xd, n, metric)
    res = path(Qd, k1, k2, return_search_lists=True)
    own = path(Qd, k1, k2)
    assert path.last_rerun == 0
    assert torch.equal(own.ids, res.ids) and torch.equal(own.scores, res.scores)
    mid = {"cosine": osr.METRIC_COSINE, "l2": osr.METRIC_L2}[metric]
    rs, ri = osr.exact_search(Q, X, k1, mid)
    assert np.array_equal(res.search_ids.cpu().numpy(), ri)
    f = oq.amplitude_fidelity_batch(Q, X[ri])
    order = oq.rank_rows(f, k2)
    assert np.array_equal(res.ids.cpu().numpy(), np.take_along_axis(ri, order, 1))
    assert np.allclose(res.scores.cpu().numpy(), np.take_along_axis(f, order, 1), rtol=1e-12, atol=1e-16)
    for G in (2, 8):
        top, ids = _emulated_packed(Xd, Qd, G, k1, k2, metric)
        assert torch.equal(ids, res.ids) and torch.equal(top, res.scores), f"packed G={G}"


def test_owner_finalize_vs_oracle_random_records(cuda):
    """qrag_owner_finalize against the NumPy restatement on synthetic records: ties in score (id decides) and in
    fidelity (merged position decides), short and empty lists, flagged shards, padding queries, k2 == k1."""
    import torch
    from quantum_rag_b200 import api
    rng = np.random.RandomState(11)
    for G, per, kk, k1, k2, metric in ((1, 3, 40, 40, 40, "cosine"), (2, 4, 17, 20, 5, "l2"), (8, 5, 320, 1000, 10, "cosine"),
                                       (3, 2, 64, 100, 100, "ip"), (5, 3, 8, 64, 3, "l2")):
        mid = {"cosine": osr.METRIC_COSINE, "l2": osr.METRIC_L2, "ip": osr.METRIC_IP}[metric]
        recv = np.zeros((G, per, 3 * kk + 1), dtype=np.int64)
        for j in range(per):
            ids = rng.permutation(G * kk * 4)[:G * kk].reshape(G, kk)
            for g in range(G):
                n = int(rng.choice([0, 1, kk // 2, kk]))
                sc = np.round(rng.standard_normal(n), 1)                          # coarse: many exact score ties
                fid = np.round(rng.random_sample(n), 1)                           # and fidelity ties
                key = sc if mid == osr.METRIC_L2 else -sc
                order = np.lexsort((ids[g, :n], key))
                s = np.full((1, kk), 0.0); i = np.full((1, kk), -1, dtype=np.int64); f = np.full((1, kk), 0.0)
                s[0, :n], i[0, :n], f[0, :n] = sc[order], ids[g, :n][order], fid[order]
                recv[g, j] = osr.pack_records(s, i, f, kk, mid, bad=np.array([int(rng.random_sample() < 0.1)]))[0]
        nq = 100 + per - 1                                                        # the last owned query is padding
        want = osr.owner_finalize(recv, kk, k1, k2, mid, 100, nq)
        got = api.owner_finalize(torch.from_numpy(recv).cuda(), kk, k1, k2, metric, 100, nq)
        assert np.array_equal(got.cpu().numpy(), want), (G, per, kk, k1, k2, metric)


def test_submit_pipelined_and_graph_replay(cuda):
    """``submit`` queues batches without a host sync (results read later, still verified); ``graph=True`` replays the
    batch from a CUDA graph.  Both give the bits of the synchronous call, for changing queries of one shape."""
    import torch
    from quantum_rag_b200.sharded import ShardedSearchRerank
    rng = np.random.RandomState(21)
    X = torch.from_numpy(rng.standard_normal((20000, 384)).astype(np.float32)).cuda()
    Qs = [torch.from_numpy(rng.standard_normal((9, 384)).astype(np.float32)).cuda() for _ in range(4)]
    path = ShardedSearchRerank(X, 20000, "cosine")
    want = [path(Q, 100, 10) for Q in Qs]
    pend = [path.submit(Q, 100, 10) for Q in Qs]                       # four batches in flight
    for p, w in zip(pend, want):
        r = p.result()
        assert torch.equal(r.ids, w.ids) and torch.equal(r.scores, w.scores)
    pend = [path.submit(Q, 100, 10, graph=True) for Q in Qs]           # captured on the first call, replayed after
    for p, w in zip(pend, want):
        r = p.result()
        assert torch.equal(r.ids, w.ids) and torch.equal(r.scores, w.scores)
    path.close()


def test_functional_form_and_cache(cuda):
    """``sharded_search_rerank(q, X_shard, k1, k2, group)`` = the class, shard prepared once."""
    import torch
    from quantum_rag_b200 import sharded
    rng = np.random.RandomState(9)
    X = torch.from_numpy(rng.standard_normal((5000, 384)).astype(np.float32)).cuda()
    Q = torch.from_numpy(rng.standard_normal((5, 384)).astype(np.float32)).cuda()
    a = sharded.sharded_search_rerank(Q, X, k1=50, k2=7)
    path = next(iter(sharded._PATHS.values()))[1]
    b = sharded.sharded_search_rerank(Q, X, k1=50, k2=7)
    assert next(iter(sharded._PATHS.values()))[1] is path and len(sharded._PATHS) == 1
    ref = sharded.ShardedSearchRerank(X, 5000, "cosine")(Q, 50, 7)
    for r in (a, b):
        assert torch.equal(r.ids, ref.ids) and torch.equal(r.scores, ref.scores)


def test_config4_full_size_properties(cuda):
    """BASELINE config 4 at full corpus size (10M x 384, top-1000 -> rerank -> top-10), 128 queries: planted answers,
    the exact CUDA-core search on a subset, and shard invariance (8 emulated shards == 1 shard, bit for bit)."""
    import torch
    from quantum_rag_b200 import api
    from quantum_rag_b200.sharded import ShardedSearchRerank
    N, D, nq, k1, k2 = 10_000_000, 384, 128, 1000, 10
    X = torch.empty((N, D), dtype=torch.float32, device="cuda")
    for b in range(0, N, 1 << 20):                               # generated in blocks: no 15 GB temporaries
        g = torch.Generator(device="cuda").manual_seed(4000 + b)
        n = min(1 << 20, N - b)
        X[b:b + n] = torch.nn.functional.normalize(torch.randn(n, D, generator=g, device="cuda"), dim=1)
    g = torch.Generator(device="cuda").manual_seed(1234 + 4)
    Q = torch.nn.functional.normalize(torch.randn(nq, D, generator=g, device="cuda"), dim=1)
    planted = torch.arange(nq, device="cuda") * 78_001 + 11      # spread over all eight shards
    X[planted] = Q
    path = ShardedSearchRerank(X, N, "cosine")
    res = path(Q, k1, k2, return_search_lists=True)
    own = path(Q, k1, k2)
    assert path.last_rerun == 0 and torch.equal(own.ids, res.ids) and torch.equal(own.scores, res.scores)
    assert torch.equal(res.search_ids[:, 0], planted) and torch.equal(res.ids[:, 0], planted)
    assert torch.allclose(res.scores[:, 0], torch.ones(nq, dtype=torch.float64, device="cuda"), atol=1e-12)
    assert torch.all(res.search_scores[:, :-1] >= res.search_scores[:, 1:]) and torch.all(res.scores[:, :-1] >= res.scores[:, 1:])
    assert all(len(set(row)) == k1 for row in res.search_ids[:4].cpu().tolist())          # no duplicates in a list
    sub = torch.tensor([0, 63, 127], device="cuda")
    es, ei = api.search_topk(Q[sub], X, k1, "cosine")
    assert torch.equal(res.search_ids[sub], ei) and torch.equal(res.search_scores[sub], es)
    top, ids, ss, si = _emulated(X, Q, 8, k1, k2, "cosine", True)
    assert torch.equal(si, res.search_ids) and torch.equal(ss, res.search_scores)
    assert torch.equal(ids, res.ids) and torch.equal(top, res.scores)
    for G in (2, 8):
        top, ids = _emulated_packed(X, Q, G, k1, k2, "cosine", check_owner_vs_oracle=(G == 8))
        assert torch.equal(ids, res.ids) and torch.equal(top, res.scores), f"packed G={G}"
