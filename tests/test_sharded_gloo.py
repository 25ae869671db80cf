"""Multi-rank logic of the sharded search + rerank on CPU: world_size 2 and 3 over gloo.

The exchange steps run for real over ``torch.distributed``: row partition and id_base, the packed form (two
threshold all-gathers, ONE all-to-all of per-query records to the query's owner, owner merge, result all-gather)
and the all-gather form (lists all-gathered, merge, owner-computes rerank, all-reduce(MAX), stable sort).  The
per-rank compute is the NumPy oracle plugged in as the engine (the CUDA engine is exercised by
tests/test_gpu_sharded.py).  The G-rank result must equal the 1-rank result bit for bit, ties included.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import quantum as oq
from oracle import search as osr
from quantum_rag_b200.sharded import ShardedSearchRerank, exchange_len, shard_bounds


class OracleEngine:
    """CPU stand-in for CudaEngine: every per-row number depends on that row only (no BLAS blocking)."""

    def __init__(self, X_shard, metric, id_base, flag_query=None):
        self.X = np.asarray(X_shard, dtype=np.float32)
        self.metric, self.id_base = metric, id_base
        self.flag_query = flag_query          # this shard reports that query as uncertified
        self.device = torch.device("cpu")

    def _scores(self, Q):
        Q64, X64 = np.asarray(Q, np.float32).astype(np.float64), self.X.astype(np.float64)
        out = np.empty((Q64.shape[0], X64.shape[0]))
        for i, q in enumerate(Q64):
            if self.metric == osr.METRIC_L2:
                out[i] = ((X64 - q) ** 2).sum(axis=1)
            else:
                ip = (X64 * q).sum(axis=1)
                if self.metric == osr.METRIC_COSINE:
                    den = (q * q).sum() * (X64 * X64).sum(axis=1)
                    ip = np.where(den > 0, ip / np.sqrt(np.where(den > 0, den, 1.0)), 0.0)
                out[i] = ip
        return out

    def search(self, Q, k):
        s, i = osr.topk_from_scores(self._scores(Q.numpy()), k, self.metric, self.id_base)
        return torch.from_numpy(np.ascontiguousarray(s)), torch.from_numpy(np.ascontiguousarray(i))

    def merge(self, scores, ids, k_out):
        s, i = osr.merge_topk(scores.numpy(), ids.numpy(), k_out, self.metric)
        return torch.from_numpy(np.ascontiguousarray(s)), torch.from_numpy(np.ascontiguousarray(i))

    def fidelity_rows(self, Q, local_idx):
        idx = local_idx.numpy()
        rows = self.X[np.maximum(idx, 0)] if self.X.shape[0] else np.zeros(idx.shape + (Q.shape[1],), np.float32)
        f = oq.amplitude_fidelity_batch(Q.numpy(), rows)
        return torch.from_numpy(np.where(idx < 0, -np.inf, f))

    def sort_scores(self, scores, k):
        order = oq.rank_rows(scores.numpy(), k)
        return torch.from_numpy(order.astype(np.int32)), torch.from_numpy(np.take_along_axis(scores.numpy(), order, 1))

    # ---- the packed form: same protocol as CudaEngine (thresholds from the union of the cut lists) ----
    def _keys(self, s, i):
        key = s if self.metric == osr.METRIC_L2 else -s
        return torch.where(i >= 0, key, torch.full_like(key, float("inf")))

    def packed_begin(self, Q, k, shards):
        s, i = self.search(Q, k)
        self._pk = (Q, k, shards, s, i, self._keys(s, i))
        return self._pk[5][:, :exchange_len(k, shards)].contiguous() if shards > 1 else None

    def packed_filter(self, bm_all):
        """The threshold (k-th best key of the union of the cut lists: valid, exact unless one shard holds > kt) is
        known after the all-gather; what gets summed over the shards is how many entries each holds at or above it."""
        Q, k, shards, s, i, key = self._pk
        nq = key.shape[0]
        if shards == 1:
            return None
        assert tuple(bm_all.shape) == (shards, nq, exchange_len(k, shards))
        self._kth = torch.sort(bm_all.permute(1, 0, 2).reshape(nq, -1), dim=1).values[:, k - 1:k]
        return ((key <= self._kth) & (i >= 0)).sum(dim=1, keepdim=True).to(torch.int32)

    def packed_finish(self, hist_all, kk, pack):
        Q, k, shards, s, i, key = self._pk
        nq = key.shape[0]
        keep = i >= 0
        if shards > 1:
            keep = keep & (key <= self._kth)
            # the all-reduce really summed over the shards: a finite threshold has at least k entries at or above it
            assert hist_all.shape == (nq, 1) and bool((hist_all[:, 0] >= k * torch.isfinite(self._kth[:, 0])).all())
        pad_s = float("inf") if self.metric == osr.METRIC_L2 else float("-inf")
        s = torch.where(keep, s, torch.full_like(s, pad_s))
        i = torch.where(keep, i, torch.full_like(i, -1))
        f = self.fidelity_rows(Q, torch.where(keep, i - self.id_base, torch.full_like(i, -1)))
        bad = np.zeros(nq, dtype=np.int64)
        if self.flag_query is not None:
            bad[self.flag_query] = 1
        pack[:nq] = torch.from_numpy(osr.pack_records(s.numpy(), i.numpy(), f.numpy(), kk, self.metric, bad))
        return pack

    def owner_finalize(self, recv, kk, k1, k2, q_base, nq, out):
        out.copy_(torch.from_numpy(osr.owner_finalize(recv.numpy(), kk, k1, k2, self.metric, q_base, nq)))
        return out


def _data(n, d, nq, seed=0):
    rng = np.random.RandomState(seed)
    X = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((nq, d)).astype(np.float32)
    X[n - 2] = X[1]                 # duplicates in different shards: exact ties, global id decides
    X[n // 2] = X[1]
    Q[0] = X[1]
    return X, Q


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, nq, k1, k2, metric, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, Q = _data(n, d, nq)
        lo, hi = shard_bounds(n, world, rank)
        eng = OracleEngine(X[lo:hi], metric, lo)
        path = ShardedSearchRerank(torch.from_numpy(X[lo:hi]), n, metric, engine=eng)
        own = path(torch.from_numpy(Q), k1, k2)                                   # query-partitioned form (all-to-all)
        assert own.search_ids is None
        res = path(torch.from_numpy(Q), k1, k2, return_search_lists=True)         # all-gather form
        out[rank] = (own.scores.numpy(), own.ids.numpy(), res.scores.numpy(), res.ids.numpy(),
                     res.search_scores.numpy(), res.search_ids.numpy())
    finally:
        dist.destroy_process_group()


def _single(n, d, nq, k1, k2, metric):
    X, Q = _data(n, d, nq)
    eng = OracleEngine(X, metric, 0)
    path = ShardedSearchRerank(torch.from_numpy(X), n, metric, engine=eng)
    own = path(torch.from_numpy(Q), k1, k2)                                       # packed form, one shard
    assert own.search_ids is None
    res = path(torch.from_numpy(Q), k1, k2, return_search_lists=True)
    return (own.scores.numpy(), own.ids.numpy(), res.scores.numpy(), res.ids.numpy(),
            res.search_scores.numpy(), res.search_ids.numpy())


def test_shard_bounds_partition():
    for n in (0, 1, 7, 10, 1000003):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world,metric,n,k1,k2", [(2, osr.METRIC_COSINE, 301, 40, 7), (3, osr.METRIC_L2, 100, 50, 10),
                                                  (2, osr.METRIC_IP, 9, 16, 16), (4, osr.METRIC_COSINE, 257, 33, 5)])
def test_sharded_equals_single_rank(world, metric, n, k1, k2):
    d, nq = 24, 5
    want = _single(n, d, nq, k1, k2, metric)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, d, nq, k1, k2, metric, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        got = out[rank]
        for a, b in zip(got, want):
            assert np.array_equal(a, b), f"rank {rank} differs from the single-rank result"
    # and against the plain oracle search on the whole corpus
    X, Q = _data(n, d, nq)
    rs, ri = osr.topk_from_scores(OracleEngine(X, metric, 0)._scores(Q), k1, metric)
    assert np.array_equal(want[5], ri)


class PhasedOracleEngine(OracleEngine):
    """OracleEngine with the phase-split search of the CUDA engine in the all-gather form too: a shard returns only
    ITS members of the global top-k (padded with id -1) plus a per-query status, so the status reduce and the exact
    rerun of ShardedSearchRerank.search run on CPU.  ``flag_query`` makes one rank report that query as uncertified."""

    def search_exact(self, Q, k):
        return self.search(Q, k)

    def search_sharded(self, Q, k, all_gather, all_reduce_sum, shards):
        s, i = self.search(Q, k)                                              # this shard's best k
        key = s if self.metric == osr.METRIC_L2 else -s
        key = torch.where(i >= 0, key, torch.full_like(key, float("inf")))
        allk = all_gather(key)                                                # [G, nq, k]
        kth = torch.sort(allk.permute(1, 0, 2).reshape(key.shape[0], -1), dim=1).values[:, k - 1:k]
        keep = (key <= kth) & (i >= 0)                                        # ties at the k-th key are kept: a superset
        pad_s = float("inf") if self.metric == osr.METRIC_L2 else float("-inf")
        s = torch.where(keep, s, torch.full_like(s, pad_s))
        i = torch.where(keep, i, torch.full_like(i, -1))
        status = torch.zeros(key.shape[0], dtype=torch.int32)
        if self.flag_query is not None:
            status[self.flag_query] = 1
            s[self.flag_query] = pad_s                                        # an uncertified list may hold anything
            i[self.flag_query] = -1
        return s, i, status


def _phased_worker(rank, world, port, n, d, nq, k1, k2, metric, skew, flag, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, Q = _data(n, d, nq)
        if skew:                                     # every neighbour of query 2 lives in shard 0: its list overflows the cut
            X[:k1 + 5] = Q[2] + 1e-3 * np.random.RandomState(5).standard_normal((k1 + 5, d)).astype(np.float32)
        lo, hi = shard_bounds(n, world, rank)
        eng = PhasedOracleEngine(X[lo:hi], metric, lo, flag_query=(1 if flag and rank == world - 1 else None))
        path = ShardedSearchRerank(torch.from_numpy(X[lo:hi]), n, metric, engine=eng)
        res = path(torch.from_numpy(Q), k1, k2)
        out[rank] = (res.scores.numpy(), res.ids.numpy(), res.search_ids is not None)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("skew,flag", [(False, False), (True, False), (False, True)])
def test_phased_search_cut_and_fallback_routes(skew, flag):
    """(no skew, no flag) takes the query-partitioned form with cut lists; a skewed corpus overflows the cut and a
    flagged query fails the certificate: both must fall back to the all-gather form and still give the exact answer."""
    world, metric, n, d, nq, k1, k2 = 4, osr.METRIC_COSINE, 2000, 16, 6, 300, 9
    X, Q = _data(n, d, nq)
    if skew:
        X[:k1 + 5] = Q[2] + 1e-3 * np.random.RandomState(5).standard_normal((k1 + 5, d)).astype(np.float32)
    eng = OracleEngine(X, metric, 0)
    want = ShardedSearchRerank(torch.from_numpy(X), n, metric, engine=eng)(torch.from_numpy(Q), k1, k2)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_phased_worker, args=(world, _free_port(), n, d, nq, k1, k2, metric, skew, flag, out), nprocs=world, join=True)
    for rank in range(world):
        scores, ids, fell_back = out[rank]
        assert np.array_equal(ids, want.ids.numpy()) and np.array_equal(scores, want.scores.numpy()), rank
        assert fell_back == (skew or flag)            # the all-gather form materialises the merged search lists


def _pipelined_worker(rank, world, port, n, d, nq, k1, k2, metric, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, _ = _data(n, d, nq)
        lo, hi = shard_bounds(n, world, rank)
        rng = np.random.RandomState(77)
        batches = [torch.from_numpy(rng.standard_normal((nq, d)).astype(np.float32)) for _ in range(4)]
        # batch 2 is flagged by the last rank (its certificate fails): it alone reruns through the all-gather form,
        # in the middle of the queue, identically on every rank
        engines = [PhasedOracleEngine(X[lo:hi], metric, lo, flag_query=(0 if b == 2 and rank == world - 1 else None))
                   for b in range(4)]
        path = ShardedSearchRerank(torch.from_numpy(X[lo:hi]), n, metric, engine=engines[0])
        pend = []
        for b, Q in enumerate(batches):
            path.engine = engines[b]
            pend.append(path.submit(Q, k1, k2))
        res = []
        for b, p in enumerate(pend):
            path.engine = engines[b]
            res.append(p.result())
        out[rank] = [(r.scores.numpy(), r.ids.numpy(), r.search_ids is not None) for r in res]
    finally:
        dist.destroy_process_group()


def test_submit_keeps_several_batches_in_flight_and_reruns_only_the_flagged_one():
    """``submit`` queues batches back to back (shared exchange buffers, results copied out per batch); ``result()`` reads
    each batch's certificate later.  Four different batches in flight over 3 gloo ranks, the third one flagged."""
    world, metric, n, d, nq, k1, k2 = 3, osr.METRIC_L2, 600, 12, 5, 60, 8
    X, _ = _data(n, d, nq)
    rng = np.random.RandomState(77)
    batches = [rng.standard_normal((nq, d)).astype(np.float32) for _ in range(4)]
    eng = OracleEngine(X, metric, 0)
    single = ShardedSearchRerank(torch.from_numpy(X), n, metric, engine=eng)
    want = [single(torch.from_numpy(Q), k1, k2) for Q in batches]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_pipelined_worker, args=(world, _free_port(), n, d, nq, k1, k2, metric, out), nprocs=world, join=True)
    for rank in range(world):
        for b, (scores, ids, reran) in enumerate(out[rank]):
            assert np.array_equal(ids, want[b].ids.numpy()) and np.array_equal(scores, want[b].scores.numpy()), (rank, b)
            assert reran == (b == 2)


class LanedOracleEngine(PhasedOracleEngine):
    """OracleEngine with ``lane()`` views (separate per-batch state), which switches ShardedSearchRerank to two lanes:
    alternate batches go through alternate process groups, workspaces and exchange buffers."""

    def lane(self, i):
        views = self.__dict__.setdefault("_views", {})
        if i not in views:
            import copy
            v = copy.copy(self)
            v.lane_id = i
            views[i] = v
        return views[i]


def _laned_worker(rank, world, port, n, d, nq, k1, k2, metric, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X, _ = _data(n, d, nq)
        lo, hi = shard_bounds(n, world, rank)
        rng = np.random.RandomState(91)
        batches = [torch.from_numpy(rng.standard_normal((nq, d)).astype(np.float32)) for _ in range(5)]
        path = ShardedSearchRerank(torch.from_numpy(X[lo:hi]), n, metric, engine=LanedOracleEngine(X[lo:hi], metric, lo))
        assert path.n_lanes == 2 and path._lane_groups[1] is not None and path._lane_groups[1] is not path._lane_groups[0]
        pend = [path.submit(Q, k1, k2) for Q in batches]
        used = sorted(v.lane_id for v in path.engine._views.values())
        res = [p.result() for p in pend]
        again = path(batches[1], k1, k2)                     # one at a time after the queue drained
        nbufs = [len(b) for b in path._lane_bufs]
        with pytest.raises(ValueError):
            path.use_lanes(3)                                # only two lanes were created
        path.use_lanes(1)                                    # the bench's A/B: the same queue through lane 0 alone
        one = [path.submit(Q, k1, k2) for Q in batches[:2]]
        assert path._next_lane == 0
        for b, p_ in enumerate(one):
            r = p_.result()
            assert torch.equal(r.ids, res[b].ids) and torch.equal(r.scores, res[b].scores)
        path.use_lanes(2)
        path.close()                                         # drops the second lane's group: one lane from here on
        assert path.n_lanes == 1 and len(path._lane_groups) == 1
        closed = path(batches[2], k1, k2)
        out[rank] = ([(r.scores.numpy(), r.ids.numpy()) for r in res + [again, closed]], used, nbufs)
    finally:
        dist.destroy_process_group()


def test_two_lanes_alternate_process_groups_and_give_the_single_rank_result():
    world, metric, n, d, nq, k1, k2 = 3, osr.METRIC_COSINE, 500, 10, 6, 40, 7
    X, _ = _data(n, d, nq)
    rng = np.random.RandomState(91)
    batches = [rng.standard_normal((nq, d)).astype(np.float32) for _ in range(5)]
    single = ShardedSearchRerank(torch.from_numpy(X), n, metric, engine=OracleEngine(X, metric, 0))
    assert single.n_lanes == 1
    want = [single(torch.from_numpy(Q), k1, k2) for Q in batches]
    want.append(want[1])
    want.append(want[2])
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_laned_worker, args=(world, _free_port(), n, d, nq, k1, k2, metric, out), nprocs=world, join=True)
    for rank in range(world):
        got, used, nbufs = out[rank]
        assert used == [0, 1] and nbufs == [1, 1]            # both lanes ran, each with its own exchange buffers
        for b, (scores, ids) in enumerate(got):
            assert np.array_equal(ids, want[b].ids.numpy()) and np.array_equal(scores, want[b].scores.numpy()), (rank, b)
