"""Row-sharded search + quantum rerank over ``torch.distributed`` (one process per GPU).

The reference is a single CPU process with no retrieval step at all;
this is the multi-GPU form BASELINE.json's config 4 asks for:

    corpus rows   contiguous row partition, rank g owns rows [lo_g, hi_g); global id = lo_g + row
    queries       replicated on every rank
    search        per shard: tcgen05 filter GEMM + exact rescoring (FlatIndexTC); the filter
                  thresholds are global (two [nq, k1] fp32 all-gathers between the phases), so a
                  shard rescores only its ~1/G share of the global top-k1
    exchange 1    ONE all-gather of the per-shard lists, [nq, k1] x (fp64 score, int64 id)
    merge         every rank merges the G lists to the global top-k1 (same kernel as the
                  single-GPU merge, so the order is the canonical (score, id) order)
    rerank        owner-computes: a rank scores only the members of the global list that live
                  in its shard (amplitude-encoded fidelity, rows gathered by TMA); no embedding
                  ever crosses NVLink
    exchange 2    ONE all-reduce(MAX) of [nq, k1] fp64 (non-owners hold -inf)
    final         stable sort by (fidelity desc, position in the merged list asc), top-k2

Every per-row number (search score, fidelity) is computed by the same kernel from the same fp32
row whatever the sharding, and the merges are total orders, so the result for G ranks is
bit-identical to the result for 1 rank (tests/test_sharded_gloo.py, tests/test_gpu_sharded.py).

The compute is behind a small engine object so that the exchange logic can be exercised on CPU
with the ``gloo`` backend (the tests plug the NumPy oracle in); the default engine is the CUDA one
and it raises without a GPU -- there is no CPU fallback in the product path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row partition: the first ``n_rows % world`` ranks get one extra row."""
    base, extra = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class CudaEngine:
    """libqrag kernels (the product path)."""

    def __init__(self, X_shard, metric: str, id_base: int):
        from . import api
        self.api = api
        self.index = api.FlatIndexTC(X_shard, metric, id_base=id_base)
        self.metric = metric
        self.device = self.index.X.device

    def search(self, Q, k):
        return self.index.search(Q, k)

    def sync_corpus_bound(self, all_reduce_max):
        """The filter's error bound uses max |x| and the largest bf16 rounding-error norm over the WHOLE corpus
        (aux[0], aux[1]): reduce them once per index."""
        all_reduce_max(self.index.aux[:2])

    def search_sharded(self, Q, k, all_gather, shards):
        """This shard's members of the global top-k (thresholds exchanged through ``all_gather``)."""
        return self.index.search_sharded(Q, k, all_gather, shards)

    def search_exact(self, Q, k):
        """Exact CUDA-core search of this shard (the rerun path for queries the filter could not certify)."""
        return self.api.search_topk(Q, self.index.X, k, self.metric, self.index.id_base)

    def merge(self, scores, ids, k_out):
        return self.api.topk_merge(scores, ids, k_out, self.metric)

    def fidelity_rows(self, Q, local_idx):
        """fp64 [nq, C] amplitude fidelity against shard rows ``local_idx`` (-1 -> -inf)."""
        return self.api.amp_fidelity(Q, X=self.index.X, idx=local_idx)

    def sort_scores(self, scores, k):
        return self.api.sort_scores(scores, k, descending=True)


@dataclass
class ShardedResult:
    scores: torch.Tensor          # [nq, k2] fp64 fidelity, best first
    ids: torch.Tensor             # [nq, k2] int64 global document ids
    search_scores: torch.Tensor   # [nq, k1] fp64 merged search scores
    search_ids: torch.Tensor      # [nq, k1] int64 merged search ids


class ShardedSearchRerank:
    """Search the row-sharded corpus, merge over the process group, quantum-rerank the merged list."""

    def __init__(self, X_shard, n_total: int, metric: str = "cosine", group: Optional[dist.ProcessGroup] = None,
                 engine=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = int(n_total)
        self.lo, self.hi = shard_bounds(self.n_total, self.world, self.rank)
        if X_shard.shape[0] != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} owns rows [{self.lo}, {self.hi}) but got {X_shard.shape[0]} rows")
        self.metric = metric
        self.profile = None          # set to {} to collect per-stage CUDA-event timings (ms) of the next call
        self._marks = []
        self.engine = engine if engine is not None else CudaEngine(X_shard, metric, self.lo)
        if hasattr(self.engine, "sync_corpus_bound"):
            self.engine.sync_corpus_bound(self._all_reduce_max)

    def _mark(self, name: str) -> None:
        if self.profile is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((name, ev))

    def _flush_marks(self) -> None:
        if self.profile is not None and self._marks:
            torch.cuda.synchronize()
            for (_, a), (name, b) in zip(self._marks[:-1], self._marks[1:]):
                self.profile[name] = self.profile.get(name, 0.0) + a.elapsed_time(b)
            self._marks = []

    # ------------------------------------------------------------------ exchange steps
    def _all_gather(self, t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return t[None]
        t = t.contiguous()
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)           # concatenation along dim 0
        return out.view((self.world,) + tuple(t.shape))

    def _all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    # ------------------------------------------------------------------------- the path
    def search(self, Q, k1: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Global top-k1 (identical on every rank): per-shard search, all-gather of the lists, merge.

        With the CUDA engine and G > 1 the shards also exchange their filter thresholds (two [nq, k1]
        fp32 all-gathers inside the search), so that each shard rescores only its ~1/G share of the
        global list; any query a shard could not certify makes every rank rerun it exactly.
        """
        if self.world > 1 and hasattr(self.engine, "search_sharded"):
            s, i, status = self.engine.search_sharded(Q, k1, self._all_gather, self.world)
            self._mark("search_phases")
            bad = self._all_reduce_max(status.clone())
            flagged = torch.nonzero(bad).flatten()
            self._mark("status_sync")
            if flagged.numel():                                   # identical on every rank after the reduce
                Qd = Q if isinstance(Q, torch.Tensor) else torch.as_tensor(Q)
                s2, i2 = self.engine.search_exact(Qd.to(s.device)[flagged], k1)
                s[flagged] = s2
                i[flagged] = i2
        else:
            s, i = self.engine.search(Q, k1)
            self._mark("search_phases")
        # one collective for both arrays: scores and ids are both 8 bytes wide
        packed = torch.stack([s.view(torch.int64), i], dim=0)                 # [2, nq, k1]
        gathered = self._all_gather(packed)                                   # [G, 2, nq, k1]
        gs = gathered[:, 0].contiguous().view(torch.float64)
        gi = gathered[:, 1].contiguous()
        self._mark("gather_lists")
        out = self.engine.merge(gs, gi, k1)
        self._mark("merge")
        return out

    def rerank(self, Q, search_ids: torch.Tensor, k2: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Amplitude-fidelity rerank of the merged list: owner computes, all-reduce(MAX), stable top-k2."""
        own = (search_ids >= self.lo) & (search_ids < self.hi)
        local = torch.where(own, search_ids - self.lo, torch.full_like(search_ids, -1))
        self._mark("ownership")
        f = self.engine.fidelity_rows(Q, local)                                # -inf where not owned / padding
        self._mark("rerank_fidelity")
        f = self._all_reduce_max(f)
        self._mark("allreduce_fidelity")
        pos, top = self.engine.sort_scores(f, k2)
        ids = torch.gather(search_ids, 1, pos.long())
        self._mark("final_sort")
        self._flush_marks()
        return top, ids

    def _all_to_all(self, t: torch.Tensor) -> torch.Tensor:
        """t [G, ...]: slice g goes to rank g; returns [G, ...] with slice g received from rank g."""
        out = torch.empty_like(t)
        dist.all_to_all_single(out, t.contiguous(), group=self.group)
        return out

    def _owner_pipeline(self, Q, k1: int, k2: int) -> Optional[ShardedResult]:
        """G > 1: everything after the shard search is partitioned BY QUERY, so it scales with 1/G too.

        Each rank scores the fidelity of its own list entries (its rows, no communication), then one
        all-to-all sends (search score, id, fidelity) of query q to the rank owning q; the owner merges the
        G lists into the global top-k1 order, ranks by (fidelity desc, position asc) and keeps k2; one small
        all-gather returns the [nq, k2] result to every rank.  Lists are cut to the longest valid prefix found
        on any rank before they travel.  Returns None if some query could not be certified (caller falls back).
        """
        dev_lists = self.engine.search_sharded(Q, k1, self._all_gather, self.world) \
            if hasattr(self.engine, "search_sharded") else self.engine.search(Q, k1) + (None,)
        s, i, status = dev_lists
        self._mark("search_phases")
        nq = s.shape[0]
        # Lists travel cut to kk entries.  A shard holds ~ k1 / G members of the global top-k1; kk leaves 2x that
        # plus 64, and the cut is VERIFIED at the end of the call (one host sync, after everything is queued): if
        # any shard had more valid entries than kk, or flagged a query, the caller reruns the all-gather form.
        kk = k1 if status is None else max(1, min(k1, -(-(2 * k1 // self.world + 64) // 32) * 32))
        valid = (i >= 0).sum(dim=1).max().to(torch.int64).reshape(1)
        flag = status.max().to(torch.int64).reshape(1) if status is not None else torch.zeros_like(valid)
        meta = self._all_reduce_max(torch.cat([flag, valid]))
        self._mark("status_reduce")
        s, i = s[:, :kk].contiguous(), i[:, :kk].contiguous()
        own = i >= 0
        f = self.engine.fidelity_rows(Q, torch.where(own, i - self.lo, torch.full_like(i, -1)))
        self._mark("rerank_fidelity")
        # pack [nq_pad, 3, kk] int64 (score bits, id, fidelity bits), queries padded to a multiple of G
        per = -(-nq // self.world)
        pack = torch.full((per * self.world, 3, kk), -1, dtype=torch.int64, device=s.device)
        pack[:nq, 0], pack[:nq, 1], pack[:nq, 2] = s.view(torch.int64), i, f.view(torch.int64)
        recv = self._all_to_all(pack.view(self.world, per, 3, kk))           # [G (source shard), per, 3, kk]
        self._mark("all_to_all")
        gs = recv[:, :, 0].contiguous().view(torch.float64)
        gi = recv[:, :, 1].contiguous()
        gf = recv[:, :, 2].contiguous().view(torch.float64)
        # carry the source slot through the merge in the low 20 bits of the tag: ids are unique, so the order
        # (score, id << 20 | slot) is the canonical (score, id) order
        slot = (torch.arange(self.world, device=s.device)[:, None, None] * kk +
                torch.arange(kk, device=s.device)[None, None, :]).expand(self.world, per, kk)
        tagged = torch.where(gi >= 0, (gi << 20) | slot, gi)
        k_m = min(k1, self.world * kk)
        ms, mt = self.engine.merge(gs, tagged, k_m)
        mi = torch.where(mt >= 0, mt >> 20, mt)
        src = torch.where(mt >= 0, mt & 0xFFFFF, torch.zeros_like(mt))
        f_all = gf.permute(1, 0, 2).reshape(per, self.world * kk)
        mf = torch.gather(f_all, 1, src)
        mf = torch.where(mt >= 0, mf, torch.full_like(mf, float("-inf")))
        self._mark("merge")
        k2 = min(k2, k_m)
        pos, top = self.engine.sort_scores(mf, k2)
        ids = torch.gather(mi, 1, pos.long())
        out = self._all_gather(torch.stack([top.view(torch.int64), ids], dim=0))          # [G, 2, per, k2]
        top_all = out[:, 0].reshape(self.world * per, k2)[:nq].contiguous().view(torch.float64)
        ids_all = out[:, 1].reshape(self.world * per, k2)[:nq].contiguous()
        self._mark("final_sort")
        flagged, max_valid = (int(v) for v in meta.cpu().tolist())          # the one host sync of the path
        self._flush_marks()
        if flagged or max_valid > kk:
            return None
        return ShardedResult(top_all, ids_all, None, None)

    def __call__(self, Q, k1: int = 1000, k2: int = 10, return_search_lists: bool = False) -> ShardedResult:
        """Top-k2 by amplitude fidelity among the global top-k1 of the search (identical on every rank).

        ``return_search_lists`` also materialises the merged [nq, k1] search lists on every rank (the
        all-gather form of the path); without it, G > 1 takes the query-partitioned form.
        """
        if self.n_total >= (1 << 40):
            raise ValueError("corpus too large: ids must stay below 2**40")
        self._mark("start")
        if self.world > 1 and not return_search_lists:
            res = self._owner_pipeline(Q, k1, k2)
            if res is not None:
                return res
            self._marks = []
            self._mark("start")
        ss, si = self.search(Q, k1)
        top, ids = self.rerank(Q, si, min(k2, k1))
        return ShardedResult(top, ids, ss, si)


_PATHS: dict = {}


def sharded_search_rerank(q, X_shard, k1: int = 1000, k2: int = 10, group: Optional[dist.ProcessGroup] = None,
                          metric: str = "cosine", n_total: Optional[int] = None) -> ShardedResult:
    """Functional form named in SURVEY.md section 8b: ``sharded_search_rerank(q, X_shard, k1, k2, group)``.

    The prepared shard (bf16 shadow, corpus bound) is cached per (shard tensor, metric, group), so repeated calls pay
    only for the search.  ``n_total`` defaults to the sum of the shard sizes over the group (one small all-reduce on the
    first call); shards must follow ``shard_bounds`` (contiguous, balanced).
    """
    key = (X_shard.data_ptr(), tuple(X_shard.shape), metric, id(group))
    hit = _PATHS.get(key)
    # the cache entry keeps the shard tensor alive, so a matching pointer means the same storage; rows modified in
    # place after the first call need a fresh ShardedSearchRerank (the bf16 shadow is built once)
    path = hit[1] if hit is not None and hit[0] is X_shard else None
    if path is None:
        if n_total is None:
            n = torch.tensor([X_shard.shape[0]], dtype=torch.int64, device=X_shard.device)
            if dist.is_initialized() and dist.get_world_size(group) > 1:
                dist.all_reduce(n, group=group)
            n_total = int(n[0])
        path = ShardedSearchRerank(X_shard, n_total, metric, group)
        _PATHS.clear()                      # one resident shard per process: do not pin stale corpora
        _PATHS[key] = (X_shard, path)
    return path(q, k1, k2)
