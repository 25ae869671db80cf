"""Row-sharded search + quantum rerank over ``torch.distributed`` (one process per GPU).

The reference is a single CPU process with no retrieval step at all;
this is the multi-GPU form BASELINE.json's config 4 asks for:

    corpus rows   contiguous row partition, rank g owns rows [lo_g, hi_g); global id = lo_g + row
    queries       replicated on every rank
    search        per shard: tcgen05 filter GEMM + exact rescoring (FlatIndexTC); the filter
                  thresholds are global (an all-gather of [nq, kt] fp32 bucket maxima, kt ~ 2 k1 / G,
                  and an all-reduce of a [nq, 256] int32 survivor histogram between the phases), so a
                  shard rescores only its ~1/G share of the global top-k1
    rerank        fused into the rescoring: the kernel that reads a candidate row for its exact
                  search score also emits its amplitude-encoded fidelity (same bits as
                  qrag_amp_fidelity), so no row is read twice and no embedding crosses NVLink
    exchange      ONE all-to-all: the shard's sorted list of query q -- one record of
                  (score, id, fidelity) x kk, written by the sort kernel in the send layout --
                  goes to the rank that owns q (queries are partitioned over the ranks)
    owner         one kernel per rank: merge the G lists by rank (binary searches, nothing moves),
                  keep the global top-k1, order by (fidelity desc, position asc), keep k2
    result        one small all-gather of [nq / G, k2] (fidelity, id, status) -> every rank

Four collectives per batch, three of them latency-sized; no torch arithmetic between the kernels.
``return_search_lists=True`` runs the all-gather form instead (every rank materialises the merged
[nq, k1] search lists): per-shard lists all-gathered, merged on every rank, owner-computes rerank,
all-reduce(MAX).  It is also the rerun route when some shard could not certify a query.

Every per-row number (search score, fidelity) is computed by the same device code from the same fp32
row whatever the sharding, and the merges are total orders, so the result for G ranks is
bit-identical to the result for 1 rank (tests/test_sharded_gloo.py, tests/test_gpu_sharded.py).

The compute is behind a small engine object so that the exchange logic can be exercised on CPU
with the ``gloo`` backend (the tests plug the NumPy oracle in); the default engine is the CUDA one
and it raises without a GPU -- there is no CPU fallback in the product path.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def exchange_len(k: int, world: int) -> int:
    """Entries of the threshold lists and of the packed records that travel between the shards: k for one
    shard, else min(k, 2k/G + 64 rounded up to 32) (= qrag_search_tc_exchange_len; tests/test_host_logic.py)."""
    if world <= 1:
        return int(k)
    return int(min(k, -(-(2 * k // world + 64) // 32) * 32))


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
        self._views = [self]

    def lane(self, i: int) -> "CudaEngine":
        """View number ``i`` of this engine: the same rows and 16-bit shadow, its own search workspace and phase
        state, so that two batches can be between their exchange points at the same time (view 0 is the engine)."""
        while len(self._views) <= i:
            v = copy.copy(self)
            v.index = self.index.fork()
            self._views.append(v)
        return self._views[i]

    def search(self, Q, k):
        return self.index.search(Q, k)

    def sync_corpus_bound(self, all_reduce_max):
        """The filter's error bound uses max |x| and the largest 16-bit rounding-error norm over the WHOLE corpus
        (aux[0], aux[1]): reduce them once per index."""
        all_reduce_max(self.index.aux[:2])

    def search_sharded(self, Q, k, all_gather, all_reduce_sum, shards):
        """This shard's members of the global top-k (thresholds exchanged through ``all_gather``)."""
        return self.index.search_sharded(Q, k, all_gather, all_reduce_sum, shards)

    # ---- packed search + rerank (include/qrag.h: qrag_search_tc_finish_packed / qrag_owner_finalize) ----
    def packed_begin(self, Q, k, shards):
        return self.index.tc_begin(Q, k, shards)

    def packed_filter(self, bm_all):
        return self.index.tc_filter(bm_all)

    def packed_finish(self, hist_all, kk, pack):
        return self.index.tc_finish_packed(hist_all, kk, pack)

    def owner_finalize(self, recv, kk, k1, k2, q_base, nq, out):
        return self.api.owner_finalize(recv, kk, k1, k2, self.metric, q_base, nq, out)

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


class PendingResult:
    """A batch queued by ``ShardedSearchRerank.submit``.  ``result()`` waits for it, reads the certificate flag (the one
    host synchronisation of the path) and, if some shard could not certify a query or had to cut a list, reruns the
    batch through the all-gather form -- identically on every rank, because the flag was all-gathered with the result."""

    def __init__(self, path, Q, k1, k2, top, ids, flag, event=None):
        self.path, self.Q, self.k1, self.k2 = path, Q, k1, k2
        self.top, self.ids, self.flag = top, ids, flag
        self.event = event            # end of the batch on its lane's stream (None: queued on the caller's stream)
        self._done = None

    def result(self) -> ShardedResult:
        if self._done is None:
            if self.event is not None:
                cur = torch.cuda.current_stream()
                cur.wait_event(self.event)
                for t in (self.top, self.ids, self.flag):   # allocated on the lane's stream, read on the caller's
                    t.record_stream(cur)
            bad = int(self.flag.item())
            self.path._flush_marks()
            if bad:
                self.path.last_rerun = 1
                self.path._marks = []
                ss, si = self.path.search(self.Q, self.k1)
                top, ids = self.path.rerank(self.Q, si, self.k2)
                self._done = ShardedResult(top, ids, ss, si)
            else:
                self._done = ShardedResult(self.top, self.ids, None, None)
        return self._done


class ShardedSearchRerank:
    """Search the row-sharded corpus, merge over the process group, quantum-rerank the merged list."""

    def __init__(self, X_shard, n_total: int, metric: str = "cosine", group: Optional[dist.ProcessGroup] = None,
                 engine=None, lanes: Optional[int] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = int(n_total)
        if self.n_total < self.world:
            # every rank computes the same bounds, so every rank raises: no rank is left waiting in a collective
            raise ValueError(f"{self.n_total} rows cannot be sharded over {self.world} ranks (an empty shard)")
        self.lo, self.hi = shard_bounds(self.n_total, self.world, self.rank)
        if X_shard.shape[0] != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} owns rows [{self.lo}, {self.hi}) but got {X_shard.shape[0]} rows")
        self.metric = metric
        self.profile = None          # set to {} to collect per-stage CUDA-event timings (ms) of the next call
        self._marks = []
        self._graphs = {}
        self.last_rerun = 0          # 1 if the last call had to rerun through the all-gather form
        self.engine = engine if engine is not None else CudaEngine(X_shard, metric, self.lo)
        if hasattr(self.engine, "sync_corpus_bound"):
            self.engine.sync_corpus_bound(self._all_reduce_max)
        # Lanes: ``submit`` alternates between `lanes` independent (stream, process group, workspace, exchange buffers)
        # sets, so that one batch's collectives and per-query kernels run beside the next batch's filter GEMM instead
        # of in front of it.  Each lane has its own communicator: NCCL runs a communicator's collectives in issue
        # order on one internal stream, and a single one would chain the two batches together again.  Every rank
        # creates the groups here, in the same order (new_group is collective).
        can = hasattr(self.engine, "lane")
        self.n_lanes = max(1, int(lanes)) if lanes is not None else (2 if can else 1)
        if self.n_lanes > 1 and not can:
            raise ValueError("this engine has no lane() views: lanes must be 1")
        self._lane_groups = [group]
        for _ in range(1, self.n_lanes):
            g = None
            if self.world > 1:
                g = dist.new_group(ranks=dist.get_process_group_ranks(group) if group is not None else None)
            self._lane_groups.append(g)
        dev = getattr(self.engine, "device", None)
        on_gpu = dev is not None and torch.device(dev).type == "cuda"
        self._lane_streams = [torch.cuda.Stream(device=dev) if (on_gpu and self.n_lanes > 1) else None
                              for _ in range(self.n_lanes)]
        self._lane_bufs = [dict() for _ in range(self.n_lanes)]
        self._next_lane = 0

    def close(self) -> None:
        """Drop the captured CUDA graphs (``submit(graph=True)``) and the extra lanes' communicators; the object then
        runs on one lane.  Call it before ``destroy_process_group``: a graph that holds NCCL kernels must not outlive
        the communicator (observed: the teardown hangs otherwise)."""
        self._graphs = {}
        for side in getattr(self, "_lane_streams", []):
            if side is not None:
                side.synchronize()
        groups, self._lane_groups = self._lane_groups, self._lane_groups[:1]
        self.n_lanes, self._next_lane = 1, 0
        for g in groups[1:]:                                  # the extra lanes' communicators (collective: every rank closes)
            if g is not None:
                try:
                    dist.destroy_process_group(g)
                except Exception:
                    pass

    def _join_lanes(self) -> None:
        """Order the caller's stream after everything queued on the lanes.  The all-gather form (``search``: the rerun
        route, ``return_search_lists=True``) runs on the caller's stream with lane 0's search workspace, which a batch
        still in flight on lane 0's own stream may be using."""
        for side in self._lane_streams:
            if side is not None:
                torch.cuda.current_stream().wait_stream(side)

    def use_lanes(self, n: int) -> None:
        """Run the following ``submit`` calls over the first ``n`` lanes (1 = every batch behind the previous one on one
        stream and one communicator; the bench's A/B).  Every rank must make the same call, between batches."""
        if not 1 <= int(n) <= len(self._lane_groups):
            raise ValueError(f"lanes must be in [1, {len(self._lane_groups)}]")
        self.n_lanes, self._next_lane = int(n), 0

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
    def _all_gather(self, t: torch.Tensor, group=None) -> torch.Tensor:
        if self.world == 1:
            return t[None]
        t = t.contiguous()
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group if group is None else group)  # concatenation along dim 0
        return out.view((self.world,) + tuple(t.shape))

    def _all_reduce_sum(self, t: torch.Tensor, group=None) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group if group is None else group)
        return t

    def _all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    # ------------------------------------------------------------------------- the path
    def search(self, Q, k1: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Global top-k1 (identical on every rank): per-shard search, all-gather of the lists, merge.

        With the CUDA engine and G > 1 the shards also exchange their filter thresholds (an all-gather and
        an all-reduce inside the search), so that each shard rescores only its ~1/G share of the
        global list; any query a shard could not certify makes every rank rerun it exactly.
        """
        self._join_lanes()
        if self.world > 1 and hasattr(self.engine, "search_sharded"):
            s, i, status = self.engine.search_sharded(Q, k1, self._all_gather, self._all_reduce_sum, self.world)
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

    def exact_reference(self, Q, k1: int, k2: int) -> ShardedResult:
        """The same answer by the plain route, for checking the packed path inside a run (every rank must call it):
        exact CUDA-core search of each shard (no tensor cores, no thresholds), all-gather of the lists, merge on
        every rank, owner-computes fidelity with the stand-alone kernel, all-reduce(MAX), stable sort."""
        s, i = self.engine.search_exact(Q, k1)
        gathered = self._all_gather(torch.stack([s.view(torch.int64), i], dim=0))
        ss, si = self.engine.merge(gathered[:, 0].contiguous().view(torch.float64), gathered[:, 1].contiguous(), k1)
        top, ids = self.rerank(Q, si, min(k2, k1))
        return ShardedResult(top, ids, ss, si)

    def _buffers(self, lane: int, nq: int, kk: int, k2: int, device):
        """Exchange buffers of the packed path, one set per lane, allocated once per shape (send rows beyond nq stay
        zero: empty records)."""
        key = (nq, kk, k2, str(device))
        b = self._lane_bufs[lane].get(key)
        if b is None:
            per = -(-nq // self.world)
            rec, orec = 3 * kk + 1, 2 * k2 + 1
            b = {"per": per,
                 "send": torch.zeros((per * self.world, rec), dtype=torch.int64, device=device),
                 "recv": torch.empty((self.world, per, rec), dtype=torch.int64, device=device),
                 "out": torch.empty((per, orec), dtype=torch.int64, device=device),
                 "res": torch.empty((self.world, per, orec), dtype=torch.int64, device=device)}
            self._lane_bufs[lane] = {key: b}
        return b

    def _enqueue_packed(self, Q, k1: int, k2: int, lane: int = 0):
        """Queue one batch of the packed path on the current stream (no host synchronisation), with lane ``lane``'s
        workspace, exchange buffers and process group:
        search phases with their two threshold exchanges, ONE all-to-all of per-query records to the query's owner,
        one owner kernel, one small all-gather.  Returns (fidelity [nq, k2], ids [nq, k2], flag [1]) device tensors;
        flag != 0 means some shard could not certify a query or had to cut a list."""
        G = self.world
        eng = self.engine.lane(lane) if hasattr(self.engine, "lane") else self.engine
        grp = self._lane_groups[lane]
        nq = Q.shape[0]
        kk = exchange_len(k1, G)
        bm = eng.packed_begin(Q, k1, G)
        bm_all = self._all_gather(bm, grp) if G > 1 else None
        self._mark("begin+gather")
        hist = eng.packed_filter(bm_all)
        if G > 1:
            self._all_reduce_sum(hist, grp)
        self._mark("filter+reduce")
        dev = getattr(eng, "device", Q.device)
        buf = self._buffers(lane, nq, kk, k2, dev)
        per = buf["per"]
        eng.packed_finish(hist, kk, buf["send"])
        self._mark("finish_packed")
        if G > 1:
            dist.all_to_all_single(buf["recv"].view(G * per, -1), buf["send"], group=grp)
            recv = buf["recv"]
        else:
            recv = buf["send"].view(1, per, -1)
        self._mark("all_to_all")
        out = eng.owner_finalize(recv, kk, k1, k2, self.rank * per, nq, buf["out"])
        self._mark("owner_finalize")
        if G > 1:
            dist.all_gather_into_tensor(buf["res"].view(G * per, -1), out, group=grp)
            res = buf["res"].view(G * per, -1)[:nq]
        else:
            res = out[:nq]
        top = res[:, :k2].contiguous().view(torch.float64)         # copies: the exchange buffers are reused by the next batch
        ids = res[:, k2:2 * k2].contiguous()
        flag = res[:, 2 * k2].max().reshape(1) if nq else torch.zeros(1, dtype=torch.int64, device=dev)
        self._mark("result_gather")
        return top, ids, flag

    def submit(self, Q, k1: int = 1000, k2: int = 10, graph: bool = False) -> "PendingResult":
        """Queue one batch of the packed path and return at once; ``PendingResult.result()`` is where the host waits.
        Batches submitted back to back keep the GPU busy while the host queues the next one (a serving loop).
        ``graph=True`` replays the batch from a CUDA graph captured on the first call with this shape (collectives
        included), so the host cost per batch is one graph launch.  Every rank must make the same calls."""
        if self.n_total >= (1 << 40):
            raise ValueError("corpus too large: ids must stay below 2**40")
        if not isinstance(Q, torch.Tensor):
            Q = torch.as_tensor(Q)
        k2 = min(k2, k1)
        if not graph:
            lane = self._next_lane
            self._next_lane = (lane + 1) % self.n_lanes
            side = self._lane_streams[lane]
            if side is None:
                self._mark("start")
                top, ids, flag = self._enqueue_packed(Q, k1, k2, lane)
                return PendingResult(self, Q, k1, k2, top, ids, flag)
            if Q.device != side.device:
                Q = Q.to(side.device, non_blocking=True)
            side.wait_stream(torch.cuda.current_stream())           # Q may have been produced on the caller's stream
            with torch.cuda.stream(side):
                self._mark("start")
                top, ids, flag = self._enqueue_packed(Q, k1, k2, lane)
                done = torch.cuda.Event()
                done.record(side)
            return PendingResult(self, Q, k1, k2, top, ids, flag, done)
        self._join_lanes()                                          # the replay shares lane 0's workspace and buffers
        key = (tuple(Q.shape), k1, k2)
        g = self._graphs.get(key)
        if g is None:
            static_q = torch.empty_like(Q, device=getattr(self.engine, "device", Q.device))
            static_q.copy_(Q)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                           # warm-up outside the capture (lazy inits, allocations)
                self._enqueue_packed(static_q, k1, k2)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg, capture_error_mode="thread_local"):   # the NCCL watchdog thread keeps polling
                outs = self._enqueue_packed(static_q, k1, k2)
            g = (cg, static_q, outs)
            self._graphs = {key: g}                                 # one captured shape at a time
        cg, static_q, (top, ids, flag) = g
        static_q.copy_(Q, non_blocking=True)
        cg.replay()
        # the graph's outputs are overwritten by the next replay: hand out copies (stream-ordered, no host sync)
        return PendingResult(self, Q, k1, k2, top.clone(), ids.clone(), flag.clone())

    def __call__(self, Q, k1: int = 1000, k2: int = 10, return_search_lists: bool = False) -> ShardedResult:
        """Top-k2 by amplitude fidelity among the global top-k1 of the search (identical on every rank).

        ``return_search_lists`` also materialises the merged [nq, k1] search lists on every rank (the
        all-gather form of the path); without it the packed, query-partitioned form runs (any G).
        """
        if self.n_total >= (1 << 40):
            raise ValueError("corpus too large: ids must stay below 2**40")
        self.last_rerun = 0
        if not return_search_lists and hasattr(self.engine, "packed_finish"):
            return self.submit(Q, k1, k2).result()
        self._mark("start")
        ss, si = self.search(Q, k1)
        top, ids = self.rerank(Q, si, min(k2, k1))
        return ShardedResult(top, ids, ss, si)


_PATHS: dict = {}


def sharded_search_rerank(q, X_shard, k1: int = 1000, k2: int = 10, group: Optional[dist.ProcessGroup] = None,
                          metric: str = "cosine", n_total: Optional[int] = None) -> ShardedResult:
    """Functional form named in SURVEY.md section 8b: ``sharded_search_rerank(q, X_shard, k1, k2, group)``.

    The prepared shard (16-bit shadow, corpus bound) is cached per (shard tensor, metric, group), so repeated calls pay
    only for the search.  ``n_total`` defaults to the sum of the shard sizes over the group (one small all-reduce on the
    first call); shards must follow ``shard_bounds`` (contiguous, balanced).
    """
    key = (X_shard.data_ptr(), tuple(X_shard.shape), metric, id(group))
    hit = _PATHS.get(key)
    # the cache entry keeps the shard tensor alive, so a matching pointer means the same storage; rows modified in
    # place after the first call need a fresh ShardedSearchRerank (the 16-bit shadow is built once)
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
