#!/usr/bin/env python
"""Multi-GPU check + timing of the sharded search -> NCCL merge -> quantum rerank path (config 4).

    torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        tools/sharded_check.py [--N 2000000 --nq 1024 --k1 1000 --k2 10]

Every rank builds the same synthetic corpus shard by shard (seeded per 1M-row block, so the corpus
is identical for every G), runs the path, and rank 0 prints the time per batch (CUDA events, max over
ranks) and a checksum of the result that must not depend on G.
"""
import argparse
import hashlib
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def corpus_rows(lo, hi, D, device):
    """Rows [lo, hi) of the synthetic corpus: block b (2^18 rows) is randn(seed 1238 + b), normalised."""
    B = 1 << 18
    out = torch.empty((hi - lo, D), dtype=torch.float32, device=device)
    b = lo // B
    while b * B < hi:
        g = torch.Generator(device=device).manual_seed(1238 + b)
        blk = torch.nn.functional.normalize(torch.randn(B, D, generator=g, device=device), dim=1)
        a0, a1 = max(lo, b * B), min(hi, (b + 1) * B)
        out[a0 - lo:a1 - lo] = blk[a0 - b * B:a1 - b * B]
        b += 1
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=2_000_000)
    ap.add_argument("--D", type=int, default=384)
    ap.add_argument("--nq", type=int, default=1024)
    ap.add_argument("--k1", type=int, default=1000)
    ap.add_argument("--k2", type=int, default=10)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--graph", action="store_true", help="also time the CUDA-graph replay of the batch")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from quantum_rag_b200 import _lib
    from quantum_rag_b200.sharded import ShardedSearchRerank, shard_bounds
    if rank == 0:
        _lib.build()
    if world > 1:
        dist.barrier()
    dev = torch.device("cuda", local)
    lo, hi = shard_bounds(a.N, world, rank)
    X = corpus_rows(lo, hi, a.D, dev)
    g = torch.Generator(device=dev).manual_seed(2238)
    Q = torch.nn.functional.normalize(torch.randn(a.nq, a.D, generator=g, device=dev), dim=1)
    path = ShardedSearchRerank(X, a.N, "cosine")
    for _ in range(2):
        res = path(Q, a.k1, a.k2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    def timed(fn):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), out

    # (1) one batch at a time: the host waits for every batch's certificate before it queues the next
    ms_sync, res = timed(lambda: [path(Q, a.k1, a.k2) for _ in range(a.steps)][-1])
    # (2) a serving loop: batches queued back to back, certificates read afterwards (all inside the timed region)
    ms_pipe, res_p = timed(lambda: [p.result() for p in [path.submit(Q, a.k1, a.k2) for _ in range(a.steps)]][-1])
    lanes = path.n_lanes
    queue = lambda: [p.result() for p in [path.submit(Q, a.k1, a.k2) for _ in range(a.steps)]][-1]
    two, one = [ms_pipe], []
    for _ in range(3):                                     # A/B, alternating: the same queue through one lane / two lanes
        path.use_lanes(1)
        t, res_1 = timed(queue)
        one.append(t)
        path.use_lanes(lanes)
        two.append(timed(queue)[0])
    ms_one, ms_pipe2 = sorted(one)[1], sorted(two)[len(two) // 2]
    modes = {"sync_per_batch_ms": ms_sync, "pipelined_ms": ms_pipe, "pipelined_again_ms": ms_pipe2, "lanes": lanes,
             "pipelined_single_lane_ms": ms_one,
             "runs_two_lanes": [round(x, 4) for x in two], "runs_single_lane": [round(x, 4) for x in one],
             "pipelined_equal": bool(torch.equal(res_p.ids, res.ids) and torch.equal(res_p.scores, res.scores) and
                                     torch.equal(res_1.ids, res.ids) and torch.equal(res_1.scores, res.scores))}
    if a.graph:
        for _ in range(2):
            path.submit(Q, a.k1, a.k2, graph=True).result()
        ms_g, res_g = timed(lambda: [path.submit(Q, a.k1, a.k2, graph=True).result() for _ in range(a.steps)][-1])
        ms_gp, _ = timed(lambda: [p.result() for p in [path.submit(Q, a.k1, a.k2, graph=True) for _ in range(a.steps)]][-1])
        modes.update({"graph_sync_per_batch_ms": ms_g, "graph_pipelined_ms": ms_gp,
                      "graph_equal": bool(torch.equal(res_g.ids, res.ids) and torch.equal(res_g.scores, res.scores))})
    path.profile = {}
    res = path(Q, a.k1, a.k2)
    prof = {k: round(v, 3) for k, v in path.profile.items()}
    path.profile = None
    h = hashlib.sha256()
    for x in (res.ids, res.scores):
        h.update(x.cpu().numpy().tobytes())
    # the packed NCCL route against the plain one (exact CUDA-core search, all-gather, merge, stand-alone fidelity)
    sub = torch.arange(0, a.nq, max(1, a.nq // 8), device=dev)[:8]
    ref = path.exact_reference(Q[sub], a.k1, a.k2)
    same = bool(torch.equal(ref.ids, res.ids[sub]) and torch.equal(ref.scores, res.scores[sub]))
    # ... and the all-gather form (the rerun route: thresholds exchanged, per-shard lists all-gathered, merged on every rank)
    full = path(Q, a.k1, a.k2, return_search_lists=True)
    same_full = bool(torch.equal(full.ids, res.ids) and torch.equal(full.scores, res.scores))
    sub_ok = bool(torch.equal(full.search_ids[sub], ref.search_ids) and torch.equal(full.search_scores[sub], ref.search_scores))
    if rank == 0:
        ms = ms_sync
        print(json.dumps({"gpus": world, "modes": modes, "N": a.N, "nq": a.nq, "k1": a.k1, "k2": a.k2, "ms_per_batch": ms,
                          "search_scores_per_s": a.nq * a.N / ms * 1e3, "reranked_queries_per_s": a.nq / ms * 1e3,
                          "rerun_all_gather_form": path.last_rerun, "equals_exact_route_on_8_queries": same,
                          "all_gather_form_equal": same_full, "all_gather_form_lists_equal_exact_search": sub_ok,
                          "stage_ms_rank0": prof, "result_sha256": h.hexdigest()}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
