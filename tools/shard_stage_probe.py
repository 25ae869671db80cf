#!/usr/bin/env python
"""One-GPU probe of what ONE rank of the G-GPU sharded search + rerank (config 4) spends per stage.

A shard of N / G rows runs the packed path (quantum_rag_b200/sharded.py); the collectives are emulated on the
device (the threshold exchanges by replicating the shard's own list / histogram G times, the all-to-all by handing the
owner kernel its own records G times: same sizes and statistics as G equal shards, no NVLink time).  Reports
CUDA-event time per stage, the eager total and the total replayed from a CUDA graph (no CPU launch gaps).

    python tools/shard_stage_probe.py [--G 8 --N 10000000 --nq 1024 --k1 1000 --k2 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_rag_b200 import _lib, api  # noqa: E402
from quantum_rag_b200.sharded import exchange_len  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--G", type=int, default=8)
    ap.add_argument("--N", type=int, default=10_000_000)
    ap.add_argument("--nq", type=int, default=1024)
    ap.add_argument("--k1", type=int, default=1000)
    ap.add_argument("--k2", type=int, default=10)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--samples", default="", help="comma-separated pass-1 sample rates to compare (QRAG_TC_SAMPLE: "
                    "builds libqrag with -DQRAG_TUNING for the run and restores the product build afterwards)")
    a = ap.parse_args()
    if not a.samples:
        _lib.build()
        probe(a)
        return
    _lib.build(force=True, tuning=True)
    try:
        for smp in a.samples.split(","):
            os.environ["QRAG_TC_SAMPLE"] = smp
            probe(a, {"sample": int(smp)})
    finally:
        os.environ.pop("QRAG_TC_SAMPLE", None)
        _lib.build(force=True)


def probe(a, extra=None):
    G, nq, k1, k2 = a.G, a.nq, a.k1, a.k2
    n = a.N // G
    g = torch.Generator(device="cuda").manual_seed(1238)
    X = torch.nn.functional.normalize(torch.randn(n, 384, generator=g, device="cuda"), dim=1)
    Q = torch.nn.functional.normalize(torch.randn(nq, 384, generator=g, device="cuda"), dim=1)
    index = api.FlatIndexTC(X, "cosine")
    kk, per = exchange_len(k1, G), -(-nq // G)
    send = torch.zeros((per * G, 3 * kk + 1), dtype=torch.int64, device="cuda")
    out = torch.empty((per, 2 * k2 + 1), dtype=torch.int64, device="cuda")
    fake = (lambda t: t[None].expand(G, *t.shape).contiguous()) if G > 1 else (lambda t: None)
    marks = []

    def mark(name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append((name, ev))

    def batch(timed=False):
        m = mark if timed else (lambda name: None)
        m("start")
        bm = index.tc_begin(Q, k1, G)
        m("begin (query_prepare, bucket GEMM, bucket_topk)")
        bm_all = fake(bm)
        m("[emulated all-gather 1]")
        hist = index.tc_filter(bm_all)
        m("filter (tau_union, filter GEMM + survivor histogram)")
        if G > 1:
            hist *= G
        m("[emulated all-reduce]")
        index.tc_finish_packed(hist, kk, send)
        m("finish_packed (collect, rescore + fidelity, sort + pack)")
        recv = send.view(G, per, -1)[:1].expand(G, per, 3 * kk + 1).contiguous()
        ids = recv[:, :, 1 + kk:1 + 2 * kk]                           # copies of one shard's lists: make the ids distinct
        ids += (ids >= 0) * (torch.arange(G, device="cuda")[:, None, None] * n)
        m("[emulated all-to-all]")
        api.owner_finalize(recv, kk, k1, k2, "cosine", 0, nq, out)
        m("owner_finalize")

    for _ in range(3):
        batch()
    torch.cuda.synchronize()
    stage = {}
    for _ in range(a.steps):
        marks.clear()
        batch(timed=True)
        torch.cuda.synchronize()
        for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
            stage[name] = stage.get(name, 0.0) + e0.elapsed_time(e1) / a.steps

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps

    eager = timed(batch)
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        batch()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(graph):
        batch()
    rep = timed(graph.replay)
    hdr = send[:nq, 0]
    res = {**(extra or {}), "G": G, "shard_rows": n, "nq": nq, "k1": k1, "k2": k2, "kk": kk, "stage_ms": {k: round(v, 4) for k, v in stage.items()},
           "eager_ms": round(eager, 4), "graph_replay_ms": round(rep, 4),
           "valid_entries_max": int((hdr & 0xFFFFFFFF).max()), "flagged": int((hdr >> 32).count_nonzero())}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
