#!/usr/bin/env python
"""One-GPU probe of what a rank of the 8-GPU sharded search (config 4) spends outside its big kernels.

A shard of 10M / G rows is searched with the phase-split tcgen05 search; the two threshold exchanges are
emulated by replicating the shard's own lists G times (same statistics as G equal shards).  Reports the time
per batch launched eagerly, and replayed from a CUDA graph (no CPU launch gaps), so the difference is what the
host costs.

    python tools/shard_stage_probe.py [--G 8 --N 10000000 --nq 1024 --k1 1000]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_rag_b200 import _lib, api  # noqa: E402


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--G", type=int, default=8)
    ap.add_argument("--N", type=int, default=10_000_000)
    ap.add_argument("--nq", type=int, default=1024)
    ap.add_argument("--k1", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    _lib.build()
    n = a.N // a.G
    g = torch.Generator(device="cuda").manual_seed(1238)
    X = torch.nn.functional.normalize(torch.randn(n, 384, generator=g, device="cuda"), dim=1)
    Q = torch.nn.functional.normalize(torch.randn(a.nq, 384, generator=g, device="cuda"), dim=1)
    index = api.FlatIndexTC(X, "cosine")
    fake = lambda t: t[None].expand(a.G, *t.shape).contiguous()          # noqa: E731
    out = {}

    def search():
        out["r"] = index.search_sharded(Q, a.k1, fake, a.G)

    def full():
        s, i, st = index.search_sharded(Q, a.k1, fake, a.G)
        kk = max(1, min(a.k1, -(-(2 * a.k1 // a.G + 64) // 32) * 32))
        s, i = s[:, :kk].contiguous(), i[:, :kk].contiguous()
        f = api.amp_fidelity(Q, X=X, idx=torch.where(i >= 0, i, torch.full_like(i, -1)))
        out["f"] = f

    for name, fn in (("search phases", search), ("search + own-entry fidelity", full)):
        eager = timed(fn, a.steps)
        graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(graph):
            fn()
        rep = timed(graph.replay, a.steps)
        print(f"{name}: shard of {n} rows, G={a.G}: eager {eager:.3f} ms, CUDA graph replay {rep:.3f} ms")
    s, i, st = out["r"]
    print("valid entries per query: max", int((i >= 0).sum(1).max()), "flagged", int(st.count_nonzero()))


if __name__ == "__main__":
    main()
