#!/usr/bin/env python
"""A/B sweep of the tensor-core search's per-query stages (config 3 shape by default).

    python tools/sweep_search.py [--N 1000000 --D 384 --nq 1024 --k 100 --metric cosine --steps 200]

Builds libqrag with -DQRAG_TUNING (the switches below exist only in that build), times whole batches with CUDA
events for each setting, checks that every setting returns the same (scores, ids), and restores the product build.
  QRAG_TC_TUNE=2          register form of tc_rescore where the bulk-copy form would be used
  QRAG_TC_TUNE=4          the bulk-copy form leaves the sort to tc_sort / tc_sort_pack
  QRAG_TC_TUNE=8          plain launches (no programmatic dependent launch between the kernels of a batch)
  QRAG_TC_RESCORE_Y=n     chunks of a query's candidates side by side (grid.y of tc_rescore)
  QRAG_TC_SORT_THREADS=n  block size of tc_sort / tc_sort_pack
  QRAG_TC_SAMPLE=n        pass 1 looks at every n-th document tile
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_rag_b200 import _lib, api  # noqa: E402

KNOBS = ("QRAG_TC_TUNE", "QRAG_TC_RESCORE_Y", "QRAG_TC_SORT_THREADS", "QRAG_TC_SAMPLE")

SETTINGS = [
    ("plain launches, register-form rescore", {"QRAG_TC_TUNE": "10"}),
    ("plain launches", {"QRAG_TC_TUNE": "8"}),
    ("product (chained launches)", {}),
    ("chained launches, separate sort", {"QRAG_TC_TUNE": "4"}),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1_000_000)
    ap.add_argument("--D", type=int, default=384)
    ap.add_argument("--nq", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--metric", default="cosine")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--rounds", type=int, default=15)
    a = ap.parse_args()
    _lib.build(force=True, tuning=True)
    g = torch.Generator(device="cuda").manual_seed(1237)
    X = torch.nn.functional.normalize(torch.randn(a.N, a.D, generator=g, device="cuda"), dim=1)
    Q = torch.nn.functional.normalize(torch.randn(a.nq, a.D, generator=g, device="cuda"), dim=1)
    index = api.FlatIndexTC(X, a.metric)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ref = None
    times = {name: [] for name, _ in SETTINGS}
    info = {}

    def use(env):
        for kname in KNOBS:
            os.environ.pop(kname, None)
        os.environ.update(env)

    for name, env in SETTINGS:                       # results first: every setting must return the same lists
        use(env)
        for _ in range(3):
            _, s, i, st = index.search_async(Q, a.k)
        torch.cuda.synchronize()
        if ref is None:
            ref = (s.clone(), i.clone())
        info[name] = (int(st.count_nonzero()), bool(torch.equal(s, ref[0]) and torch.equal(i, ref[1])))
    for _ in range(a.rounds):                        # then short timed bursts, the settings interleaved (the GEMM runs
        for name, env in SETTINGS:                   # at the power cap: a long run of one setting drifts by more than
            use(env)                                 # the differences looked for)
            index.search_async(Q, a.k)
            e0.record()
            for _ in range(a.steps):
                index.search_async(Q, a.k)
            e1.record()
            torch.cuda.synchronize()
            times[name].append(e0.elapsed_time(e1) / a.steps)
    for name, _ in SETTINGS:
        t = sorted(times[name])
        print(f"{name:46s} median {t[len(t) // 2]:.4f} min {t[0]:.4f} max {t[-1]:.4f} ms/batch  "
              f"flagged={info[name][0]} same_as_first={info[name][1]}", flush=True)


if __name__ == "__main__":
    try:
        main()
    finally:
        _lib.build(force=True)                # never leave the tuning build in the tree
