#!/usr/bin/env python
"""Tuning sweep for the streaming amplitude-rerank kernel (config 2 shape by default).

    python tools/sweep_amp.py [--nq 1000 --C 100 --D 384 --k 10]

Sets QRAG_AMP_STREAM_G / QRAG_AMP_STREAM_RB (debug overrides read by amp_stream.cu) and times
qrag_amp_rerank with CUDA events over rotating input sets (larger than L2 in total).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_rag_b200 import _lib, api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nq", type=int, default=1000)
    ap.add_argument("--C", type=int, default=100)
    ap.add_argument("--D", type=int, default=384)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--configs", default="default,4x4,8x4,8x2")
    ap.add_argument("--overlap", default="0,1,2,3")
    ap.add_argument("--gather", action="store_true", help="candidates named by random ids into a resident corpus")
    ap.add_argument("--region", type=int, default=0, help="also time regions of this many steps with a synchronise "
                    "between them (what bench.py --steps K measures), median of 15")
    ap.add_argument("--cps", default="1", help="comma list of QRAG_AMP_STREAM_CPS values (CTAs per SM and launch, interleaved shape)")
    ap.add_argument("--no-build", action="store_true", help="the tuning build is already in the tree (built where nvcc is fast)")
    a = ap.parse_args()
    if not a.no_build:
        _lib.build(force=True, tuning=True)   # the switches below only exist in a -DQRAG_TUNING build
    lib = _lib.load()
    g = torch.Generator().manual_seed(7)
    nsets = max(2, int(520e6 / (a.nq * a.C * a.D * 4)) + 1)
    sets = [(torch.randn(a.nq, a.D, generator=g).cuda(), torch.randn(a.nq, a.C, a.D, generator=g).cuda())
            for _ in range(nsets)]
    scores = torch.empty((a.nq, a.k), dtype=torch.float64, device="cuda")
    pos = torch.empty((a.nq, a.k), dtype=torch.int32, device="cuda")
    nbytes = a.nq * a.C * a.D * 4

    if a.gather:
        os.environ["QRAG_AMP_STREAM_GATHER"] = "1"        # the sweep is about the streaming kernel
    corpus = torch.cat([c.view(-1, a.D) for _, c in sets], dim=0) if a.gather else None
    idxs = [torch.randint(0, corpus.shape[0], (a.nq, a.C), generator=g).cuda() for _ in range(nsets)] if a.gather else None

    def step(i):
        Q, c = sets[i % nsets]
        if a.gather:
            _lib.check(lib.qrag_amp_rerank(api._ptr(Q), a.nq, None, api._ptr(corpus), corpus.shape[0], api._ptr(idxs[i % nsets]), a.C, a.D,
                                           api.qubits_for(a.D), a.k, api._ptr(scores), api._ptr(pos), None, api._stream()))
            return
        _lib.check(lib.qrag_amp_rerank(api._ptr(Q), a.nq, api._ptr(c), None, 0, None, a.C, a.D, api.qubits_for(a.D), a.k,
                                       api._ptr(scores), api._ptr(pos), None, api._stream()))

    ref = None
    for cfg, ov, cps in [(c, int(o), x) for c in a.configs.split(",") for o in a.overlap.split(",") for x in a.cps.split(",")]:
        api.set_overlap(ov)
        os.environ["QRAG_AMP_STREAM_CPS"] = cps
        os.environ.pop("QRAG_AMP_STREAM_G", None)
        os.environ.pop("QRAG_AMP_STREAM_RB", None)
        if cfg != "default":
            gg, rb = cfg.split("x")
            os.environ["QRAG_AMP_STREAM_G"] = gg
            os.environ["QRAG_AMP_STREAM_RB"] = rb
        step(0)
        torch.cuda.synchronize()
        cur = (scores.clone(), pos.clone())
        if ref is None:
            ref = cur
        same = bool(torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]))
        for i in range(20):
            step(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(a.steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / a.steps
        reg = ""
        if a.region:
            ts = []
            for _ in range(15):
                torch.cuda.synchronize()
                e0.record()
                for i in range(a.region):
                    step(i)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3 / a.region)
            ts.sort()
            reg = f"  regions of {a.region}: median {ts[7]:.2f} us/launch ({nbytes / ts[7] / 1e3:.0f} GB/s), min {ts[0]:.2f}"
        print(f"GxRB={cfg:8s} overlap={ov} cps={cps} {us:8.2f} us/launch  {nbytes / us / 1e3:8.1f} GB/s  same_as_first={same}{reg}", flush=True)


if __name__ == "__main__":
    try:
        main()
    finally:
        if "--no-build" not in sys.argv:
            _lib.build(force=True)            # never leave the tuning build in the tree: what ships reads no environment
