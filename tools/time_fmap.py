#!/usr/bin/env python
"""Timing of the statevector kernels: config 5 (amplitude state + L feature-map layers, n = 10) and the
reference angle circuit at n = 4 / 9 (config 1 / 2 variants).

    python tools/time_fmap.py [--nq 256 --C 1000 --D 1024 --layers 4]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_rag_b200 import _lib, api  # noqa: E402


def timed(fn, steps):
    for _ in range(2):
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
    ap.add_argument("--nq", type=int, default=256)
    ap.add_argument("--C", type=int, default=1000)
    ap.add_argument("--D", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    _lib.build()
    g = torch.Generator(device="cuda").manual_seed(1239)
    Q = torch.randn(a.nq, a.D, generator=g, device="cuda")
    cand = torch.randn(a.nq, a.C, a.D, generator=g, device="cuda")
    n = api.qubits_for(a.D)
    ms = timed(lambda: api.amp_fidelity(Q, cand=cand, n_qubits=n, layers=a.layers), a.steps)
    scores = a.nq * a.C
    flop = a.layers * 12 * n * 2 ** n + 8 * 2 ** n
    print(f"config 5 shape: {a.nq} x {a.C} x {a.D}, n={n}, L={a.layers}: {ms:.3f} ms, {scores / ms * 1e3:.3e} scores/s, "
          f"{scores * flop / ms / 1e9:.2f} TFLOP/s fp64-equivalent ({flop} flop/score), "
          f"{scores * 4 * a.D / ms / 1e6:.1f} GB/s of candidate rows")
    if n == 10:
        msf = timed(lambda: api.fmap_filter_scores(Q, cand=cand, layers=a.layers), a.steps)
        print(f"complex64 filter pass alone: {msf:.3f} ms, {scores / msf * 1e3:.3e} scores/s")
        msr = timed(lambda: api.quantum_rerank_batch(Q, cand=cand, top_k=10, n_qubits=n, layers=a.layers), a.steps)
        print(f"rerank top-10, filter + certify: {msr:.3f} ms, {scores / msr * 1e3:.3e} scores/s")
    ms0 = timed(lambda: api.amp_fidelity(Q, cand=cand, n_qubits=n, layers=0), a.steps)
    print(f"same shape, layers=0 (HBM-bound streaming kernel): {ms0:.3f} ms, {scores / ms0 * 1e3:.3e} scores/s, "
          f"{scores * 4 * a.D / ms0 / 1e6:.1f} GB/s")
    for nqb, nd in ((4, 1_000_000), (9, 200_000), (10, 100_000)):
        rng = np.random.RandomState(nqb)
        qv = torch.from_numpy(rng.random_sample((1000, 2 * nqb))).cuda()
        dv = torch.from_numpy(rng.random_sample((nd, 2 * nqb))).cuda()
        ms1 = timed(lambda: api.sv_fidelity_angle(qv, dv, docs_per_query=nd // 1000, n_qubits=nqb), a.steps)
        print(f"reference circuit n={nqb}: {nd} pairs in {ms1:.3f} ms = {nd / ms1 * 1e3:.3e} pairs/s")


if __name__ == "__main__":
    main()
