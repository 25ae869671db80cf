#!/usr/bin/env python
"""K1 alone: the reference circuit (quantum.py:138-167) on batches of (query, document) angle vectors, per n_qubits.
Prints ms per batch and pairs/s (CUDA events); --layers for the multi-layer extension."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_rag_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nq", type=int, default=1000)
    ap.add_argument("--per", type=int, default=100)
    ap.add_argument("--layers", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--n", type=int, nargs="*", default=[4, 5, 6, 9, 10, 12])
    a = ap.parse_args()
    for n in a.n:
        g = torch.Generator(device="cuda").manual_seed(n)
        vec_len = 2 * n if n <= 5 else n
        q = torch.rand(a.nq, vec_len, generator=g, device="cuda", dtype=torch.float64)
        d = torch.rand(a.nq * a.per, vec_len, generator=g, device="cuda", dtype=torch.float64)
        for _ in range(3):
            out = api.sv_fidelity_angle(q, d, docs_per_query=a.per, n_qubits=n, layers=a.layers)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            out = api.sv_fidelity_angle(q, d, docs_per_query=a.per, n_qubits=n, layers=a.layers)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        print(f"n={n:2d} layers={a.layers} pairs={a.nq * a.per}: {ms * 1e3:9.1f} us per batch, "
              f"{a.nq * a.per / ms * 1e3:.3e} pairs/s, checksum {float(out.sum()):.12f}", flush=True)


if __name__ == "__main__":
    main()
