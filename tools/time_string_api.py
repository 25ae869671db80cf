#!/usr/bin/env python
"""Latency of the drop-in string API (config 1): RerankerController.rerank on 20 documents, reference circuit n = 4."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from src.reranker.classical import Document  # noqa: E402
from src.reranker.controller import RerankerController  # noqa: E402


def main():
    ctl = RerankerController()
    rng = np.random.RandomState(0)
    words = ["sponsor", "episode", "panel", "election", "discount", "code", "piers", "morgan", "news", "tonight"]
    docs = [Document(str(i), " ".join(rng.choice(words, 12))) for i in range(20)]
    query = "which segments contain a sponsor advertisement"
    for _ in range(20):
        ctl.rerank(query, docs, top_k=5, reranker_type="quantum")
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        res = ctl.rerank(query, docs, top_k=5, reranker_type="quantum")
    dt = (time.perf_counter() - t0) / n
    print(f"quantum rerank of 20 documents through the string API: {dt * 1e6:.1f} us per call "
          f"({20 / dt:.0f} pairs/s); top id {res['documents'][0][0].id}")
    docs = [Document(str(i), " ".join(rng.choice(words, 12))) for i in range(1000)]
    for _ in range(3):
        ctl.rerank(query, docs, top_k=10, reranker_type="quantum")
    t0 = time.perf_counter()
    for _ in range(20):
        ctl.rerank(query, docs, top_k=10, reranker_type="quantum")
    dt = (time.perf_counter() - t0) / 20
    print(f"quantum rerank of 1000 documents through the string API: {dt * 1e3:.2f} ms per call ({1000 / dt:.0f} pairs/s)")


if __name__ == "__main__":
    main()
