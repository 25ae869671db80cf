#!/usr/bin/env python
"""cProfile of the drop-in string API (RerankerController.rerank, 20 documents, n_qubits = 4): where the host time goes."""
import cProfile, pstats, sys, io
sys.path.insert(0, ".")
import numpy as np, torch
from quantum_rag_b200.reranker import Document, RerankerController
ctl = RerankerController()
rng = np.random.RandomState(0)
words = ["sponsor", "episode", "panel", "election", "discount", "code", "piers", "morgan", "news", "tonight"]
docs = [Document(str(i), " ".join(rng.choice(words, 12))) for i in range(20)]
q = "which segments contain a sponsor advertisement"
for _ in range(50): ctl.rerank(q, docs, top_k=5, reranker_type="quantum")
pr = cProfile.Profile(); pr.enable()
for _ in range(500): ctl.rerank(q, docs, top_k=5, reranker_type="quantum")
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
