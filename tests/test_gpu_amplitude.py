"""K2 parity: amplitude-encoded fidelity, fused rerank and feature-map layers vs the oracle."""
import numpy as np
import pytest

from oracle import quantum as oq

pytestmark = pytest.mark.gpu

REL = 1e-12


@pytest.mark.parametrize("D", [3, 10, 128, 256, 384, 512, 1000, 1024, 1536])
@pytest.mark.parametrize("C", [1, 5, 100])
def test_amp_fidelity_dense(cuda, D, C):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(D + C)
    nq = 4
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    cand = rng.standard_normal((nq, C, D)).astype(np.float32)
    if C > 2:
        cand[1, 2] = 0.0                              # zero candidate -> fidelity 0
        cand[2, 1] = Q[2] * 3.0                       # parallel -> fidelity 1
    got, got32 = api.amp_fidelity(Q, cand=cand, want_fp32=True)
    want = oq.amplitude_fidelity_batch(Q, cand)
    assert np.allclose(got.cpu().numpy(), want, rtol=REL, atol=1e-16)   # atol: cancellation in q.d when F ~ 1e-7
    assert np.allclose(got32.cpu().numpy(), want, rtol=1e-5, atol=1e-30)      # north_star: 1e-5 relative in fp32
    if C > 2:
        assert float(got[1, 2]) == 0.0
        assert float(got[2, 1]) == pytest.approx(1.0, abs=1e-14)


def test_amp_fidelity_gather_and_padding_ids(cuda):
    """Gathered candidates (the plain-load kernel; the shipped library reads no tuning environment variable).
    Padding ids (-1) and ids beyond the corpus (>= N) both score -inf and sort last: never an out-of-bounds read."""
    from quantum_rag_b200 import api
    rng = np.random.RandomState(11)
    X = rng.standard_normal((300, 384)).astype(np.float32)
    Q = rng.standard_normal((6, 384)).astype(np.float32)
    idx = rng.randint(0, 300, size=(6, 33)).astype(np.int64)
    idx[0, 4] = idx[0, 9]                             # same row twice -> exact tie
    idx[3, 30:] = -1                                  # padding ids from a short search
    idx[5, 2], idx[5, 20] = 300, 2 ** 40              # out of range for X [300, D]: treated as padding
    got = api.amp_fidelity(Q, X=X, idx=idx).cpu().numpy()
    want = oq.amplitude_fidelity_batch(Q, X[np.where((idx >= 0) & (idx < 300), idx, 0)])
    want[3, 30:] = -np.inf
    want[5, 2] = want[5, 20] = -np.inf
    assert np.allclose(got, want, rtol=REL)
    assert got[0, 4] == got[0, 9]
    scores, pos, ids = api.quantum_rerank_batch(Q, X=X, idx=idx, top_k=33)
    order = oq.rank_rows(want)
    assert np.array_equal(pos.cpu().numpy(), order)
    assert np.array_equal(ids.cpu().numpy(), np.take_along_axis(idx, order, 1))
    assert ids[3, -3:].cpu().tolist() == [-1, -1, -1]


@pytest.mark.parametrize("C,top_k", [(1, 1), (20, 20), (100, 10), (128, 128), (129, 7), (1000, 10), (4096, 100)])
def test_fused_rerank_order_is_bit_exact(cuda, C, top_k):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(C)
    nq, D = 5, 384
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    cand = rng.standard_normal((nq, C, D)).astype(np.float32)
    if C >= 20:
        cand[:, 11] = cand[:, 3]                      # duplicate rows must tie and keep input order
        cand[:, 17] = cand[:, 3]
    scores, pos, ids = api.quantum_rerank_batch(Q, cand=cand, top_k=top_k, n_qubits=9)
    assert ids is None
    want = oq.amplitude_fidelity_batch(Q, cand)
    order = oq.rank_rows(want, top_k)
    assert np.array_equal(pos.cpu().numpy(), order)
    assert np.allclose(scores.cpu().numpy(), np.take_along_axis(want, order, 1), rtol=REL)
    # unfused path must agree bit for bit with the fused one
    full = api.amp_fidelity(Q, cand=cand)
    perm, srt = api.sort_scores(full, top_k)
    assert np.array_equal(perm.cpu().numpy(), pos.cpu().numpy())
    assert np.array_equal(srt.cpu().numpy(), scores.cpu().numpy())


@pytest.mark.parametrize("nq,C,D,top_k", [(3000, 1, 128, 1), (2500, 3, 384, 2), (1500, 7, 256, 7), (700, 20, 512, 5),
                                           (300, 150, 1024, 9), (40, 300, 1536, 300), (9, 2000, 64, 50),
                                           (2, 1000, 4096, 10), (333, 100, 384, 10)])
def test_fused_rerank_many_queries_all_shapes(cuda, nq, C, D, top_k):
    """Exercises every pipeline regime of the streaming kernel: more queries in flight than query slots,
    rings shorter than the warp count, ragged last tiles, both ranking algorithms."""
    from quantum_rag_b200 import api
    rng = np.random.RandomState(nq + C + D)
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    cand = rng.standard_normal((nq, C, D)).astype(np.float32)
    if C >= 3:
        cand[:, 2] = cand[:, 0]
    scores, pos, _ = api.quantum_rerank_batch(Q, cand=cand, top_k=top_k)
    want = oq.amplitude_fidelity_batch(Q, cand)
    order = oq.rank_rows(want, top_k)
    assert np.array_equal(pos.cpu().numpy(), order)
    assert np.allclose(scores.cpu().numpy(), np.take_along_axis(want, order, 1), rtol=REL, atol=1e-16)
    full = api.amp_fidelity(Q, cand=cand)
    assert np.allclose(full.cpu().numpy(), want, rtol=REL, atol=1e-16)
    assert np.array_equal(np.take_along_axis(full.cpu().numpy(), order, 1), scores.cpu().numpy())


def test_one_query_many_candidates_unfused(cuda):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(5)
    Q = rng.standard_normal((1, 384)).astype(np.float32)
    cand = rng.standard_normal((1, 50001, 384)).astype(np.float32)
    got = api.amp_fidelity(Q, cand=cand).cpu().numpy()
    assert np.allclose(got, oq.amplitude_fidelity_batch(Q, cand), rtol=REL, atol=1e-16)


def test_golden_amplitude_and_feature_map(cuda, kat):
    from quantum_rag_b200 import api
    for c in kat["oracle"]["amplitude"]:
        q = np.float32(c["q"])[None]
        d = np.float32(c["d"])[None, None]
        assert float(api.amp_fidelity(q, cand=d, n_qubits=c["n"])[0, 0]) == pytest.approx(c["f"], rel=REL)
    for c in kat["oracle"]["feature_map"]:
        q = np.float32(c["q"])[None]
        d = np.float32(c["d"])[None, None]
        got = float(api.amp_fidelity(q, cand=d, n_qubits=c["n"], layers=c["layers"])[0, 0])
        assert got == pytest.approx(c["f"], rel=REL)


@pytest.mark.parametrize("D,n,layers", [(5, 3, 1), (20, 5, 2), (48, 6, 3), (384, 9, 2), (1024, 10, 4)])
def test_feature_map_layers_match_oracle(cuda, D, n, layers):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(D)
    nq, C = 2, 4
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    cand = rng.standard_normal((nq, C, D)).astype(np.float32)
    cand[1, 1] = 0.0
    cand[0, 2] = Q[0]
    got = api.amp_fidelity(Q, cand=cand, n_qubits=n, layers=layers).cpu().numpy()
    want = np.array([[oq.feature_map_fidelity(Q[i], cand[i, j], n, layers) for j in range(C)] for i in range(nq)])
    assert np.allclose(got, want, rtol=1e-11, atol=1e-15)
    assert got[1, 1] == 0.0 and got[0, 2] == pytest.approx(1.0, abs=1e-12)


@pytest.mark.parametrize("D,layers", [(1024, 1), (1024, 4), (1000, 3), (515, 2), (7, 5), (1022, 9)])
def test_feature_map_warp_kernel_n10(cuda, D, layers):
    """n = 10 runs the warp-per-state register kernel: vs the oracle, and vs the shared-memory kernel."""
    import torch
    from quantum_rag_b200 import api
    rng = np.random.RandomState(D + layers)
    nq, C = 3, 29                                        # more items than warps in a CTA, ragged last round
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    cand = rng.standard_normal((nq, C, D)).astype(np.float32)
    cand[1, 1] = 0.0
    cand[0, 2] = Q[0]
    cand[2, 28] = cand[2, 5]                             # exact tie
    got = api.amp_fidelity(Q, cand=cand, n_qubits=10, layers=layers).cpu().numpy()
    sel = [(i, j) for i in range(nq) for j in (0, 1, 2, 5, 11, 28)]
    for i, j in sel:
        want = oq.feature_map_fidelity(Q[i], cand[i, j], 10, layers)
        assert got[i, j] == pytest.approx(want, rel=1e-11, abs=1e-15), (i, j)
    assert got[1, 1] == 0.0 and got[0, 2] == pytest.approx(1.0, abs=1e-12) and got[2, 28] == got[2, 5]
    api.set_fmap_kernel(api.FMAP_GENERIC)
    try:
        ref = api.amp_fidelity(Q, cand=cand, n_qubits=10, layers=layers).cpu().numpy()
    finally:
        api.set_fmap_kernel(api.FMAP_AUTO)
    assert np.allclose(got, ref, rtol=1e-11, atol=1e-15)
    # gathered rows with a padding id and a zero query
    X = cand.reshape(-1, D)
    idx = torch.from_numpy(rng.randint(0, X.shape[0], size=(nq, 40)))
    idx[1, 7] = -1
    Qz = Q.copy()
    Qz[2] = 0.0
    g = api.amp_fidelity(Qz, X=X, idx=idx, n_qubits=10, layers=layers).cpu().numpy()
    assert g[1, 7] == -np.inf and np.all(g[2] == 0.0)
    for j in (0, 39):
        want = oq.feature_map_fidelity(Q[0], X[int(idx[0, j])], 10, layers)
        assert g[0, j] == pytest.approx(want, rel=1e-11, abs=1e-15)


def test_feature_map_warp_kernel_many_units(cuda):
    """Enough queries that CTAs loop over several units, and a chunked query (nq small, C large)."""
    from quantum_rag_b200 import api
    rng = np.random.RandomState(77)
    for nq, C in ((400, 3), (1, 700)):
        Q = rng.standard_normal((nq, 1024)).astype(np.float32)
        cand = rng.standard_normal((nq, C, 1024)).astype(np.float32)
        got = api.amp_fidelity(Q, cand=cand, n_qubits=10, layers=2).cpu().numpy()
        api.set_fmap_kernel(api.FMAP_GENERIC)
        try:
            ref = api.amp_fidelity(Q, cand=cand, n_qubits=10, layers=2).cpu().numpy()
        finally:
            api.set_fmap_kernel(api.FMAP_AUTO)
        assert np.allclose(got, ref, rtol=1e-11, atol=1e-15)
        for i, j in ((0, 0), (nq - 1, C - 1)):
            assert got[i, j] == pytest.approx(oq.feature_map_fidelity(Q[i], cand[i, j], 10, 2), rel=1e-11, abs=1e-15)


def test_too_few_qubits_is_an_error(cuda):
    from quantum_rag_b200 import api
    from quantum_rag_b200._lib import QragError
    with pytest.raises(QragError, match="does not fit"):
        api.amp_fidelity(np.ones((1, 384), np.float32), cand=np.ones((1, 2, 384), np.float32), n_qubits=8)


def test_config2_full_size_properties(cuda):
    """BASELINE config 2 (1k queries x 100 candidates x 384-d, 9 qubits) through size-independent properties."""
    import torch
    from quantum_rag_b200 import api
    g = torch.Generator(device="cpu").manual_seed(1234 + 2)
    nq, C, D = 1000, 100, 384
    Q = torch.nn.functional.normalize(torch.randn(nq, D, generator=g), dim=1).cuda()
    cand = torch.nn.functional.normalize(torch.randn(nq, C, D, generator=g), dim=2).cuda()
    cand[:, 42] = Q                                    # planted self-match: fidelity 1, must rank first
    cand[:, 77] = cand[:, 5]                           # planted tie
    scores, pos, _ = api.quantum_rerank_batch(Q, cand=cand, top_k=C, n_qubits=9)
    full = api.amp_fidelity(Q, cand=cand)
    assert torch.all(pos[:, 0] == 42) and torch.allclose(scores[:, 0], torch.ones(nq, dtype=torch.float64, device="cuda"),
                                                         atol=1e-6)
    assert torch.all(scores[:, :-1] >= scores[:, 1:])                          # sortedness
    assert torch.equal(torch.gather(full, 1, pos.long()), scores)              # scores are the gathered fidelities
    assert torch.all(torch.sort(pos.long(), dim=1).values == torch.arange(C, device="cuda"))   # a permutation
    p5 = (pos == 5).nonzero()[:, 1]
    p77 = (pos == 77).nonzero()[:, 1]
    assert torch.all(p77 == p5 + 1)                                            # tie keeps input order, adjacent
    assert float(full.min()) >= 0.0 and float(full.max()) <= 1.0 + 1e-12
    # scale invariance of the state preparation
    full2 = api.amp_fidelity(Q * 4.0, cand=cand * 0.5)      # power-of-two scalings are exact in fp32
    assert torch.allclose(full, full2, rtol=1e-12, atol=1e-18)
    # sample rows against the oracle
    rows = [0, 499, 999]
    want = oq.amplitude_fidelity_batch(Q[rows].cpu().numpy(), cand[rows].cpu().numpy())
    assert np.allclose(full[rows].cpu().numpy(), want, rtol=REL)


def test_config5_full_size_properties(cuda):
    """BASELINE config 5 (4096 queries x 1000 candidates x 1024-d, 10 qubits, 4 feature-map layers): 4.1M scores,
    candidates named by ids into a resident corpus (the dense form would be 16.8 GB), size-independent properties."""
    import torch
    from quantum_rag_b200 import api
    g = torch.Generator(device="cuda").manual_seed(1234 + 5)
    nq, C, D, N, L = 4096, 1000, 1024, 50_000, 4
    X = torch.randn(N, D, generator=g, device="cuda")
    qrow = torch.randint(0, N, (nq,), generator=g, device="cuda")
    Q = X[qrow].clone()                                      # every query is a corpus row
    idx = torch.randint(0, N, (nq, C), generator=g, device="cuda")
    idx[:, 17] = qrow                                        # planted self-match
    idx[:, 900] = idx[:, 3]                                  # planted duplicate
    idx[5, 10] = -1                                          # padding id
    F = api.amp_fidelity(Q, X=X, idx=idx, n_qubits=10, layers=L)
    assert F.shape == (nq, C)
    assert torch.allclose(F[:, 17], torch.ones(nq, dtype=torch.float64, device="cuda"), atol=1e-12)   # |<psi|psi>|^2 = 1
    assert torch.equal(F[:, 900], F[:, 3])                   # same row -> bit-identical score
    assert float(F[5, 10]) == float("-inf")
    fin = F[torch.isfinite(F)]
    assert float(fin.min()) >= 0.0 and float(fin.max()) <= 1.0 + 1e-12
    # symmetry: F(q, d) = F(d, q), evaluated by swapping the roles for one column
    d_rows = idx[:, 1]
    Fsw = api.amp_fidelity(X[d_rows].contiguous(), X=X, idx=qrow[:, None].contiguous(), n_qubits=10, layers=L)
    assert torch.allclose(Fsw[:, 0], F[:, 1], rtol=1e-10, atol=1e-18)
    # scale invariance of state preparation and angles (power-of-two scaling is exact in fp32)
    F2 = api.amp_fidelity(Q[:64] * 8.0, X=X * 0.25, idx=idx[:64], n_qubits=10, layers=L)
    assert torch.allclose(F2[torch.isfinite(F2)], F[:64][torch.isfinite(F[:64])], rtol=1e-10, atol=1e-18)
    # rerank of the full lists: sorted, stable, the self-match first
    scores, pos, ids = api.quantum_rerank_batch(Q, X=X, idx=idx, top_k=10, n_qubits=10, layers=L)
    assert torch.all(pos[:, 0] <= 17) and torch.all(ids[:, 0] == qrow)      # (a random id may name the same row earlier)
    assert torch.all(scores[:, :-1] >= scores[:, 1:])
    # spot checks against the oracle
    for i, j in ((0, 0), (2048, 999), (4095, 500)):
        want = oq.feature_map_fidelity(Q[i].cpu().numpy(), X[idx[i, j]].cpu().numpy(), 10, L)
        assert float(F[i, j]) == pytest.approx(want, rel=1e-10, abs=1e-16)


def test_overlap_policies_give_identical_results_back_to_back(cuda):
    """Programmatic dependent launch: consecutive launches on one stream overlap, results must not change."""
    import torch
    from quantum_rag_b200 import api
    rng = np.random.RandomState(21)
    batches = [(torch.from_numpy(rng.standard_normal((600, 384)).astype(np.float32)).cuda(),
                torch.from_numpy(rng.standard_normal((600, 100, 384)).astype(np.float32)).cuda()) for _ in range(3)]
    results = {}
    old = api.set_overlap(api.OVERLAP_SAFE)
    try:
        for mode in (api.OVERLAP_NONE, api.OVERLAP_SAFE, api.OVERLAP_INPUTS_STABLE):
            api.set_overlap(mode)
            outs = []
            for rep in range(4):                         # 12 launches back to back, no host sync in between
                for Q, cand in batches:
                    outs.append(api.quantum_rerank_batch(Q, cand=cand, top_k=10)[:2])
            torch.cuda.synchronize()
            results[mode] = outs
    finally:
        api.set_overlap(old)
    base = results[api.OVERLAP_NONE]
    for mode, outs in results.items():
        for (s, p), (bs, bp) in zip(outs, base):
            assert torch.equal(s, bs) and torch.equal(p, bp), f"overlap mode {mode}"
    want = oq.rank_rows(oq.amplitude_fidelity_batch(batches[0][0].cpu().numpy(), batches[0][1].cpu().numpy()), 10)
    assert np.array_equal(base[0][1].cpu().numpy(), want)
    with pytest.raises(Exception):
        api.set_overlap(7)


@pytest.mark.parametrize("nq,C,top_k,descending", [(3, 5000, None, True), (2, 100000, 25, True), (1, 100000, None, False),
                                                   (5, 4097, 4097, True), (4, 12289, 7, False), (1, 4096, None, True)])
def test_long_stable_sort_vs_oracle(cuda, nq, C, top_k, descending):
    """Lists longer than the shared-memory sort (block sort + global-memory merge levels, no library sort on the
    path): the reference's `sorted(..., reverse=True)` order -- ties keep their input order -- at C = 5 000 and 100 000."""
    import torch
    from quantum_rag_b200 import api
    rng = np.random.RandomState(C % 1000 + nq)
    scores = np.round(rng.standard_normal((nq, C)), 2)                  # two decimals: thousands of exact ties
    scores[:, C // 2] = scores[:, 1]
    perm, srt = api.sort_scores(torch.from_numpy(scores).cuda(), top_k, descending=descending)
    k = C if top_k is None else top_k
    for q in range(nq):
        want = oq.stable_rank(scores[q].tolist(), k) if descending else \
            sorted(range(C), key=lambda i: scores[q, i])[:k]             # Python's sort is stable either way
        assert perm[q].cpu().tolist() == want
        assert np.array_equal(srt[q].cpu().numpy(), scores[q][want])


def test_interleaved_overlap_mode_back_to_back_batches(cuda):
    """QRAG_OVERLAP_INTERLEAVED: half-size CTAs, two consecutive launches share every SM, results staged in shared
    memory and written after the previous kernel completed.  Back-to-back batches into the SAME output arrays (the
    write-after-write case the staging exists for) and into fresh ones: every batch equals the oracle's ranking."""
    import ctypes
    import torch
    from quantum_rag_b200 import _lib, api
    rng = np.random.RandomState(31)
    nq, C, D, k = 300, 100, 384, 10
    batches = []
    for b in range(6):
        Q = rng.standard_normal((nq, D)).astype(np.float32)
        cand = rng.standard_normal((nq, C, D)).astype(np.float32)
        cand[:, 40] = cand[:, 7]                                     # exact ties
        batches.append((torch.from_numpy(Q).cuda(), torch.from_numpy(cand).cuda(), Q, cand))
    old = api.set_overlap(api.OVERLAP_INTERLEAVED)
    try:
        fresh = [api.quantum_rerank_batch(Qd, cand=cd, top_k=k, n_qubits=9) for Qd, cd, _, _ in batches]
        lib = _lib.load()
        scores = torch.empty((nq, k), dtype=torch.float64, device="cuda")
        pos = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        snaps = []
        for Qd, cd, _, _ in batches:                                 # same outputs every launch, copied out in stream order
            _lib.check(lib.qrag_amp_rerank(api._ptr(Qd), nq, api._ptr(cd), None, 0, None, C, D, 9, k, api._ptr(scores),
                                           api._ptr(pos), None, api._stream()))
            snaps.append((scores.clone(), pos.clone()))
        torch.cuda.synchronize()
    finally:
        api.set_overlap(old)
    for (s1, p1, _), (s2, p2), (_, _, Q, cand) in zip(fresh, snaps, batches):
        want = oq.amplitude_fidelity_batch(Q, cand)
        order = oq.rank_rows(want, k)
        assert np.array_equal(p1.cpu().numpy(), order) and np.array_equal(p2.cpu().numpy(), order)
        assert np.allclose(s1.cpu().numpy(), np.take_along_axis(want, order, 1), rtol=REL)
        assert torch.equal(s1, s2)


def test_feature_map_filter_error_is_inside_the_bound(cuda):
    """The complex64 filter pass of the feature-map rerank against the complex128 kernel: max |F32 - F| MEASURED and
    compared with delta = qrag_fmap_filter_error_bound(layers).  Ordinary data (scaled-rotation path), large angles
    (direct-rotation path: some |x^_i| > 1/2), near-duplicates of the query (fidelity ~ 1, where the bound's factor
    2 |<d|q>| is largest), short rows (D = 7) and the deepest circuit the register kernel takes."""
    import torch
    from quantum_rag_b200 import api
    rng = np.random.RandomState(3)
    worst = 0.0
    for D, layers in ((1024, 4), (1024, 1), (600, 6), (7, 5), (1024, 9)):
        nq, C = 6, 200
        Q = rng.standard_normal((nq, D)).astype(np.float32)
        cand = rng.standard_normal((nq, C, D)).astype(np.float32)
        cand[:, :20] = Q[:, None, :] + 1e-2 * rng.standard_normal((nq, 20, D)).astype(np.float32)   # fidelity near 1
        cand[:, 20] = Q
        if D >= 600:
            cand[:, 21:40, 3] = 40.0                                        # dominant component: |x^_3| > 1/2, direct rotations
        f64 = api.amp_fidelity(Q, cand=cand, n_qubits=10, layers=layers)
        f32 = api.fmap_filter_scores(Q, cand=cand, layers=layers)
        delta = api.fmap_filter_error_bound(layers)
        err = float((f32 - f64).abs().max())
        assert err <= delta, (D, layers, err, delta)
        assert err > 0.0                                                    # it IS a different arithmetic
        worst = max(worst, err / delta)
    assert worst < 0.5, worst                                              # the bound has room (worst-case analysis x 2)
    print(f"feature-map filter: max |F32 - F| = {worst:.3f} delta")


@pytest.mark.parametrize("C,top_k,gathered", [(1000, 10, False), (300, 40, True), (64, 64, False), (4096, 5, False)])
def test_feature_map_rerank_equals_the_complex128_path(cuda, C, top_k, gathered):
    """qrag_fmap_rerank (complex64 filter + complex128 certification of the top-k boundary) returns the positions,
    ids and score BITS of the all-complex128 path (qrag_amp_fidelity + stable sort): exact duplicates tie in input
    order, near-ties closer than the filter's error are certified, padding ids sort last."""
    import torch
    from quantum_rag_b200 import api
    rng = np.random.RandomState(C + top_k)
    nq, D, L = 5, 1024, 4
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    if gathered:
        X = rng.standard_normal((500, D)).astype(np.float32)
        idx = rng.randint(0, 500, size=(nq, C)).astype(np.int64)          # C of 500 rows: repeated rows = exact ties
        idx[2, 17:25] = -1                                                 # padding
        idx[4, 3] = 10 ** 9                                                # out of range = padding
        kw = {"X": X, "idx": idx}
        rows = X[np.where((idx >= 0) & (idx < 500), idx, 0)]
    else:
        cand = rng.standard_normal((nq, C, D)).astype(np.float32)
        cand[:, 5] = cand[:, 2]                                            # exact tie
        cand[1, 9] = cand[1, 2] * (1 + 1e-7)                               # near-tie far inside the filter's error
        kw = {"cand": cand}
    scores, pos, ids = api.quantum_rerank_batch(Q, top_k=top_k, n_qubits=10, layers=L, **kw)
    es, ep, ei = api.quantum_rerank_batch(Q, top_k=top_k, n_qubits=10, layers=L, certify=False, **kw)
    assert torch.equal(pos, ep) and torch.equal(scores, es)
    if gathered:
        assert torch.equal(ids, ei)
        assert int(pos[2].min()) >= 0 or top_k > C - 8                     # padding never beats a real candidate
    # and against the oracle on a few entries: the ranking is the oracle's
    q0 = 1
    want = np.array([oq.feature_map_fidelity(Q[q0], (rows if gathered else cand)[q0, int(p)], 10, L)
                     for p in pos[q0, :min(top_k, 6)].cpu().tolist()])
    assert np.allclose(scores[q0, :min(top_k, 6)].cpu().numpy(), want, rtol=1e-11)


def test_feature_map_rerank_crowded_margin_is_flagged_and_rerun(cuda):
    """More candidates inside the filter's error margin than the certify list holds (here: one row repeated 300 times
    at the top): status flags the query and the API reruns it through the exact path -- same answer, never silent."""
    import ctypes
    import torch
    from quantum_rag_b200 import _lib, api
    rng = np.random.RandomState(5)
    nq, C, D, L, k = 3, 400, 1024, 2, 10
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    cand = rng.standard_normal((nq, C, D)).astype(np.float32)
    cand[1, 50:350] = Q[1]                                                 # 300 exact copies of the query: all tie at F = 1
    Qd, cd = torch.from_numpy(Q).cuda(), torch.from_numpy(cand).cuda()
    lib = _lib.load()
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.qrag_fmap_rerank_workspace(nq, C, k, ctypes.byref(nbytes)))
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    s = torch.empty((nq, k), dtype=torch.float64, device="cuda")
    p = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    st = torch.empty(nq, dtype=torch.int32, device="cuda")
    _lib.check(lib.qrag_fmap_rerank(api._ptr(Qd), nq, api._ptr(cd), None, 0, None, C, D, 10, L, k, api._ptr(s), api._ptr(p), None,
                                    api._ptr(st), api._ptr(ws), nbytes.value, api._stream()))
    assert st.cpu().tolist() == [0, 1, 0]
    scores, pos, _ = api.quantum_rerank_batch(Q, cand=cand, top_k=k, n_qubits=10, layers=L)
    es, ep, _ = api.quantum_rerank_batch(Q, cand=cand, top_k=k, n_qubits=10, layers=L, certify=False)
    assert torch.equal(pos, ep) and torch.equal(scores, es)
    assert pos[1].cpu().tolist() == list(range(50, 60))                    # the first ten copies, in input order


def test_host_pipelines_equal_the_device_api(cuda):
    """The two end-to-end forms the bench times (pinned host buffers in, pinned host results out, slices on two
    streams): dense candidates, and candidate ids against a resident corpus.  Same bits as the one-launch device API."""
    import torch
    from quantum_rag_b200 import api
    rng = np.random.RandomState(23)
    nq, C, D, k = 203, 60, 384, 7                                        # 203 queries: the last slice is ragged
    Q = torch.from_numpy(rng.standard_normal((nq, D)).astype(np.float32)).pin_memory()
    cand = torch.from_numpy(rng.standard_normal((nq, C, D)).astype(np.float32)).pin_memory()
    pipe = api.HostRerankPipeline(nq, C, D, k, 9, chunks=4)
    hS, hP = pipe(Q, cand)
    s, p, _ = api.quantum_rerank_batch(Q, cand=cand, top_k=k, n_qubits=9)
    assert torch.equal(hS, s.cpu()) and torch.equal(hP, p.cpu())
    X = torch.from_numpy(rng.standard_normal((5000, D)).astype(np.float32))
    idx = torch.from_numpy(rng.randint(0, 5000, size=(nq, C)).astype(np.int64)).pin_memory()
    idx[5, 10:20] = -1
    s2, p2, i2 = api.quantum_rerank_batch(Q, X=X, idx=idx, top_k=k, n_qubits=9)
    for depth in (1, 3):
        idp = api.HostIdRerankPipeline(X, nq, C, k, 9, depth=depth)
        for _ in range(2):                                                # reusable: the second call gives the same answer
            hS2, hO2 = idp(Q, idx)
        assert torch.equal(hS2, s2.cpu()) and torch.equal(hO2, i2.cpu())
    # several batches in flight (one library call each: copies in, kernel, copies out on the slot's stream); every
    # ticket returns its own batch; a fourth submit before any result() is refused
    Qb = [torch.from_numpy(rng.standard_normal((nq, D)).astype(np.float32)).pin_memory() for _ in range(3)]
    Ib = [torch.from_numpy(rng.randint(-2, 5003, size=(nq, C)).astype(np.int64)).pin_memory() for _ in range(3)]
    tickets = [idp.submit(q, i) for q, i in zip(Qb, Ib)]
    with pytest.raises(RuntimeError):
        idp.submit(Qb[0], Ib[0])
    for t, q, i in zip(tickets, Qb, Ib):
        hS3, hO3 = idp.result(t)
        s3, _, i3 = api.quantum_rerank_batch(q, X=X, idx=i, top_k=k, n_qubits=9)
        assert torch.equal(hS3, s3.cpu()) and torch.equal(hO3, i3.cpu())
    with pytest.raises(RuntimeError):
        idp.result(tickets[0])
    with pytest.raises(ValueError):
        idp.submit(Qb[0][:5], Ib[0][:5])
    want = oq.rank_rows(oq.amplitude_fidelity_batch(Q.numpy()[:3], X.numpy()[idx.numpy()[:3]]), k)
    assert np.array_equal(p2[:3].cpu().numpy(), want)
