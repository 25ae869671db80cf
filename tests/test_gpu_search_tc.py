"""K3/K4 parity: the tcgen05 search returns exactly what the exact CUDA-core search and the oracle return."""
import numpy as np
import pytest

from oracle import search as osr

pytestmark = pytest.mark.gpu

METRICS = [("ip", osr.METRIC_IP), ("l2", osr.METRIC_L2), ("cosine", osr.METRIC_COSINE)]


def _compare_with_exact(api, Q, X, k, name, id_base=0, allow_fallback=False):
    import torch
    index = api.FlatIndexTC(X, name, id_base=id_base)
    s, i = index.search(Q, k)
    if not allow_fallback:
        assert index.last_fallback == 0, f"{index.last_fallback} queries fell back to the exact path"
    es, ei = api.search_topk(Q, X, k, name, id_base=id_base)
    assert torch.equal(i, ei), name
    assert torch.equal(s, es), name                     # bit-identical fp64 scores (same scoring code)
    return index


@pytest.mark.parametrize("name,mid", METRICS)
@pytest.mark.parametrize("nq,N,D,k", [(5, 3000, 64, 10), (130, 20000, 384, 100), (1, 70000, 384, 100),
                                      (300, 66000, 128, 37), (64, 50000, 96, 1000), (7, 9000, 40, 7),
                                      (33, 30011, 1536, 20), (9, 5000, 50, 12), (3, 4000, 33, 5),     # D % 4 != 0: scalar rows
                                      # >= 592 queries: one rescoring CTA per query, which loops over the list's chunks
                                      # and (rows long enough to hold the list in their stages) sorts it too
                                      (640, 40000, 384, 300), (600, 20000, 64, 200)])
def test_tc_search_equals_exact_search(cuda, name, mid, nq, N, D, k):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(nq + N + D + k)
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    X[17] = X[3]
    X[N - 1] = X[3]                                     # duplicates far apart: exact ties, id order decides
    Q[0] = X[3]
    _compare_with_exact(api, Q, X, k, name, id_base=12345)


@pytest.mark.parametrize("name,mid", METRICS)
def test_tc_search_matches_oracle_unit_rows(cuda, name, mid):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(3)
    X = rng.standard_normal((12000, 384)).astype(np.float32)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    Q = rng.standard_normal((20, 384)).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    index = api.FlatIndexTC(X, name)
    s, i = index.search(Q, 50)
    rs, ri = osr.exact_search(Q, X, 50, mid)
    assert np.array_equal(i.cpu().numpy(), ri)
    assert np.allclose(s.cpu().numpy(), rs, rtol=1e-12, atol=1e-13)


def test_tc_search_on_reference_fixture(cuda, piers, kat):
    """Config 1 through the tensor-core path: the reference's own index, every row as the query."""
    from quantum_rag_b200 import api
    x = piers["vectors"]
    index = api.FlatIndexTC(x, "l2")
    s, i = index.search(x, 20)
    assert np.array_equal(i.cpu().numpy(), piers["top20_ids"])
    assert i[0].cpu().tolist() == kat["survey"]["fixture_top20_row0"]
    assert np.allclose(s.cpu().numpy(), piers["top20_dist"], rtol=1e-12, atol=1e-13)


def test_tc_search_scaled_and_clustered_data(cuda):
    """Norm spread and heavy clustering stress the error bound and the survivor lists."""
    from quantum_rag_b200 import api
    rng = np.random.RandomState(9)
    centers = rng.standard_normal((8, 128)).astype(np.float32)
    X = (centers[rng.randint(0, 8, 40000)] + 0.05 * rng.standard_normal((40000, 128))).astype(np.float32)
    X *= rng.uniform(0.1, 30.0, size=(40000, 1)).astype(np.float32)
    Q = (centers[rng.randint(0, 8, 50)] + 0.05 * rng.standard_normal((50, 128))).astype(np.float32)
    for name, _ in METRICS:
        _compare_with_exact(api, Q, X, 64, name, allow_fallback=True)


def test_tc_search_k_larger_than_shard(cuda):
    from quantum_rag_b200 import api
    rng = np.random.RandomState(1)
    X = rng.standard_normal((40, 32)).astype(np.float32)
    Q = rng.standard_normal((3, 32)).astype(np.float32)
    _compare_with_exact(api, Q, X, 64, "ip")


def test_adversarial_bf16_rounding_is_covered_by_the_bound(cuda):
    """Operands built so that bf16 rounding moves the approximate scores as far as it can: the query's first 32
    components and the A documents sit just below a rounding midpoint (round DOWN by 2^-8 relative), the other 32
    components and the B documents just above it (round UP).  Exactly, the 40 A documents beat the 200 B documents
    (32.274 vs 32.258); in bf16 every B document beats every A document by 0.47.  A filter margin from an assumed
    2^-9 per operand (0.36 here) drops the true top-40; the measured bound (0.69) must keep them."""
    import torch
    from quantum_rag_b200 import api
    D, N, k = 72, 4096, 40
    d = 2.0 ** -13
    q = np.zeros(D, np.float32)
    q[:32], q[32:64], q[64] = 1 + 2.0 ** -8 - d, 1 + 2.0 ** -8 + d, 0.25
    xa = np.zeros(D, np.float32)
    xa[:32], xa[64] = 1 + 2.0 ** -8 - d, 0.125
    xb = np.zeros(D, np.float32)
    xb[32:64] = 1 + 2.0 ** -8 + d
    X = (np.random.RandomState(0).standard_normal((N, D)) * 0.01).astype(np.float32)
    ia, ib = np.arange(100, 140), np.arange(1000, 1200)
    X[ia], X[ib] = xa, xb
    Q = np.stack([q, q * 2.0, q])
    # the construction really is adversarial: in bf16 arithmetic the B documents win
    bf = lambda a: torch.from_numpy(a).bfloat16().float().numpy().astype(np.float64)     # noqa: E731
    approx = bf(X) @ bf(q)
    assert approx[ib].min() - approx[ia].max() > 0.45
    for metric in ("ip", "l2"):
        index = api.FlatIndexTC(X, metric)
        s, i = index.search(Q, k)
        es, ei = api.search_topk(Q, X, k, metric)
        assert index.last_fallback == 0
        assert torch.equal(i, ei) and torch.equal(s, es), metric
    s, i = api.FlatIndexTC(X, "ip").search(Q, k)
    assert i[0].cpu().tolist() == ia.tolist() and i[1].cpu().tolist() == ia.tolist()   # ties in id order


def test_config3_full_size_properties(cuda):
    """BASELINE config 3 (1M x 384 unit rows, top-100) through size-independent properties."""
    import torch
    from quantum_rag_b200 import api
    g = torch.Generator(device="cuda").manual_seed(1234 + 3)
    N, D, nq, k = 1_000_000, 384, 1024, 100
    X = torch.nn.functional.normalize(torch.randn(N, D, generator=g, device="cuda"), dim=1)
    Q = torch.nn.functional.normalize(torch.randn(nq, D, generator=g, device="cuda"), dim=1)
    planted = torch.arange(nq, device="cuda") * 971 + 7
    osub = [0, 341, 682, 1023]                           # checked against the NumPy oracle below
    twice = torch.ones(nq, dtype=torch.bool, device="cuda")
    twice[osub] = False                                  # (no exact ties there: a BLAS product need not tie bit for bit)
    X[planted] = Q                                       # every query has itself in the corpus
    X[planted[twice] + 1] = Q[twice]                     # ... twice: exact tie, smaller id first
    index = api.FlatIndexTC(X, "cosine")
    s, i = index.search(Q, k)
    assert index.last_fallback == 0
    assert torch.equal(i[:, 0], planted) and torch.equal(i[twice, 1], planted[twice] + 1)
    assert torch.all(s[twice, 0] == s[twice, 1]) and torch.allclose(s[:, 0], torch.ones_like(s[:, 0]), atol=1e-6)
    assert torch.all(s[:, :-1] >= s[:, 1:])              # sorted
    # a subset of queries against the exact CUDA-core search: identical ids and bits
    sub = torch.arange(0, nq, 37, device="cuda")
    es, ei = api.search_topk(Q[sub], X, k, "cosine")
    assert torch.equal(i[sub], ei) and torch.equal(s[sub], es)
    # ... and four queries against the NumPy oracle over the whole 1M-row corpus (chunked): identical ids,
    # scores within fp64 rounding of the oracle's (different summation order)
    rs, ri = osr.exact_search_chunked(Q[osub].cpu().numpy(), X.cpu().numpy(), k, osr.METRIC_COSINE)
    assert np.array_equal(i[osub].cpu().numpy(), ri)
    assert np.allclose(s[osub].cpu().numpy(), rs, rtol=1e-12, atol=1e-15)
    # sharding: merging two half-corpus searches reproduces the full search
    h = N // 2
    a = api.FlatIndexTC(X[:h], "cosine", id_base=0).search(Q, k)
    b = api.FlatIndexTC(X[h:], "cosine", id_base=h).search(Q, k)
    ms, mi = api.topk_merge(torch.stack([a[0], b[0]]), torch.stack([a[1], b[1]]), k, "cosine")
    assert torch.equal(mi, i) and torch.equal(ms, s)


def test_flat_index_file_roundtrip_and_search(cuda, piers, tmp_path):
    """(f)1: an IxF2 file as the reference's ingest tool writes it -> searchable device index, faiss-style (D, I)."""
    from quantum_rag_b200 import index as qidx
    x = piers["vectors"]
    path = tmp_path / "piers.faiss"
    path.write_bytes(qidx.dump_ixf(x, qidx.FAISS_METRIC_L2))
    labels = [f"row{i}" for i in range(x.shape[0])]
    import pickle
    (tmp_path / "piers.pkl").write_bytes(pickle.dumps(labels))
    idx = qidx.FlatIndex.read(str(path), str(tmp_path / "piers.pkl"))
    assert (idx.d, idx.ntotal, idx.metric) == (1536, 119, "l2")
    D, I = idx.search(x[:5], 20)
    assert np.array_equal(I.cpu().numpy(), piers["top20_ids"][:5])
    assert np.allclose(D.cpu().numpy(), piers["top20_dist"][:5], rtol=1e-12, atol=1e-13)
    _, names = idx.search_labels(x[:1], 3)
    assert names[0] == [labels[i] for i in piers["top20_ids"][0, :3]]
    idx.write(str(tmp_path / "again.faiss"))
    assert (tmp_path / "again.faiss").read_bytes() == path.read_bytes()
    # write() left the JSON side-car next to the index; read() picks it up without being told (never the pickle)
    again = qidx.FlatIndex.read(str(tmp_path / "again.faiss"))
    assert again.labels == labels and (tmp_path / "again_labels.json").exists()
    assert qidx.FlatIndex.read(str(path)).labels is None               # piers.faiss has only a pickle beside it


@pytest.mark.parametrize("D", [384, 1536])
def test_tensor_core_accumulation_error_is_inside_the_bound(cuda, D):
    """The third term of the filter's error bound (search_tc.cu tc_eps: g = Kp * 2.4e-7 of |a||b| for the fp32
    accumulation inside the tensor core) MEASURED: operands exactly representable in bf16, so that operand rounding
    contributes nothing and |TMEM score - fp64 dot of the same operands| is the accumulation error alone.
    Data: random signs (cancellation), all-positive (the accumulator grows, every add rounds at its largest ulp),
    and magnitudes spread over 2^-6 .. 2^6."""
    import torch
    from quantum_rag_b200 import api
    g = torch.Generator(device="cuda").manual_seed(77 + D)
    N, nq = 8192, 256
    bf = lambda t: t.bfloat16().float()                                            # noqa: E731
    X = torch.randn(N, D, generator=g, device="cuda")
    X[N // 4:N // 2] = X[N // 4:N // 2].abs() + 1.0                                # all positive, similar size
    X[N // 2:3 * N // 4] *= torch.exp2(torch.randint(-6, 7, (N // 4, D), generator=g, device="cuda").float())
    Q = torch.randn(nq, D, generator=g, device="cuda")
    Q[nq // 2:] = Q[nq // 2:].abs() + 1.0
    X, Q = bf(X), bf(Q)
    Kp = (D + 15) // 16 * 16
    worst = 0.0
    for metric in ("ip",):
        index = api.FlatIndexTC(X, metric)
        assert torch.equal(index.Xb[:, :D].float(), X)                             # the shadow IS the data: no rounding
        S = index.approx_scores(Q).double()
        exact = Q.double() @ X.double().T
        scale = Q.double().norm(dim=1)[:, None] * X.double().norm(dim=1)[None, :]
        ratio = ((S - exact).abs() / scale).max().item()
        worst = max(worst, ratio)
    # measured on B200: ~1e-7 * small factor; the bound's constant is Kp * 2.4e-7
    assert worst <= Kp * 2.4e-7, f"accumulation error {worst:.3e} |a||b| exceeds the bound's {Kp * 2.4e-7:.3e}"
    assert worst > 0.0                                                             # the probe really measures something
    print(f"D={D}: measured max accumulation error = {worst:.3e} |a||b|; bound term = {Kp * 2.4e-7:.3e}")


def test_approx_scores_within_eps_of_exact(cuda):
    """The whole bound, measured: |approximate - exact| <= eps for every (query, document) pair on ordinary data,
    with eps recomputed here from the bound's published formula and the measured rounding-error norms."""
    import torch
    from quantum_rag_b200 import api
    g = torch.Generator(device="cuda").manual_seed(5)
    N, nq, D = 20000, 64, 384
    X = torch.randn(N, D, generator=g, device="cuda") * torch.rand(N, 1, generator=g, device="cuda") * 3
    Q = torch.randn(nq, D, generator=g, device="cuda")
    index = api.FlatIndexTC(X, "ip")
    S = index.approx_scores(Q).double()
    exact = Q.double() @ X.double().T
    qb = Q.bfloat16().float()
    qn, qe = Q.double().norm(dim=1), (Q - qb).double().norm(dim=1)
    xmax, xe = float(index.aux[0]), float(index.aux[1])
    gk = 384 * 2.4e-7
    eps = qe * (xmax + xe) + qn * xe + gk * (qn + qe) * (xmax + xe)
    err = (S - exact).abs().max(dim=1).values
    assert torch.all(err <= eps), (err / eps).max().item()
    assert (err / eps).max().item() > 0.01                                         # and it is not vacuous by orders of magnitude


def test_cosine_fp16_operands_bound_and_flush_to_zero(cuda):
    """Cosine runs the filter GEMM on fp16 operands (both sides normalised).  (a) |approximate - exact cosine| <= eps
    for every pair, eps recomputed from the published formula and the measured rounding-error norms; (b) rows whose
    normalised components fall below fp16's smallest normal (flushed to zero in the shadow, counted in the measured
    error) and queries of extreme scale (1e-30 ... 1e+30: the operand is q/|q|) still give the exact search's bits."""
    import torch
    from quantum_rag_b200 import api
    g = torch.Generator(device="cuda").manual_seed(8)
    N, nq, D = 20000, 64, 384
    X = torch.randn(N, D, generator=g, device="cuda") * torch.rand(N, 1, generator=g, device="cuda") * 3
    Q = torch.randn(nq, D, generator=g, device="cuda")
    index = api.FlatIndexTC(X, "cosine")
    assert index.Xb.dtype == torch.float16
    S = index.approx_scores(Q).double()
    Xd, Qd = X.double(), Q.double()
    exact = (Qd @ Xd.T) / (Qd.norm(dim=1)[:, None] * Xd.norm(dim=1)[None, :])
    a = (Qd / Qd.norm(dim=1, keepdim=True)).float()                        # the query operand before the fp16 rounding
    back = a.half().float()
    back[a.abs() < 2.0 ** -14] = 0.0
    qn, qe = a.double().norm(dim=1), (a - back).double().norm(dim=1)
    xe = float(index.aux[1])
    gk, xt = 384 * 2.4e-7, 1.0000002 + xe
    eps = (qe + 2.4e-7 * qn) * xt + qn * (xe + 2.4e-7) + gk * (qn + qe) * xt
    err = (S - exact).abs().max(dim=1).values
    assert torch.all(err <= eps), (err / eps).max().item()
    assert xe < 1e-3                                                       # fp16: ~2e-4 |x| (bf16 would measure ~1.7e-3)
    # (b) extreme dynamic range inside a row, extreme scale across queries
    rng = np.random.RandomState(12)
    Xh = (rng.standard_normal((30000, 128)) * 1e-6).astype(np.float32)
    Xh[np.arange(30000), rng.randint(0, 128, 30000)] += rng.choice([-1.0, 1.0], 30000).astype(np.float32)
    Qh = (rng.standard_normal((40, 128))).astype(np.float32)
    Qh[:10] *= 1e30
    Qh[10:20] *= 1e-30
    Qh[20] = 0.0
    _compare_with_exact(api, Qh, Xh, 50, "cosine", allow_fallback=True)


@pytest.mark.parametrize("nq", [4, 200])
def test_clustered_corpus_does_not_fall_back(cuda, nq):
    """near-duplicates stored in ADJACENT rows put a query's whole top-k into one 256-row tile, i.e. into
    the four survivor segments of one CTA.  The segments are sized for that (256 slots each where memory allows), so a
    clustered corpus is searched by the tensor-core path without the exact rerun (which re-reads the corpus per query)."""
    import torch
    from quantum_rag_b200 import api
    rng = np.random.RandomState(17)
    N, D, k = 60000, 384, 100
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    for q in range(min(nq, 8)):                       # 230 near-copies of query q in adjacent rows of one tile
        base = 256 * (10 + 7 * q) + 5
        X[base:base + 230] = Q[q] + 0.02 * rng.standard_normal((230, D)).astype(np.float32)
    for name in ("cosine", "ip", "l2"):
        index = _compare_with_exact(api, Q, X, k, name)          # asserts last_fallback == 0 and identity with the exact search
        assert index.last_fallback == 0
