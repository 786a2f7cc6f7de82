"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle and the committed
golden fixtures.  Integer results (ranks, hits, top-k indices) must be bit-exact; floating-point
results carry the tolerance BASELINE.json's north_star states (1e-4 rel fp32, 2e-2 rel bf16).
Nothing here reads /root/reference (it does not exist on the GPU box)."""
import ctypes
import math

import numpy as np
import pytest
import torch

from oracle import vtc_oracle as O
from vtc_b200.synthetic import make_batch_pair, make_cam_inputs, make_retrieval_pair

pytestmark = pytest.mark.gpu

METRICS = {"l2": O.METRIC_L2, "dot": O.METRIC_DOT}


def _np(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------ H1
def test_normalize_matches_reference_semantics(cuda_dev):
    from vtc_b200 import ops

    x = torch.randn(777, 512)
    x[5] = 0.0  # zero row -> NaN, no eps (model/model.py:26-27)
    y = ops.normalize(x.to(cuda_dev))
    want = O.normalize(x)
    assert torch.isnan(y[5]).all()
    ok = torch.ones(777, dtype=torch.bool)
    ok[5] = False
    np.testing.assert_allclose(_np(y)[ok.numpy()], want.numpy()[ok.numpy()], rtol=2e-6, atol=1e-7)
    inv, sq = ops.row_norms(x.to(cuda_dev))
    np.testing.assert_allclose(_np(sq), (x.double() ** 2).sum(1).numpy(), rtol=1e-5)
    y3 = ops.normalize(torch.randn(6, 11, 96, device=cuda_dev))
    np.testing.assert_allclose(_np(y3.norm(dim=-1)), 1.0, rtol=1e-5)


# ------------------------------------------------------------------------------------------ H2
@pytest.mark.parametrize("N,M,D", [(128, 256, 64), (300, 700, 512), (1000, 1000, 512),
                                   (257, 513, 768), (64, 100, 100), (130, 40, 1024)])
@pytest.mark.parametrize("precision", ["bf16", "exact"])
def test_sim_matrix_tensor_core_gemm(cuda_dev, N, M, D, precision):
    """The tcgen05 GEMM core (EPI_STORE) against fp64; also the evidence for the guard band:
    |tc - exact| / (|a||b|) must stay well inside the library's guard_rel (csrc/api.cu
    guard_rel_for: (K'/16 + 8) * 2^-24, plus 4.62e-5 for the bf16x3 split)."""
    from vtc_b200 import ops

    g = torch.Generator().manual_seed(N * 7 + M)
    a = torch.randn(N, D, generator=g)
    b = torch.randn(M, D, generator=g)
    a = a / a.norm(dim=-1, keepdim=True) * (0.5 + torch.rand(N, 1, generator=g))
    b = b / b.norm(dim=-1, keepdim=True)
    scale = 3.25
    out = _np(ops.sim_matrix(a.to(cuda_dev), b.to(cuda_dev), scale, precision))
    a_ref, b_ref = a, b
    if precision == "bf16":
        a_ref, b_ref = a.bfloat16().float(), b.bfloat16().float()
    want = scale * (a_ref.double() @ b_ref.double().t()).numpy()
    norms = (a_ref.norm(dim=-1, keepdim=True) * b_ref.norm(dim=-1, keepdim=True).t()).double().numpy()
    rel = np.abs(out - want) / (scale * norms)
    kp = -(-(D if precision == "bf16" else 3 * D) // 64) * 64
    guard = (kp // 16 + 8) * 2.0 ** -24 + (0.0 if precision == "bf16" else 4.62e-5)
    print(f"\n[guard-band evidence] {precision} N={N} M={M} D={D}: max rel err {rel.max():.3e} "
          f"(guard {guard:.3e}, margin x{guard / max(rel.max(), 1e-30):.1f})")
    assert rel.max() < guard / 3
    # against fp32 inputs the north_star tolerances hold
    full = scale * (a.double() @ b.double().t()).numpy()
    tol = 2e-2 if precision == "bf16" else 1e-4
    assert np.abs(out - full).max() <= tol * np.abs(full).max()


# ------------------------------------------------------------------------------------- R1 / R3
def _oracle_ranks(Q, G, metric, precision, gt=None):
    if precision == "bf16":
        Q, G = O.bf16_round(Q), O.bf16_round(G)
    return O.rank0_exact(Q, G, gt=gt, metric=METRICS[metric])


@pytest.mark.parametrize("precision", ["brute", "exact", "bf16"])
@pytest.mark.parametrize("metric", ["l2", "dot"])
@pytest.mark.parametrize("N,M,D,sigma", [(1000, 1000, 512, 6.0), (333, 1201, 96, 2.0),
                                         (129, 257, 768, 7.0), (5, 3000, 64, 1.5)])
def test_rank_bit_exact(cuda_dev, precision, metric, N, M, D, sigma):
    from vtc_b200 import ops

    T, V = make_retrieval_pair(min(N, M), M, D, sigma=sigma, seed=N + M)
    if N > M:
        T = torch.cat([T, T[: N - M]])
    T = T[:N].contiguous()
    rank0, gts = ops.sim_rank(T.to(cuda_dev), V.to(cuda_dev), metric=metric, precision=precision)
    hits, medr = ops.rank_finalize(rank0, gts, M, [1, 5, 10])
    want = _oracle_ranks(T, V, metric, precision)
    np.testing.assert_array_equal(_np(rank0), want)
    np.testing.assert_array_equal(_np(hits), [np.sum(want < k) for k in (1, 5, 10)])
    assert _np(medr)[0] == O.medr(want)


@pytest.mark.parametrize("precision", ["exact", "bf16"])
@pytest.mark.parametrize("D", [64, 512])
def test_rank_structured_embeddings_ties_stay_exact(cuda_dev, precision, D):
    """Constant-magnitude sign embeddings whose elements sit just below a bf16 rounding midpoint
    (the inputs on which the bf16x3 split loses the most: 2.4e-5 of |q||x|, tests/test_guard_band.py)
    with duplicated gallery rows, so that exactly tied columns abound.  A tie must reach the fp64
    re-check (index tie-break), which it only does if the guard band really bounds the tensor-core
    error -- the round-1 constant (1.2e-5) did not."""
    from vtc_b200 import ops

    rng = np.random.default_rng(D)
    base = np.sign(rng.standard_normal(D)).astype(np.float32)
    M, N = 900, 400
    flip = rng.random((M, D)) < 0.1
    V = (np.where(flip, -base, base) * np.float32(1.00385 / 32)).astype(np.float32)
    V[300:600] = V[0:300]          # every row of [0, 300) has an exact duplicate further down
    V[700] = V[5]
    T = V[:N].copy()               # queries identical to their ground truth: d(t, gt) ties with the copies
    T[350:] = V[350:N] * np.float32(0.5)
    Vt, Tt = torch.from_numpy(V), torch.from_numpy(T)
    for metric in ("l2", "dot"):
        rank0, gts = ops.sim_rank(Tt.to(cuda_dev), Vt.to(cuda_dev), metric=metric, precision=precision)
        ops.rank_finalize(rank0, gts, M, [1])
        np.testing.assert_array_equal(_np(rank0), _oracle_ranks(Tt, Vt, metric, precision))


@pytest.mark.parametrize("precision", ["brute", "exact", "bf16"])
def test_rank_explicit_gt_and_mixed_difficulty(cuda_dev, precision):
    from vtc_b200 import ops

    T, V = make_retrieval_pair(700, 900, 512, seed=11, mixed=True)
    perm = torch.randperm(900, generator=torch.Generator().manual_seed(3))
    Vp = V[perm].contiguous()  # gallery shuffled: gt(t) = position of row t
    inv = torch.empty(900, dtype=torch.int64)
    inv[perm] = torch.arange(900)
    gt = inv[:700].contiguous()
    rank0, gts = ops.sim_rank(T.to(cuda_dev), Vp.to(cuda_dev), gt=gt.to(cuda_dev), precision=precision)
    ops.rank_finalize(rank0, gts, 900, [1])
    want = _oracle_ranks(T, Vp, "l2", precision, gt=gt.numpy())
    np.testing.assert_array_equal(_np(rank0), want)
    assert want.max() > 50  # the mixed set really spreads ranks


@pytest.mark.parametrize("precision", ["brute", "exact", "bf16"])
def test_rank_adversarial_golden(cuda_dev, golden, precision):
    """Duplicated gallery rows (exact ties), a zero row, a non-unit row, a NaN query."""
    from vtc_b200 import ops

    g = golden("retrieval_small.npz")
    Q, G = torch.from_numpy(g["queries"]), torch.from_numpy(g["gallery"])
    for metric, key in (("l2", "rank0"), ("dot", "rank0_dot")):
        rank0, gts = ops.sim_rank(Q.to(cuda_dev), G.to(cuda_dev), metric=metric, precision=precision)
        ops.rank_finalize(rank0, gts, G.shape[0], [1])
        if precision == "bf16":
            want = _oracle_ranks(Q, G, metric, precision)
        else:
            want = g[key]
        np.testing.assert_array_equal(_np(rank0), want)


def test_rank_all_duplicates_and_overflow_fallback(cuda_dev, lib):
    """Every pair ties -> every pair is ambiguous: exercises the re-check list and, with a
    workspace too small for the list, the brute-force fallback."""
    from vtc_b200 import _ffi, ops

    row = torch.randn(1, 128)
    G = row.repeat(600, 1).contiguous()
    Q = row.repeat(400, 1).contiguous()
    want = O.rank0_exact(Q, G)
    np.testing.assert_array_equal(want, np.arange(400))  # ties broken by index
    for precision in ("exact", "bf16"):
        rank0, _ = ops.sim_rank(Q.to(cuda_dev), G.to(cuda_dev), precision=precision)
        np.testing.assert_array_equal(_np(rank0), want)
    # overflow: hand the C ABI a workspace that leaves room for ~2k pairs (240k are ambiguous)
    q, g = Q.to(cuda_dev), G.to(cuda_dev)
    need = lib.vtc_workspace_bytes(_ffi.OP_SIM_RANK, 400, 600, 128, _ffi.PREC_EXACT)
    small = need - (1 << 20) * 8 + 2048 * 8
    ws = torch.empty(small, dtype=torch.uint8, device=cuda_dev)
    rank0 = torch.full((400,), -7, dtype=torch.int32, device=cuda_dev)
    rc = lib.vtc_sim_rank(q.data_ptr(), g.data_ptr(), 400, 600, 128, _ffi.F32, None, 0, 0,
                          _ffi.METRIC_L2, _ffi.PREC_EXACT, None, None, 0, rank0.data_ptr(),
                          ws.data_ptr(), ws.numel(),
                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.vtc_strerror(rc)
    np.testing.assert_array_equal(_np(rank0), want)


@pytest.mark.parametrize("precision", ["brute", "exact", "bf16"])
def test_rank_and_topk_degenerate_shapes(cuda_dev, precision):
    """Empty / single-row / odd-D / M < k inputs (the ragged edges of every tile)."""
    from vtc_b200 import ops

    for (N, M, D) in ((1, 1, 8), (3, 2, 5), (1, 300, 37), (130, 1, 64), (257, 513, 1)):
        g = torch.Generator().manual_seed(N * 1000 + M)
        G = torch.randn(M, D, generator=g)
        Q = torch.randn(N, D, generator=g)
        gt = torch.randint(0, M, (N,), generator=g)
        rank0, gts = ops.sim_rank(Q.to(cuda_dev), G.to(cuda_dev), gt=gt.to(cuda_dev), precision=precision)
        hits, medr = ops.rank_finalize(rank0, gts, M, [1, 5])
        Qo, Go = (O.bf16_round(Q), O.bf16_round(G)) if precision == "bf16" else (Q, G)
        want = O.rank0_exact(Qo, Go, gt=gt.numpy())
        np.testing.assert_array_equal(_np(rank0), want)
        assert _np(medr)[0] == O.medr(want)
        vals, idx = ops.sim_topk(Q.to(cuda_dev), G.to(cuda_dev), 5, precision=precision)
        np.testing.assert_array_equal(_np(idx), O.topk_exact(Qo, Go, 5)[1])
    # no queries at all: nothing to do, empty results
    Q0 = torch.empty(0, 16, device=cuda_dev)
    G0 = torch.randn(9, 16, device=cuda_dev)
    r, gs = ops.sim_rank(Q0, G0, precision=precision)
    assert r.shape == (0,) and gs.shape == (0,)
    v, i = ops.sim_topk(Q0, G0, 3, precision=precision)
    assert v.shape == (0, 3) and i.shape == (0, 3)


def test_topk_full_size_1M_gallery(cuda_dev):
    """BASELINE config 5 at full gallery size on one GPU (1M x 512, bf16): the tensor-core path
    against the fp64 brute-force kernel on all rows and against the CPU oracle on a few."""
    from vtc_b200 import ops

    M, N, D, k = 1_000_000, 48, 512, 11
    T, V = make_retrieval_pair(N, M, D, seed=1023)
    q, g = T.to(cuda_dev), V.to(cuda_dev)
    vals, idx = ops.sim_topk(q, g, k, precision="bf16")
    qb, gb = q.bfloat16().float(), g.bfloat16().float()
    bv, bi = ops.sim_topk(qb, gb, k, precision="brute")
    np.testing.assert_array_equal(_np(idx), _np(bi))
    np.testing.assert_allclose(_np(vals), _np(bv), rtol=1e-5, atol=1e-6)
    wi = O.topk_exact(O.bf16_round(T[:8]), O.bf16_round(V), k)[1]
    np.testing.assert_array_equal(_np(idx[:8]), wi)
    assert (_np(idx[:, 0]) == np.arange(N)).mean() > 0.05  # queries are noisy copies of rows 0..N-1


def test_rank_chunked_gallery_is_additive(cuda_dev):
    """rank counts add over gallery chunks (the multi-GPU decomposition)."""
    from vtc_b200 import ops

    T, V = make_retrieval_pair(500, 2000, 256, sigma=4.0, seed=21)
    q, g = T.to(cuda_dev), V.to(cuda_dev)
    for precision in ("exact", "bf16", "brute"):
        whole, gts = ops.sim_rank(q, g, precision=precision)
        acc = torch.zeros(500, dtype=torch.int32, device=cuda_dev)
        gs = torch.full((500,), float("nan"), dtype=torch.float64, device=cuda_dev)
        bounds = [0, 700, 701, 1500, 2000]
        for s, e in zip(bounds[:-1], bounds[1:]):
            part = ops.gt_scores(q, g[s:e], col_offset=s, precision=precision)
            gs = torch.where(torch.isnan(gs), part, gs)
        np.testing.assert_array_equal(_np(gs), _np(gts))
        for s, e in zip(bounds[:-1], bounds[1:]):
            ops.sim_rank(q, g[s:e].contiguous(), col_offset=s, precision=precision, gt_score=gs,
                         rank0=acc, accumulate=True)
        np.testing.assert_array_equal(_np(acc), _np(whole))


def test_rank_10k_config3(cuda_dev):
    """BASELINE config 3: 10k x 10k x 512, fused similarity + rank on one GPU."""
    from vtc_b200 import ops

    T, V = make_retrieval_pair(10000, 10000, 512, seed=1023)
    q, g = T.to(cuda_dev), V.to(cuda_dev)
    want = {"exact": O.rank0_exact(T, V), "bf16": _oracle_ranks(T, V, "l2", "bf16")}
    for precision in ("exact", "bf16"):
        rank0, gts = ops.sim_rank(q, g, precision=precision)
        hits, medr = ops.rank_finalize(rank0, gts, 10000, [1, 5, 10])
        np.testing.assert_array_equal(_np(rank0), want[precision])
        assert _np(medr)[0] == O.medr(want[precision])
    # SURVEY.md §8d calibration for this seed: R@1/5/10 = 0.460/0.668/0.738, MedR 2
    r = O.recall_from_ranks(want["exact"], [1, 5, 10])
    assert [round(x, 3) for _, x in r] == [0.46, 0.668, 0.738]


def test_full_size_properties_100k(cuda_dev):
    """North-star size (100k x 100k x 512): the oracle cannot finish this in seconds, so check
    size-independent properties: (a) tensor-core ranks == fp64 brute force on a slice of the
    queries; (b) ranks are additive over gallery halves; (c) hits == count of rank0 < k."""
    from vtc_b200 import ops

    N = M = 100_000
    T, V = make_retrieval_pair(N, M, 512, seed=1023)
    q, g = T.to(cuda_dev), V.to(cuda_dev)
    rank_bf16, gts = ops.sim_rank(q, g, precision="bf16")
    sl = slice(50_000, 50_256)
    qb, gb = q[sl].bfloat16().float().contiguous(), g.bfloat16().float()
    brute, _ = ops.sim_rank(qb, gb, row_offset=50_000, precision="brute")
    np.testing.assert_array_equal(_np(rank_bf16[sl]), _np(brute))
    # the same slice against the CPU oracle (256 x 100k is ~1 s)
    want = O.rank0_exact(O.bf16_round(T[sl]), O.bf16_round(V), row_offset=50_000)
    np.testing.assert_array_equal(_np(brute), want)
    acc = torch.zeros(N, dtype=torch.int32, device=cuda_dev)
    for s, e in ((0, 50_000), (50_000, M)):
        ops.sim_rank(q, g[s:e].contiguous(), col_offset=s, precision="bf16", gt_score=gts, rank0=acc,
                     accumulate=True)
    np.testing.assert_array_equal(_np(acc), _np(rank_bf16))
    hits, medr = ops.rank_finalize(rank_bf16, gts, M, [1, 5, 10])
    r = _np(rank_bf16)
    np.testing.assert_array_equal(_np(hits), [np.sum(r < k) for k in (1, 5, 10)])
    assert _np(medr)[0] == O.medr(r)


@pytest.mark.parametrize("D", [64, 192, 100])
def test_bf16_inputs_used_in_place(cuda_dev, D):
    """bf16 rows of whole 128-byte atoms are the tensor-core operands themselves (no prep copy);
    other widths go through the prep kernel.  Same ranks / top-k either way, also on row slices
    (offset base pointers) as the pipelined staging and the multi-GPU gather hand them in."""
    from vtc_b200 import ops

    T, V = make_retrieval_pair(700, 900, D, sigma=3.0, seed=D)
    Tq, Vq = O.bf16_round(T), O.bf16_round(V)
    q16, g16 = T.to(cuda_dev).bfloat16(), V.to(cuda_dev).bfloat16()
    for prec in ("bf16", "exact"):
        r, _ = ops.sim_rank(q16, g16, precision=prec)
        np.testing.assert_array_equal(_np(r), O.rank0_exact(Tq, Vq))
        _, idx = ops.sim_topk(q16[:128], g16, 11, precision=prec)
        np.testing.assert_array_equal(_np(idx), O.topk_exact(Tq[:128], Vq, 11)[1])
    # slices: queries 300.., gallery rows 256..900 (ground truth shifted by the offsets)
    r, _ = ops.sim_rank(q16[300:], g16[256:], row_offset=300, col_offset=256, precision="bf16")
    np.testing.assert_array_equal(_np(r), O.rank0_exact(Tq[300:], Vq[256:], gt=np.arange(300, 700) - 256))


# ------------------------------------------------------------- drop-in call sites (R1, R2, R4)
def test_recall_at_k_and_compute_recall_golden(cuda_dev, golden):
    """The reference-facing calls on config-1 inputs against the fixtures produced by the
    reference's own RecallAtK.compute / compute_recall."""
    from vtc_b200.evaluation.retrieval_evaluation import compute_recall, compute_recall_full
    from vtc_b200.model.metric import RecallAtK

    g = golden("retrieval_c1.npz")
    for mixed, key, rk in ((False, "df_values", ""), (True, "df_values_mixed", "_mixed")):
        T, V = make_retrieval_pair(1000, 1000, 512, sigma=None if mixed else 6.0, seed=1023,
                                   mixed=mixed)
        df = compute_recall(V, T.unsqueeze(1))  # CPU tensors in, like the reference call site
        assert list(df.index) == list(g["df_index"])
        assert list(df.columns) == list(g["df_columns"])
        np.testing.assert_array_equal(df.values, g[key])
        full = compute_recall_full(V, T.unsqueeze(1))
        np.testing.assert_array_equal(_np(full["t2v"]["rank0"]), g["rank_t2v" + rk])
        np.testing.assert_array_equal(_np(full["v2t"]["rank0"]), g["rank_v2t" + rk])
        # numpy in, list of (k, float) out -- model/metric.py:137-161
        m = RecallAtK("videos", "titles", [1, 5, 10])
        got = m.compute(V.numpy(), T.numpy())
        assert [k for k, _ in got] == [1, 5, 10]
        np.testing.assert_array_equal(np.array([r for _, r in got]) * 100.0, g[key][:, 1])


def _reference_hit_loop(I, k_vals, num_samples):
    """model/metric.py:148-160, the loop the reference runs over the index lists faiss returns."""
    return [(k, sum(1 for target, rp in enumerate(I) if target in rp[:k]) / num_samples)
            for k in k_vals]


def test_faiss_compat_index_on_reference_fixtures(cuda_dev, golden):
    """The reference's faiss call site (model/metric.py:139-146) bound to vtc_b200.faiss_compat:
    GpuIndexFlatL2.add / search(k = max(k_vals) + 1) + the reference's own hit loop reproduce the
    fixtures the reference's RecallAtK.compute generated; numpy in -> numpy out like faiss."""
    from vtc_b200 import faiss_compat as faiss

    cfg = faiss.GpuIndexFlatConfig()
    cfg.useFloat16 = False                       # model/metric.py:112-113
    cfg.device = cuda_dev.index or 0             # :127
    g = golden("retrieval_c1.npz")
    for mixed, key in ((False, "df_values"), (True, "df_values_mixed")):
        T, V = make_retrieval_pair(1000, 1000, 512, sigma=None if mixed else 6.0, seed=1023,
                                   mixed=mixed)
        for a, b, col in ((V, T, 1), (T, V, 0)):   # gallery a, queries b; both retrieval directions
            index = faiss.GpuIndexFlatL2(faiss.StandardGpuResources(), a.shape[1], cfg)
            index.add(a.numpy())
            assert index.ntotal == 1000 and index.d == 512
            D, I = index.search(b.numpy(), 11)
            assert isinstance(I, np.ndarray) and I.dtype == np.int64 and I.shape == (1000, 11)
            assert D.dtype == np.float32 and (np.diff(D, axis=1) >= 0).all()
            got = _reference_hit_loop(I, [1, 5, 10], 1000)
            np.testing.assert_array_equal(np.array([r for _, r in got]) * 100.0, g[key][:, col])
            wv, wi = O.topk_exact(b, a, 11)
            np.testing.assert_array_equal(I, wi)
            # faiss reports the full squared distance ||q||^2 + ||x||^2 - 2 q.x
            np.testing.assert_allclose(D, wv + (b.numpy().astype(np.float64) ** 2).sum(1, keepdims=True),
                                       rtol=1e-5, atol=1e-6)
    # the adversarial fixture: exact ties, zero / non-unit rows, a NaN query, +-inf rows
    s = golden("retrieval_small.npz")
    Q, G = s["queries"], s["gallery"]
    index = faiss.GpuIndexFlatL2(faiss.StandardGpuResources(), 64, cfg)
    index.add(G[:100])
    index.add(G[100:])                            # incremental adds concatenate
    _, I = index.search(Q, 11)
    got = np.array([r for _, r in _reference_hit_loop(I, list(s["k_vals"]), G.shape[0])])
    np.testing.assert_array_equal(got, s["recall_q_from_g"])
    ok = ~np.isnan(Q).any(1)
    np.testing.assert_array_equal(I[ok], O.topk_exact(Q, G, 11)[1][ok])
    # device tensors in -> device tensors out; the bf16 mode is exact on the rounded rows; k > ntotal
    cfg16 = faiss.GpuIndexFlatConfig()
    cfg16.useFloat16 = True
    T, V = make_retrieval_pair(300, 2000, 256, sigma=3.0, seed=5)
    index = faiss.GpuIndexFlatL2(None, 256, cfg16)
    index.add(V.to(cuda_dev))
    D, I = index.search(T.to(cuda_dev), 10)
    assert I.is_cuda
    np.testing.assert_array_equal(_np(I), O.topk_exact(O.bf16_round(T), O.bf16_round(V), 10)[1])
    index.reset()
    assert index.ntotal == 0
    index.add(V[:4].numpy())
    D, I = index.search(T[:3].numpy(), 6)
    assert (I[:, 4:] == -1).all() and np.isinf(D[:, 4:]).all() and (I[:, :4] >= 0).all()
    with pytest.raises(ValueError):
        index.search(T[:3].numpy(), 17)


def test_eval_consumer_six_keys(cuda_dev, tmp_path):
    """evaluation/eval.py:97-138: a loader of (vis, title, comments, meta) batches through a model,
    features kept on the device, six floats keyed R{k}_{title_from_im,im_from_title}."""
    import json

    from vtc_b200.evaluation.eval import evaluate, recall_summary

    T, V = make_retrieval_pair(700, 700, 128, sigma=5.0, seed=3)
    batches = [(V[s:s + 96].unsqueeze(1), T[s:s + 96].unsqueeze(1), torch.zeros(min(96, 700 - s), 1),
                {"id": torch.arange(s, min(s + 96, 700))}) for s in range(0, 700, 96)]

    def model(vis, title, comments):  # stands in for the CLIP encoders: features pass through
        return vis, title, None

    out = evaluate(model, batches, cuda_dev, save_path=str(tmp_path / "r.json"))
    want_ti = O.recall_at_k(V.numpy(), T.numpy(), [1, 5, 10])   # gallery = images, queries = titles
    want_it = O.recall_at_k(T.numpy(), V.numpy(), [1, 5, 10])
    assert list(out) == ["R1_title_from_im", "R5_title_from_im", "R10_title_from_im",
                         "R1_im_from_title", "R5_im_from_title", "R10_im_from_title"]
    assert [out[f"R{k}_title_from_im"] for k in (1, 5, 10)] == [r for _, r in want_ti]
    assert [out[f"R{k}_im_from_title"] for k in (1, 5, 10)] == [r for _, r in want_it]
    assert json.load(open(tmp_path / "r.json")) == out
    assert recall_summary(V.numpy(), T.numpy()) == out          # numpy in, like the reference


def test_eval_tail_on_device(cuda_dev):
    """R5 (evaluation/retrieval_evaluation.py:238-260): lists of per-video frame features and
    per-video caption features -> [N, D] means (not renormalised) and [N, maxcap, D] with -inf
    padding, built on the device in O(1) launches; then ranked like the reference does."""
    from vtc_b200.evaluation.retrieval_evaluation import compute_recall, eval_tail

    g = torch.Generator().manual_seed(5)
    T, V = make_retrieval_pair(300, 300, 64, sigma=2.0, seed=9)
    vids = [V[i:i + 1] + 0.01 * torch.randn(int(n), 64, generator=g)
            for i, n in enumerate(torch.randint(1, 9, (300,), generator=g))]
    caps = [T[i:i + 1].repeat(int(n), 1) for i, n in enumerate(torch.randint(1, 4, (300,), generator=g))]
    v, c = eval_tail(vids, caps, cuda_dev)
    vo, co = O.eval_tail(vids, caps)
    assert v.is_cuda and c.is_cuda
    np.testing.assert_array_equal(_np(c), co.numpy())
    np.testing.assert_allclose(_np(v), vo.numpy(), rtol=1e-6, atol=1e-7)
    # one caption per video -> the reference's compute_recall call on the tail's outputs
    v1, c1 = eval_tail(vids, [x[:1] for x in caps], cuda_dev)
    df = compute_recall(v1, c1)
    want = O.compute_recall(v1.cpu(), c1.cpu())
    np.testing.assert_array_equal(df.values, want.values)


def test_recall_pipelined_host_staging(cuda_dev):
    """Large host inputs are staged chunk by chunk on a copy stream (H2D overlaps ranking); the
    result must not depend on the chunking."""
    from vtc_b200.model.metric import RecallAtK

    T, V = make_retrieval_pair(1003, 1003, 256, sigma=4.0, seed=17)
    want = O.rank0_exact(T, V)
    m = RecallAtK("videos", "titles", [1, 5, 10])
    m.PIPELINE_MIN_BYTES = 0
    for qa, ga in ((T, V), (T.numpy(), V.numpy()), (T.pin_memory(), V.to(cuda_dev))):
        full = m.compute_full(ga, qa)
        np.testing.assert_array_equal(_np(full["rank0"]), want)
        np.testing.assert_array_equal(_np(full["hits"]), [np.sum(want < k) for k in (1, 5, 10)])
        assert _np(full["medr"])[0] == O.medr(want)


def test_recall_from_cached_embedding_files(cuda_dev, tmp_path):
    """the reference's cached-feature .pth files feed the eval kernels."""
    from vtc_b200.data import recall_from_cached

    T, V = make_retrieval_pair(500, 500, 128, sigma=3.0, seed=8)
    ids = torch.arange(1000, 1500, dtype=torch.int64)
    perm = torch.randperm(500, generator=torch.Generator().manual_seed(1))
    torch.save({"reddit_ids": ids, "embeddings": V * 3.0}, tmp_path / "v.pth")          # un-normalised
    torch.save({"reddit_ids": ids[perm], "embeddings": (T * 0.5)[perm]}, tmp_path / "t.pth")  # shuffled
    df = recall_from_cached(str(tmp_path / "v.pth"), str(tmp_path / "t.pth"), split="s", dataset_name="d")
    want = O.compute_recall(O.normalize(V * 3.0), O.normalize(T * 0.5).unsqueeze(1), split="s",
                            dataset_name="d")
    assert list(df.columns) == list(want.columns)
    np.testing.assert_array_equal(df.values, want.values)


def test_recall_at_k_update_result_protocol(cuda_dev):
    from vtc_b200.model.metric import MetricTracker, RecallAtK

    T, V = make_retrieval_pair(300, 300, 128, sigma=3.0, seed=2)
    m = RecallAtK("visual", "titles", k_vals=[1, 10])
    tr = MetricTracker(m)
    for s in range(0, 300, 50):
        tr.update(0.0, (V[s:s + 50].to(cuda_dev), T[s:s + 50].to(cuda_dev)), {})
    res = tr.result()
    assert set(res) == {"titles_from_visual-recall_at_1", "titles_from_visual-recall_at_10",
                        "visual_from_titles-recall_at_1", "visual_from_titles-recall_at_10"}
    want = dict(O.recall_at_k(V.numpy(), T.numpy(), [1, 10]))
    assert res["titles_from_visual-recall_at_1"] == want[1]
    assert res["titles_from_visual-recall_at_10"] == want[10]
    assert m.avg() is None and m.is_train is False


# ------------------------------------------------------------------------------------------ K7
@pytest.mark.parametrize("precision", ["brute", "exact", "bf16"])
@pytest.mark.parametrize("metric", ["l2", "dot"])
@pytest.mark.parametrize("k", [11, 16])
def test_topk_matches_oracle(cuda_dev, precision, metric, k):
    from vtc_b200 import ops

    T, V = make_retrieval_pair(300, 5000, 256, sigma=3.0, seed=31)
    V = V.clone()
    V[100] = V[7]
    V[4000] = V[7]  # exact ties across tiles
    V[50] *= 0.5    # non-unit row
    vals, idx = ops.sim_topk(T.to(cuda_dev), V.to(cuda_dev), k, metric=metric, precision=precision,
                             col_offset=1000)
    Tq, Vq = (O.bf16_round(T), O.bf16_round(V)) if precision == "bf16" else (T.numpy(), V.numpy())
    wv, wi = O.topk_exact(Tq, Vq, k, metric=METRICS[metric], col_offset=1000)
    np.testing.assert_array_equal(_np(idx), wi)
    if metric == "l2":  # faiss-style distances carry ||q||^2
        wv = wv + (np.asarray(Tq, dtype=np.float64) ** 2).sum(1, keepdims=True)
    np.testing.assert_allclose(_np(vals), wv, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("precision", ["exact", "bf16"])
@pytest.mark.parametrize("k", [11, 16])
def test_topk_sample_pass_large_gallery(cuda_dev, monkeypatch, precision, k):
    """Galleries of >= 128 tiles take a sample pass first (per-row threshold from the first 1/16 of
    the gallery); results must equal the oracle and the single-pass path."""
    from vtc_b200 import ops

    T, V = make_retrieval_pair(257, 40_000, 64, sigma=2.0, seed=77)
    V = V.clone()
    V[39_000] = V[5]      # exact tie between a sample column and a late column
    V[100] = V[20_000]    # and the other way round
    q, g = T.to(cuda_dev), V.to(cuda_dev)
    vals, idx = ops.sim_topk(q, g, k, precision=precision)
    Tq, Vq = (O.bf16_round(T), O.bf16_round(V)) if precision == "bf16" else (T.numpy(), V.numpy())
    wv, wi = O.topk_exact(Tq, Vq, k)
    np.testing.assert_array_equal(_np(idx), wi)
    monkeypatch.setenv("VTC_TOPK_NO_SAMPLE", "1")
    vals2, idx2 = ops.sim_topk(q, g, k, precision=precision)
    np.testing.assert_array_equal(_np(idx2), wi)
    np.testing.assert_allclose(_np(vals2), _np(vals), rtol=0, atol=0)


def test_topk_small_gallery_and_merge(cuda_dev):
    from vtc_b200 import ops

    T, V = make_retrieval_pair(5, 7, 64, sigma=1.0, seed=1)
    vals, idx = ops.sim_topk(T.to(cuda_dev), V.to(cuda_dev), 11, precision="exact")
    wv, wi = O.topk_exact(T, V, 11)
    np.testing.assert_array_equal(_np(idx), wi)  # -1 fill beyond the 7 gallery rows
    assert np.isinf(_np(vals)[:, 7:]).all()
    # gallery-sharded search + merge == whole search
    T, V = make_retrieval_pair(200, 3000, 128, sigma=3.0, seed=9)
    q, g = T.to(cuda_dev), V.to(cuda_dev)
    whole_v, whole_i = ops.sim_topk(q, g, 10, precision="exact")
    parts = [ops.sim_topk(q, g[s:e].contiguous(), 10, precision="exact", col_offset=s)
             for s, e in ((0, 1000), (1000, 1001), (1001, 3000))]
    mv, mi = ops.topk_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    np.testing.assert_array_equal(_np(mi), _np(whole_i))
    np.testing.assert_allclose(_np(mv), _np(whole_v), rtol=1e-6)


# ------------------------------------------------------------------------------------- H2 + H3
@pytest.mark.parametrize("name", ["c2_s100", "c2_s14", "small", "ragged"])
@pytest.mark.parametrize("precision", ["exact", "bf16"])
@pytest.mark.parametrize("path", ["fused_small", "tcgen05"])
def test_clip_loss_fused_matches_reference_golden(cuda_dev, golden, monkeypatch, name, precision, path):
    """model/loss.py::clip_loss through the LazySim route against the value the reference's own
    clip_loss produced (fixtures), and the saved LSEs against fp64.  Both kernels are covered: the
    single-launch fused kernel used for n <= 2048 and the tcgen05 online-LSE kernel."""
    from vtc_b200 import ops
    from vtc_b200.model import LazySim, clip_loss

    if path == "tcgen05":
        monkeypatch.setenv("VTC_INFONCE_FORCE_TC", "1")

    g = golden("clip_loss.npz")
    b, D, s = g[name + "_cfg"]
    vis, txt = make_batch_pair(int(b), int(D), seed=1023)
    a, t = vis.to(cuda_dev), txt.to(cuda_dev)
    scale = torch.tensor(float(s), device=cuda_dev)
    loss = clip_loss((a, t, LazySim(a, t, scale, precision)), {})
    tol = 1e-4 if precision == "exact" else 2e-2
    assert loss.dim() == 0 and loss.device.type == "cuda"
    np.testing.assert_allclose(loss.item(), float(g[name + "_loss"]), rtol=tol)
    l2, row, col, diag = ops.infonce_fwd(a, t, scale, precision)
    p64 = O.clip_loss_parts64(vis, txt, float(s))
    atol = (1e-4 if precision == "exact" else 2e-2) * float(s)
    np.testing.assert_allclose(_np(row), p64["row_lse"], atol=atol, rtol=0)
    np.testing.assert_allclose(_np(col), p64["col_lse"], atol=atol, rtol=0)
    np.testing.assert_allclose(_np(diag), p64["diag"], atol=atol, rtol=0)


def test_clip_loss_large_batch_tcgen05(cuda_dev):
    """n = 3000 > 2048 takes the tcgen05 online-LSE path by itself (several gallery tiles)."""
    from vtc_b200 import ops

    vis, txt = make_batch_pair(3000, 256, seed=5)
    for precision, tol in (("exact", 1e-4), ("bf16", 2e-2)):
        loss, row, col, diag = ops.infonce_fwd(vis.to(cuda_dev), txt.to(cuda_dev), 50.0, precision)
        p64 = O.clip_loss_parts64(vis, txt, 50.0)
        np.testing.assert_allclose(loss.item(), p64["loss"], rtol=tol)
        np.testing.assert_allclose(_np(row), p64["row_lse"], atol=tol * 50, rtol=0)
        np.testing.assert_allclose(_np(col), p64["col_lse"], atol=tol * 50, rtol=0)


def test_clip_loss_one_pass_column_sums_and_overflow_fallback(cuda_dev):
    """The tcgen05 InfoNCE forward is ONE pass over the logits (row log-sum-exp online, column sums
    against each column's positive logit).  A negative that beats its column's positive by more than
    88 nats overflows that reference: the merge kernel notices and gates the transposed pass on --
    results must still match fp64.  Also n = 16384 (64 row tiles x 64 column tiles) against an fp64
    computation on the device."""
    from vtc_b200 import _ffi, ops

    # (a) launch count: one tensor-core pass does rows AND columns; the gated second pass exits at once
    vis, txt = make_batch_pair(3000, 256, seed=5)
    a, t = vis.to(cuda_dev), txt.to(cuda_dev)
    loss, row, col, diag = ops.infonce_fwd(a, t, 50.0, "bf16")
    p64 = O.clip_loss_parts64(O.bf16_round(vis), O.bf16_round(txt), 50.0)
    np.testing.assert_allclose(_np(col), p64["col_lse"], atol=2e-3, rtol=0)
    np.testing.assert_allclose(_np(row), p64["row_lse"], atol=2e-3, rtol=0)
    # (b) overflow of the column reference -> device-side fallback
    n, D, s = 2200, 128, 300.0
    g = torch.Generator().manual_seed(3)
    u = torch.nn.functional.normalize(torch.randn(1, D, generator=g), dim=1)
    txt2 = torch.nn.functional.normalize(u + 0.05 * torch.randn(n, D, generator=g), dim=1)
    vis2 = torch.nn.functional.normalize(u + 0.05 * torch.randn(n, D, generator=g), dim=1)
    w = torch.nn.functional.normalize(torch.randn(1, D, generator=g), dim=1)
    vis2[0] = torch.nn.functional.normalize(w - (w @ txt2[0]) * txt2[0:1], dim=1)  # positive logit ~ 0
    for precision, tol in (("exact", 1e-4), ("bf16", 2e-2)):
        loss, row, col, diag = ops.infonce_fwd(vis2.to(cuda_dev), txt2.to(cuda_dev), s, precision)
        p64 = O.clip_loss_parts64(vis2, txt2, s)
        assert np.isfinite(_np(col)).all() and np.isfinite(loss.item())
        np.testing.assert_allclose(_np(col), p64["col_lse"], atol=tol * s, rtol=0)
        np.testing.assert_allclose(_np(row), p64["row_lse"], atol=tol * s, rtol=0)
        np.testing.assert_allclose(loss.item(), p64["loss"], rtol=tol, atol=tol)
    # (c) n = 16384: forward (+ the saved statistics) against fp64 on the device
    n, D, s = 16384, 512, 100.0
    gen = torch.Generator(device=cuda_dev).manual_seed(11)
    V = torch.nn.functional.normalize(torch.randn(n, D, generator=gen, device=cuda_dev), dim=1)
    T = torch.nn.functional.normalize(V + 8.0 * torch.randn(n, D, generator=gen, device=cuda_dev) / D ** 0.5, dim=1)
    loss, row, col, diag = ops.infonce_fwd(V, T, s, "exact")
    sim = s * (V.double() @ T.double().t())
    want_row = torch.logsumexp(sim, dim=1)
    want_col = torch.logsumexp(sim, dim=0)
    want = 0.5 * ((want_row - sim.diag()).mean() + (want_col - sim.diag()).mean())
    np.testing.assert_allclose(_np(row), _np(want_row), atol=1e-4 * s, rtol=0)
    np.testing.assert_allclose(_np(col), _np(want_col), atol=1e-4 * s, rtol=0)
    np.testing.assert_allclose(loss.item(), want.item(), rtol=1e-4)


def test_clip_loss_backward_and_dense_sim(cuda_dev, golden):
    from vtc_b200.model import LazySim, clip_loss

    g = golden("clip_loss.npz")
    vis, txt = torch.from_numpy(g["small_vis"]), torch.from_numpy(g["small_txt"])
    s = float(g["small_cfg"][2])
    # oracle gradients: the reference formula under torch autograd on the CPU
    a0 = vis.clone().requires_grad_(True)
    t0 = txt.clone().requires_grad_(True)
    ls0 = torch.tensor(math.log(s), requires_grad=True)
    O.clip_loss(O.sim_matrix(a0, t0, ls0.exp())).backward()
    a = vis.to(cuda_dev).requires_grad_(True)
    t = txt.to(cuda_dev).requires_grad_(True)
    ls = torch.tensor(math.log(s), device=cuda_dev, requires_grad=True)
    loss = clip_loss((a, t, LazySim(a, t, ls.exp(), "exact")), {})
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g["small_loss"]), rtol=1e-4)
    np.testing.assert_allclose(_np(a.grad), a0.grad.numpy(), rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(_np(t.grad), t0.grad.numpy(), rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(ls.grad.item(), ls0.grad.item(), rtol=2e-3, atol=1e-5)
    # a materialised sim tensor works through the same call site (and yields d loss / d sim)
    sim = torch.from_numpy(g["small_sim"]).to(cuda_dev).requires_grad_(True)
    loss2 = clip_loss((None, None, sim), {})
    loss2.backward()
    np.testing.assert_allclose(loss2.item(), float(g["small_loss"]), rtol=1e-4)
    np.testing.assert_allclose(_np(sim.grad), g["small_dsim"], rtol=2e-3, atol=1e-6)
    # LazySim materialises to the reference's sim
    lz = LazySim(vis.to(cuda_dev), txt.to(cuda_dev), s, "exact")
    assert lz.shape == (32, 32)
    np.testing.assert_allclose(_np(lz.materialize()), g["small_sim"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(_np(torch.diagonal(lz)), np.diag(g["small_sim"]), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("M_total,N", [(700, 333), (1023, 1000), (5000, 4096), (100_000, 10_001),
                                       (3_000_000, 5000), (3_000_000, 4999), (1 << 28, 2048)])
def test_rank_finalize_radix_levels(cuda_dev, M_total, N):
    """R@K hit counts and MedR = median(rank0) + 1 with numpy semantics from the
    finalisation chain alone (csrc/rank_stage.cu: commit + radix select).  The select takes one, two or
    three digits depending on the gallery size (< 2^10, < 2^21, larger); the evaluation tests only reach
    the first two.  Even and odd N (the two middle order statistics), heavy ties, NaN ground truths."""
    from vtc_b200 import ops

    rng = np.random.default_rng(M_total % 1000 + N)
    ranks = np.minimum(rng.geometric(1.0 / max(2, M_total // 50), size=N) - 1, M_total - 1).astype(np.int32)
    ranks[: N // 3] = rng.integers(0, 12, size=N // 3)          # many small, tied ranks
    ranks[-5:] = M_total - 1                                      # the far end of the range
    gts = rng.standard_normal(N)
    gts[::97] = np.nan                                            # no ground truth -> rank M_total
    want = ranks.copy()
    want[::97] = M_total
    k_vals = [1, 5, 10, 100]
    r = torch.from_numpy(ranks).to(cuda_dev)
    hits, medr = ops.rank_finalize(r, torch.from_numpy(gts).to(cuda_dev), M_total, k_vals)
    np.testing.assert_array_equal(_np(r), want)
    np.testing.assert_array_equal(_np(hits), [(want < k).sum() for k in k_vals])
    assert medr.item() == float(np.median(want.astype(np.float64)) + 1.0)


@pytest.mark.parametrize("n,s", [(300, 100.0), (1000, 14.29), (33, 1.0)])
def test_clip_loss_on_a_materialised_sim(cuda_dev, n, s):
    """model/loss.py:18-22 handed a real tensor (the reference's own forward returns one,
    model/model.py:369): reductions over the matrix itself (csrc/infonce_dense.cu), loss and
    d loss / d sim against the reference formula in fp64."""
    from vtc_b200.model import clip_loss

    vis, txt = make_batch_pair(n, 128, seed=9)
    sim0 = (s * vis.double() @ txt.double().t()).requires_grad_(True)
    want = O.clip_loss(sim0)
    want.backward()
    sim = sim0.detach().float().to(cuda_dev).requires_grad_(True)
    loss = clip_loss((None, None, sim), {})
    loss.backward()
    np.testing.assert_allclose(loss.item(), want.item(), rtol=1e-4, atol=1e-6)
    # (entries are <= 1/n; the diagonal ones cancel against -1/n, so bound them absolutely)
    np.testing.assert_allclose(_np(sim.grad), sim0.grad.numpy(), rtol=2e-3, atol=3e-5 / n)


@pytest.mark.parametrize("n,D,precision", [(2100, 64, "exact"), (2100, 64, "bf16"),
                                           (2304, 200, "exact"), (4096, 512, "bf16")])
def test_clip_loss_backward_tensor_cores(cuda_dev, n, D, precision):
    """loss.backward() of model/loss.py:18-22 (trainer/trainer.py:79) for n > 2048: the tcgen05 path of
    vtc_infonce_bwd (logit tiles recomputed, gradient weights as bf16 operand strips) against torch
    autograd through the reference formula in fp64 on the device (oracle.clip_loss is that formula;
    n x n in fp64 is too slow for the CPU at these sizes)."""
    from vtc_b200.model import LazySim, clip_loss

    s = 30.0
    vis, txt = make_batch_pair(n, D, seed=21)
    if precision == "bf16":  # the bf16 mode differentiates the loss of the bf16-rounded features
        vis, txt = torch.from_numpy(O.bf16_round(vis.numpy())), torch.from_numpy(O.bf16_round(txt.numpy()))
    a0 = vis.to(cuda_dev).double().requires_grad_(True)
    t0 = txt.to(cuda_dev).double().requires_grad_(True)
    ls0 = torch.tensor(math.log(s), device=cuda_dev, dtype=torch.float64, requires_grad=True)
    sim = ls0.exp() * a0 @ t0.t()
    lab = torch.arange(n, device=cuda_dev)
    want = 0.5 * (torch.nn.functional.cross_entropy(sim, lab) + torch.nn.functional.cross_entropy(sim.t(), lab))
    want.backward()
    a = vis.to(cuda_dev).requires_grad_(True)
    t = txt.to(cuda_dev).requires_grad_(True)
    ls = torch.tensor(math.log(s), device=cuda_dev, requires_grad=True)
    loss = clip_loss((a, t, LazySim(a, t, ls.exp(), precision)), {})
    loss.backward()
    np.testing.assert_allclose(loss.item(), want.item(), rtol=1e-4 if precision == "exact" else 2e-2)
    # gradient rows are sums of n weighted unit vectors: compare against the largest entry
    rtol = 1e-3 if precision == "exact" else 2e-2
    for got, ref in ((a.grad, a0.grad), (t.grad, t0.grad)):
        ref = ref.float()
        err = (got - ref).abs().max().item()
        assert err <= rtol * ref.abs().max().item(), (err, ref.abs().max().item())
    np.testing.assert_allclose(ls.grad.item(), ls0.grad.item(), rtol=5 * rtol, atol=1e-6)


def test_clip_loss_forward_backward_n16384(cuda_dev):
    """n = 16384: beyond the n <= 8192 cap of round 1 (an n x n fp32 matrix in the workspace); the
    strips keep the scratch at O(R n).  Checked on a slice of rows against fp64."""
    from vtc_b200.model import LazySim, clip_loss

    n, D, s = 16384, 512, 100.0
    gen = torch.Generator(device=cuda_dev).manual_seed(11)
    V = torch.nn.functional.normalize(torch.randn(n, D, generator=gen, device=cuda_dev), dim=1)
    T = torch.nn.functional.normalize(V + 8.0 * torch.randn(n, D, generator=gen, device=cuda_dev) / D ** 0.5, dim=1)
    a, t = V.clone().requires_grad_(True), T.clone().requires_grad_(True)
    loss = clip_loss((a, t, LazySim(a, t, torch.tensor(s, device=cuda_dev), "exact")), {})
    loss.backward()
    sim = s * (V.double() @ T.double().t())
    row, col = torch.logsumexp(sim, dim=1), torch.logsumexp(sim, dim=0)
    W = (torch.exp(sim - row[:, None]) + torch.exp(sim - col[None, :])) / (2 * n)
    W -= torch.eye(n, device=cuda_dev, dtype=torch.float64) / n
    wantA = (s * W[:256] @ T.double()).float()
    wantB = (s * W[:, :256].t() @ V.double()).float()
    for got, ref in ((a.grad[:256], wantA), (t.grad[:256], wantB)):
        err = (got - ref).abs().max().item()
        assert err <= 1e-3 * ref.abs().max().item(), (err, ref.abs().max().item())


# ------------------------------------------------------------------------------------------ H4
def _make_cam(D, layers, heads, params, init_from_avg, flw, dev, precision="exact"):
    from vtc_b200.model import PretrainedCLIP_finaltf

    m = PretrainedCLIP_finaltf(D, n_layers=layers, n_heads=heads, init_from_avg=init_from_avg,
                               precision=precision)
    missing, unexpected = m.final_transformer.load_state_dict(params, strict=True)
    with torch.no_grad():
        m.final_linear.weight.copy_(flw)
    return m.to(dev).eval()


@pytest.mark.parametrize("name", ["c2_init", "c2_rand", "c2_linear", "small"])
@pytest.mark.parametrize("precision", ["exact", "bf16"])
def test_cam_adapt_feature_golden(cuda_dev, golden, name, precision):
    g = golden("cam.npz")
    b, nc, D, layers, heads, rerand, avg = [int(x) for x in g[name + "_cfg"]]
    params = O.make_cam_params(D, layers, heads, seed=1023, rerandomise=bool(rerand))
    gen = torch.Generator().manual_seed(99)
    flw = torch.randn(D, D, generator=gen) / D ** 0.5
    cam = _make_cam(D, layers, heads, params, bool(avg), flw, cuda_dev, precision)
    main, aux = make_cam_inputs(b, nc, D, seed=1023)
    with torch.no_grad():
        out = cam._adapt_feature(main.to(cuda_dev), aux.to(cuda_dev))
    want = g[name + "_adapted"]
    tol = dict(rtol=1e-4, atol=1e-5) if precision == "exact" else dict(rtol=2e-2, atol=2e-3)
    np.testing.assert_allclose(_np(out)[:want.shape[0]], want, **tol)
    np.testing.assert_allclose(_np(out.norm(dim=-1)), 1.0, rtol=1e-5)


@pytest.mark.parametrize("act", ["normalize", "squash", "squash10", "squash1p5", "tanh", "sub_mean", "bn"])
def test_cam_residual_activations_golden(cuda_dev, golden, act):
    """the residual-activation table of model/model.py:30-77 inside the fused
    read-out kernel, against the reference's own outputs."""
    from vtc_b200.model import PretrainedCLIP_finaltf

    g = golden("cam.npz")
    b, nc, D, layers, heads = 8, 3, 64, 2, 2
    params = O.make_cam_params(D, layers, heads, seed=1023, rerandomise=True)
    m = PretrainedCLIP_finaltf(D, n_layers=layers, n_heads=heads, residual_activation=act)
    m.final_transformer.load_state_dict(params, strict=True)
    if act in ("sub_mean", "bn"):
        m.mean_center_bn.running_mean.copy_(torch.from_numpy(g["act_running_mean"]))
        m.mean_center_bn.running_var.copy_(torch.from_numpy(g["act_running_var"]))
    m = m.to(cuda_dev).eval()
    main, aux = make_cam_inputs(b, nc, D, seed=1023)
    with torch.no_grad():
        out = m._adapt_feature(main.to(cuda_dev), aux.to(cuda_dev))
    np.testing.assert_allclose(_np(out), g["act_" + act], rtol=2e-4, atol=1e-5)


@pytest.mark.parametrize("avg", [True, False])
def test_cam_backward_matches_autograd_oracle(cuda_dev, avg):
    """Training path: gradients of `_adapt_feature` w.r.t. inputs and every CAM parameter against
    torch autograd through the oracle restatement (CPU)."""
    from vtc_b200.model import PretrainedCLIP_finaltf

    b, nc, D, layers, heads = 24, 5, 128, 2, 4
    params = O.make_cam_params(D, layers, heads, seed=11, rerandomise=True)
    flw = torch.randn(D, D, generator=torch.Generator().manual_seed(3)) / D ** 0.5
    main, aux = make_cam_inputs(b, nc, D, seed=6)
    w = torch.randn(b, D, generator=torch.Generator().manual_seed(9))  # loss = sum(out * w)
    skip = torch.zeros(b, dtype=torch.bool)
    skip[::5] = True
    # oracle gradients
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    mo, ao, fo = main.clone().requires_grad_(True), aux.clone().requires_grad_(True), flw.clone().requires_grad_(True)
    out_o = O.adapt_feature(mo, ao, po, layers, heads, init_from_avg=avg, final_linear_weight=fo,
                            skip_mask=skip)
    (out_o * w).sum().backward()
    # CUDA path
    m = PretrainedCLIP_finaltf(D, n_layers=layers, n_heads=heads, init_from_avg=avg).to(cuda_dev)
    m.final_transformer.load_state_dict(params, strict=True)
    with torch.no_grad():
        m.final_linear.weight.copy_(flw)
    m.train()
    m.random_skip_adapter = False  # the mask is injected below instead of drawn
    mg, ag = main.to(cuda_dev).requires_grad_(True), aux.to(cuda_dev).requires_grad_(True)
    from vtc_b200.model.model import _CamAdaptFunction, _layer_params

    plist = [p for blk in m.final_transformer.resblocks for p in _layer_params(blk)]
    cfg = (layers, heads, avg, "exact", None)
    out = _CamAdaptFunction.apply(mg, ag, skip.to(cuda_dev), cfg, None if avg else m.final_linear.weight,
                                  *plist)
    np.testing.assert_allclose(_np(out), out_o.detach().numpy(), rtol=2e-4, atol=2e-5)
    (out * w.to(cuda_dev)).sum().backward()

    def close(got, want, name):
        want = want.numpy()
        scale = np.abs(want).max() + 1e-12
        err = np.abs(_np(got) - want).max() / scale
        assert err < 3e-3, f"{name}: max err / max |grad| = {err:.3e}"

    close(mg.grad, mo.grad, "dmain")
    close(ag.grad, ao.grad, "daux")
    for n_, p in m.final_transformer.named_parameters():
        close(p.grad, po[n_].grad, n_)
    if not avg:
        close(m.final_linear.weight.grad, fo.grad, "final_linear.weight")
    # and through the module call site (train mode, autograd on) the same function is used
    m.zero_grad()
    out2 = m._adapt_feature(mg, ag)
    assert out2.requires_grad


@pytest.mark.parametrize("act,avg", [("normalize", True), ("squash", False), ("squash10", True),
                                     ("tanh", False), ("tanh", True), ("sub_mean", True), ("bn", False)])
def test_cam_backward_residual_activations(cuda_dev, act, avg):
    """SURVEY.md §8f row 4, training side: the Jacobians of the residual-activation table
    (model/model.py:30-77; sub_mean / bn on running statistics = frozen-finaltf form) inside
    vtc_cam_readout_bwd, against torch autograd through the oracle."""
    from vtc_b200.model import PretrainedCLIP_finaltf

    b, nc, D, layers, heads = 16, 3, 64, 1, 2
    params = O.make_cam_params(D, layers, heads, seed=13, rerandomise=True)
    g = torch.Generator().manual_seed(4)
    flw = torch.randn(D, D, generator=g) / D ** 0.5
    run_mean, run_var = 0.05 * torch.randn(D, generator=g), 0.5 + torch.rand(D, generator=g)
    main, aux = make_cam_inputs(b, nc, D, seed=2)
    w = torch.randn(b, D, generator=g)
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    mo, ao, fo = main.clone().requires_grad_(True), aux.clone().requires_grad_(True), flw.clone().requires_grad_(True)
    out_o = O.adapt_feature(mo, ao, po, layers, heads, init_from_avg=avg, final_linear_weight=fo,
                            residual_activation=act, bn_state=(run_mean, run_var, 1e-5))
    (out_o * w).sum().backward()

    m = PretrainedCLIP_finaltf(D, n_layers=layers, n_heads=heads, init_from_avg=avg,
                               residual_activation=act)
    if act in ("sub_mean", "bn"):
        m.branch_to_freeze = "finaltf"  # what _freeze("finaltf") records (model/model.py:268-269)
    m.final_transformer.load_state_dict(params, strict=True)
    with torch.no_grad():
        m.final_linear.weight.copy_(flw)
        if act in ("sub_mean", "bn"):
            m.mean_center_bn.running_mean.copy_(run_mean)
            m.mean_center_bn.running_var.copy_(run_var)
    m = m.to(cuda_dev).train()
    m.random_skip_adapter = False
    mg, ag = main.to(cuda_dev).requires_grad_(True), aux.to(cuda_dev).requires_grad_(True)
    out = m._adapt_feature(mg, ag)
    np.testing.assert_allclose(_np(out), out_o.detach().numpy(), rtol=2e-4, atol=2e-5)
    (out * w.to(cuda_dev)).sum().backward()

    def close(got, want, name):
        want = want.numpy()
        err = np.abs(_np(got) - want).max() / (np.abs(want).max() + 1e-12)
        assert err < 3e-3, f"{name}: max err / max |grad| = {err:.3e}"

    close(mg.grad, mo.grad, "dmain")
    close(ag.grad, ao.grad, "daux")
    for n_, p in m.final_transformer.named_parameters():
        close(p.grad, po[n_].grad, n_)
    if not avg:
        close(m.final_linear.weight.grad, fo.grad, "final_linear.weight")


def test_training_step_gradients_through_the_model(cuda_dev):
    """The whole training call chain of trainer/trainer.py:77-79 on precomputed features:
    model(vis, title, comments) -> clip_loss -> backward, against torch autograd through the oracle."""
    from vtc_b200.model import PretrainedCLIP, PretrainedCLIP_finaltf, clip_loss

    b, nc, D, layers, heads = 32, 3, 128, 2, 4
    params = O.make_cam_params(D, layers, heads, seed=21, rerandomise=True)
    g = torch.Generator().manual_seed(5)
    vis, title, comm = torch.randn(b, D, generator=g), torch.randn(b, D, generator=g), torch.randn(b, nc, D, generator=g)
    ls0 = math.log(20.0)
    # oracle
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    to, co, lo = title.clone().requires_grad_(True), comm.clone().requires_grad_(True), torch.tensor(ls0, requires_grad=True)
    ft = O.normalize(O.adapt_feature(to, co.permute(1, 0, 2), po, layers, heads))
    loss_o = O.clip_loss(O.sim_matrix(O.normalize(vis), ft, lo.exp()))
    loss_o.backward()
    # CUDA path (eval-mode routing, autograd on)
    m = PretrainedCLIP_finaltf(D, n_layers=layers, n_heads=heads, logit_scale_init=ls0).to(cuda_dev)
    m.final_transformer.load_state_dict(params, strict=True)
    m.eval()
    tg, cg = title.to(cuda_dev).requires_grad_(True), comm.to(cuda_dev).requires_grad_(True)
    out = m(vis.to(cuda_dev), tg, cg)
    loss = clip_loss(out, {})
    np.testing.assert_allclose(loss.item(), loss_o.item(), rtol=1e-4)
    loss.backward()

    def close(got, want, name, tol=5e-3):
        want = want.numpy()
        err = np.abs(_np(got) - want).max() / (np.abs(want).max() + 1e-12)
        assert err < tol, f"{name}: {err:.3e}"

    close(tg.grad, to.grad, "dtitle")
    close(cg.grad, co.grad, "dcomments")
    close(m.logit_scale.grad, lo.grad, "dlogit_scale")
    for n_, p in m.final_transformer.named_parameters():
        close(p.grad, po[n_].grad, n_)
    # averaging-fusion baseline (model/model.py:356-366) is differentiable too
    m2 = PretrainedCLIP(D, comment_fusion="averaging", logit_scale_init=ls0).to(cuda_dev)
    t2, c2 = title.to(cuda_dev).requires_grad_(True), comm.to(cuda_dev).requires_grad_(True)
    loss2 = clip_loss(m2(vis.to(cuda_dev), t2, c2), {})
    loss2.backward()
    t3, c3 = title.clone().requires_grad_(True), comm.clone().requires_grad_(True)
    l3 = O.clip_loss(O.sim_matrix(O.normalize(vis), O.averaging_fusion(t3, c3), torch.tensor(20.0)))
    l3.backward()
    np.testing.assert_allclose(loss2.item(), l3.item(), rtol=1e-4)
    close(t2.grad, t3.grad, "avg dtitle")
    close(c2.grad, c3.grad, "avg dcomments")


def test_cam_transformer_module_forward(cuda_dev):
    """The op-by-op CAMTransformer.forward (LayerNorm / tcgen05 linears / attention core as separate
    C-ABI calls) against the oracle's clip.model.Transformer restatement."""
    from vtc_b200.model import CAMTransformer

    L, b, D, layers, heads = 6, 40, 256, 2, 4
    params = O.make_cam_params(D, layers, heads, seed=5, rerandomise=True)
    tfm = CAMTransformer(D, layers, heads, "exact")
    tfm.load_state_dict(params, strict=True)
    tfm = tfm.to(cuda_dev).eval()
    x = torch.randn(L, b, D, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        got = tfm(x.to(cuda_dev))
    want = O.transformer(x, params, layers, heads)
    np.testing.assert_allclose(_np(got), want.numpy(), rtol=2e-4, atol=2e-5)
    with pytest.raises(NotImplementedError):
        tfm(x.to(cuda_dev))  # autograd on: forward-only this round


def test_cam_properties(cuda_dev):
    """Restatements of the two reference tests (tests/test_pretrained_clip.py) on synthetic
    features: skip == identity; adapting text leaves vision untouched; closed form at init."""
    from vtc_b200.model import PretrainedCLIP_finaltf

    torch.manual_seed(0)
    b, nc, D = 48, 5, 512
    vis, title = torch.randn(b, D), torch.randn(b, D)
    comm = torch.randn(b, nc, D)
    m = PretrainedCLIP_finaltf(D).to(cuda_dev).eval()
    with torch.no_grad():
        fv, ft, sim = m(vis.to(cuda_dev), title.to(cuda_dev), comm.to(cuda_dev))
        np.testing.assert_allclose(_np(fv), O.normalize(vis).numpy(), rtol=1e-5, atol=1e-6)
        want = O.cam_closed_form_at_init(title, comm.permute(1, 0, 2))
        np.testing.assert_allclose(_np(ft), want.numpy(), rtol=1e-4, atol=1e-5)
        m.branch_to_adapt_val = "skip"
        fv2, ft2, _ = m(vis.to(cuda_dev), title.to(cuda_dev), comm.to(cuda_dev))
        np.testing.assert_allclose(_np(ft2), O.normalize(title).numpy(), rtol=1e-5, atol=1e-6)
        m.branch_to_adapt_val = "image"
        fv3, ft3, _ = m(vis.to(cuda_dev), title.to(cuda_dev), comm.to(cuda_dev))
        np.testing.assert_allclose(_np(ft3), _np(ft2), rtol=0, atol=0)
        assert not np.allclose(_np(fv3), _np(fv))
    # sim is lazy, shaped [b, b], and equals (s * fv) @ ft.t() when touched
    assert tuple(sim.shape) == (b, b)
    dense = sim.materialize()
    s = math.exp(math.log(1 / 0.07))
    np.testing.assert_allclose(_np(dense), s * (_np(fv).astype(np.float64) @ _np(ft).astype(np.float64).T),
                               rtol=1e-4, atol=1e-3)


def test_averaging_fusion(cuda_dev):
    from vtc_b200.model import PretrainedCLIP

    torch.manual_seed(1)
    b, nc, D = 33, 5, 512
    vis, title, comm = torch.randn(b, D), torch.randn(b, D), torch.randn(b, nc, D)
    m = PretrainedCLIP(D, comment_fusion="averaging").to(cuda_dev).eval()
    with torch.no_grad():
        fv, ft, _ = m(vis.to(cuda_dev), title.to(cuda_dev), comm.to(cuda_dev))
    np.testing.assert_allclose(_np(ft), O.averaging_fusion(title, comm).numpy(), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------- multi-GPU (§8e)
def test_sharded_eval_single_process_equals_whole(cuda_dev):
    from vtc_b200 import ops
    from vtc_b200.parallel import sharded_rank_eval, sharded_topk

    T, V = make_retrieval_pair(600, 600, 128, sigma=3.0, seed=4)
    q, g = T.to(cuda_dev), V.to(cuda_dev)
    res = sharded_rank_eval(q, g, 600, 600, precision="exact")
    want = O.rank0_exact(T, V)
    np.testing.assert_array_equal(_np(res["rank0_local"]), want)
    np.testing.assert_array_equal(_np(res["hits"]), [np.sum(want < k) for k in (1, 5, 10)])
    assert _np(res["medr"])[0] == O.medr(want)
    v, i = sharded_topk(q, g, 600, 5, precision="exact")
    np.testing.assert_array_equal(_np(i), O.topk_exact(T, V, 5)[1])


# ----------------------------------------------------------- variants promoted in round 2
def test_more_than_eight_k_values(cuda_dev):
    """RecallAtK accepts any number of k values like the reference (model/metric.py:104-108); the C
    entry point counts 8 per call."""
    from vtc_b200.model.metric import RecallAtK

    T, V = make_retrieval_pair(400, 400, 64, sigma=3.0, seed=4)
    ks = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 20, 50]
    got = RecallAtK("v", "t", ks).compute(V.to(cuda_dev), T.to(cuda_dev))
    want = O.recall_at_k(V.numpy(), T.numpy(), ks)
    assert [k for k, _ in got] == ks
    assert [r for _, r in got] == [r for _, r in want]


@pytest.mark.parametrize("precision", ["exact", "bf16"])
def test_cached_ranking_same_ranks(cuda_dev, golden, precision):
    """vtc_rank_prepare / vtc_sim_rank_prepared: norms, norm bounds and ground-truth scores computed
    once (by vtc_rank_prepare or by the first ranking call that sees the rows) and handed to every
    later call as slices of per-gallery arrays -- the host-staging pipeline and the sharded step run
    on this.  Same ranks as the oracle, whichever call produced the cached values."""
    from vtc_b200 import ops

    def canon(x):  # cached values are tied to the canonical rows of the mode
        return x.to(cuda_dev).bfloat16() if precision == "bf16" else x.to(cuda_dev)

    cases = [make_retrieval_pair(1000, 1000, 512, sigma=4.0, seed=3) + ("l2",),
             make_retrieval_pair(333, 1201, 96, sigma=2.0, seed=4) + ("dot",),
             make_retrieval_pair(129, 257, 768, sigma=7.0, seed=5) + ("l2",)]
    g = golden("retrieval_small.npz")           # ties, zero row, non-unit row, NaN query, inf rows
    cases.append((torch.from_numpy(g["queries"]), torch.from_numpy(g["gallery"]), "l2"))
    for T, V, metric in cases:
        q, gal = canon(T), canon(V)
        N, M = q.shape[0], gal.shape[0]
        want = _oracle_ranks(T, V, metric, precision)
        gts = ops.gt_scores(q, gal, metric=metric, precision=precision)
        sq64, _ = ops.rank_prepare(gal, precision, want_qq=False)
        _, qq = ops.rank_prepare(q, precision, want_sq64=False)
        np.testing.assert_array_equal(_np(sq64), O.sqnorm64(_np(gal.float())))
        assert (_np(qq) >= O.sqnorm64(_np(q.float())))[~np.isnan(_np(qq))].all()
        # whole gallery in one call, everything handed in
        rank0 = torch.full((N,), -5, dtype=torch.int32, device=cuda_dev)
        ops.sim_rank(q, gal, metric=metric, precision=precision, gt_score=gts, rank0=rank0,
                     accumulate=False, sq64=sq64, qq=qq)
        ops.rank_finalize(rank0, gts, M, [1])
        np.testing.assert_array_equal(_np(rank0), want)
        # the ranking call itself produces the cached values: same bits as vtc_rank_prepare's
        sq_o = torch.empty(M, dtype=torch.float64, device=cuda_dev)
        qq_o = torch.empty(N, dtype=torch.float32, device=cuda_dev)
        gs_o = torch.empty(N, dtype=torch.float64, device=cuda_dev)
        r2, _ = ops.sim_rank(q, gal, metric=metric, precision=precision, gt_score_out=gs_o,
                             sq64_out=sq_o, qq_out=qq_o)
        ops.rank_finalize(r2, gs_o, M, [1])
        np.testing.assert_array_equal(_np(r2), want)
        np.testing.assert_array_equal(_np(sq_o), _np(sq64))
        np.testing.assert_array_equal(_np(gs_o), _np(gts))
        # ... and accumulated over ragged gallery chunks with slices of the same arrays
        acc = torch.zeros(N, dtype=torch.int32, device=cuda_dev)
        bounds = [0, M // 3, M // 3 + 1, M]
        for s, e in zip(bounds[:-1], bounds[1:]):
            ops.sim_rank(q, gal[s:e].contiguous(), col_offset=s, metric=metric, precision=precision,
                         gt_score=gts, rank0=acc, accumulate=True, sq64=sq64[s:e].contiguous(), qq=qq)
        ops.rank_finalize(acc, gts, M, [1])
        np.testing.assert_array_equal(_np(acc), want)


def test_host_staging_pipeline_matches_device_path(cuda_dev):
    """RecallAtK.compute from host arrays (interleaved chunk pairs, cached per-row quantities): the
    same ranks and hits as the single device-resident call, and as the oracle on a slice."""
    from vtc_b200.model.metric import RecallAtK

    N, D = 40_000, 256                           # 41 MB per side: above PIPELINE_MIN_BYTES
    T, V = make_retrieval_pair(N, N, D, sigma=5.0, seed=12)
    for precision in ("bf16", "exact"):
        m = RecallAtK("v", "t", [1, 5, 10], precision=precision)
        full = m.compute_full(V.numpy(), T.numpy())
        sl = slice(20_000, 20_200)
        want = O.rank0_exact(O.bf16_round(T[sl]) if precision == "bf16" else T[sl],
                             O.bf16_round(V) if precision == "bf16" else V, row_offset=20_000)
        np.testing.assert_array_equal(_np(full["rank0"][sl]), want)
        base = m.compute_full(V.to(cuda_dev), T.to(cuda_dev))
        np.testing.assert_array_equal(_np(full["rank0"]), _np(base["rank0"]))
        np.testing.assert_array_equal(_np(full["hits"]), _np(base["hits"]))
        assert _np(full["medr"])[0] == _np(base["medr"])[0]


@pytest.mark.parametrize("precision", ["brute", "exact", "bf16"])
def test_infinite_gallery_rows(cuda_dev, precision):
    """Gallery rows of -inf (the padding rows evaluation/retrieval_evaluation.py:238-252 creates for
    missing captions) and +inf, also as ground truth.  Canonical arithmetic gives such a column an
    infinite or NaN score (NaN comparisons false; -inf closer than everything).  The 3-term bf16
    split cannot carry inf (x - bf16(x) is NaN), so the exact mode must notice and recount in
    canonical arithmetic (round 1 silently lost the -inf scores of the DOT metric here)."""
    from vtc_b200 import ops

    T, V = make_retrieval_pair(300, 900, 128, sigma=3.0, seed=31)
    V[7] = float("-inf")          # ground truth of query 7
    V[400] = float("-inf")
    V[401, :5] = float("inf")
    V[20] = V[7]
    for metric in ("l2", "dot"):
        rank0, gts = ops.sim_rank(T.to(cuda_dev), V.to(cuda_dev), metric=metric, precision=precision)
        ops.rank_finalize(rank0, gts, 900, [1])
        want = _oracle_ranks(T, V, metric, "bf16" if precision == "bf16" else "exact")
        d0 = O.scores64(O.bf16_round(T) if precision == "bf16" else T,
                        O.bf16_round(V) if precision == "bf16" else V, METRICS[metric])[
            np.arange(300), np.arange(300)]
        want = np.where(np.isnan(d0), 900, want)     # vtc_rank_finalize: NaN score -> rank M
        np.testing.assert_array_equal(_np(rank0), want)
        # the one-call evaluation takes the same fallback
        full = ops.rank_eval(T.to(cuda_dev), V.to(cuda_dev), [1, 5, 10], metric=metric,
                             precision=precision)
        np.testing.assert_array_equal(_np(full["rank0"]), want)
        np.testing.assert_array_equal(_np(full["hits"]), [np.sum(want < k) for k in (1, 5, 10)])
        assert _np(full["medr"])[0] == O.medr(want)


@pytest.mark.parametrize("precision", ["brute", "exact", "bf16"])
@pytest.mark.parametrize("N,M,D", [(1000, 1000, 512), (333, 1201, 96), (5, 3000, 64), (700, 700, 100),
                                   (2500, 2500, 768)])
def test_rank_eval_one_call(cuda_dev, precision, N, M, D):
    """vtc_rank_eval (memset + prologue + tensor-core pass + cooperative epilogue) == vtc_sim_rank +
    vtc_rank_finalize == oracle: ranks, hit counts, median, ground-truth scores."""
    from vtc_b200 import ops

    T, V = make_retrieval_pair(min(N, M), M, D, sigma=3.0, seed=N + D, mixed=True)
    T = T[:N].contiguous()
    q, g = T.to(cuda_dev), V.to(cuda_dev)
    ks = [1, 5, 10]
    full = ops.rank_eval(q, g, ks, precision=precision)
    want = _oracle_ranks(T, V, "l2", precision)
    np.testing.assert_array_equal(_np(full["rank0"]), want)
    np.testing.assert_array_equal(_np(full["hits"]), [np.sum(want < k) for k in ks])
    assert _np(full["medr"])[0] == O.medr(want)
    r2, gts = ops.sim_rank(q, g, precision=precision)
    h2, m2 = ops.rank_finalize(r2, gts, M, ks)
    np.testing.assert_array_equal(_np(r2), want)
    np.testing.assert_array_equal(_np(h2), _np(full["hits"]))
    np.testing.assert_array_equal(_np(gts), _np(full["gt_score"]))
    assert _np(m2)[0] == _np(full["medr"])[0]


def test_rank_finalize_median_wide_ranks(cuda_dev):
    """The radix-select median over the full int32 range (three 11/11/10-bit levels), odd and even
    N, the two middle order statistics in different bins."""
    from vtc_b200 import ops

    rng = np.random.default_rng(5)
    for n, hi in ((1, 10), (2, 5_000_000), (1001, 3_000_000), (4096, 900), (50_000, 2**30)):
        r = rng.integers(0, hi, size=n).astype(np.int32)
        if n == 4096:
            r[:2048] = 3            # median straddles two far-apart values
            r[2048:] = 800
        t = torch.from_numpy(r).to(cuda_dev)
        hits, medr = ops.rank_finalize(t, None, int(hi), [1, 5, 10])
        np.testing.assert_array_equal(_np(hits), [np.sum(r < k) for k in (1, 5, 10)])
        assert _np(medr)[0] == O.medr(r)
