"""Parity tests of the OPT-IN kernel variants that have been written but not yet run on a B200
(DESIGN.md §8 "pending validation").  They are skipped unless VTC_TEST_EXPERIMENTAL=1, so that the
default `pytest -m gpu` run only exercises code that has been measured; the first GPU call of the next
round is `VTC_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q`
(scripts/gpu_experimental.sh), after which a variant is either made the default or removed.

Variants covered: VTC_RANK_FOLD / VTC_FOLD_COLS (fold epilogue), VTC_FAST_THR (thresholds from a
coalesced norm), vtc_rank_prepare + vtc_sim_rank_prepared / VTC_RANK_PREPARED (per-chunk prepared
quantities), VTC_PIPELINE_SCHEDULE (host-staging chunk sizes), plus > 8 k values.

VTC_RANK_FOLD=1 -- EPI_RANK_FOLD (csrc/fold.cu, RankFoldEpi in csrc/sim_tc_kernel.cuh): the
per-column bias and the per-row ground-truth score enter the accumulator through one extra K16 MMA
step, the epilogue counts sign bits.  Same contract as the default path: ranks bit-exact against the
fp64-sequential oracle (replacing the faiss search + hit loop of model/metric.py:137-161).

Reading a failure (nothing here has run on a GPU yet):
* cols16 fails, cols64 passes -> the assumption that a {64, rows} TMA box may exceed a 16-column
  tensor along the inner dimension (zero fill, as for a K tail) is wrong: drop VTC_FOLD_COLS=16.
* every fold test dies after ~4 s with a CUDA error -> an mbarrier wait hit its 4 s trap: the fold
  kernel differs from the measured pair kernel only in (a) nkb = num_kb + 1 k-blocks per tile in the
  producer and MMA loops, (b) 9 resident query blocks / a_full barriers, (c) a 5-stage ring when
  resident, (d) the last block issuing one K16 step -- check those four in sim_tc_kernel.cuh.
* ranks off by small counts on a few rows -> guard band: VTC_GUARD_REL_BF16 / _EXACT widen it; the
  fold step's own slack is the 4.8e-7 * S term of fold.cu::fold_q_kernel.
* ranks wildly off -> sign convention / operand layout: tests/test_fold_math.py is the numpy
  transcription the CUDA code must match (Qx = [m' pieces | 1 1 1], Gx = [1 1 1 | h pieces]).
"""
import os

import numpy as np
import pytest
import torch

from oracle import vtc_oracle as O
from vtc_b200.synthetic import make_retrieval_pair

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.skipif(os.environ.get("VTC_TEST_EXPERIMENTAL") != "1",
                       reason="opt-in variants: set VTC_TEST_EXPERIMENTAL=1"),
]

METRICS = {"l2": O.METRIC_L2, "dot": O.METRIC_DOT}


def _np(t):
    return t.detach().cpu().numpy()


def _oracle_ranks(Q, G, metric, precision, gt=None, row_offset=0):
    if precision == "bf16":
        Q, G = O.bf16_round(Q), O.bf16_round(G)
    return O.rank0_exact(Q, G, gt=gt, metric=METRICS[metric], row_offset=row_offset)


@pytest.fixture(params=[64, 16], ids=["cols64", "cols16"])
def fold(monkeypatch, request):
    """VTC_RANK_FOLD / VTC_FOLD_COLS are read by the library on every vtc_sim_rank call.  64-column
    fold operands are laid out like any other k-block; 16-column ones rely on TMA zero-filling the
    out-of-range part of the {64, rows} box (4 KB instead of 16 KB of L2 traffic per tile)."""
    monkeypatch.setenv("VTC_RANK_FOLD", "1")
    monkeypatch.setenv("VTC_FOLD_COLS", str(request.param))


# resident query tile (K' <= 512) and streamed (D = 768; every exact-mode K' = 3D), ragged edges
@pytest.mark.parametrize("precision", ["exact", "bf16"])
@pytest.mark.parametrize("metric", ["l2", "dot"])
@pytest.mark.parametrize("N,M,D,sigma", [(1000, 1000, 512, 6.0), (333, 1201, 96, 2.0),
                                         (129, 257, 768, 7.0), (300, 3000, 64, 1.5),
                                         (2000, 5000, 256, 5.0)])
def test_fold_rank_bit_exact(cuda_dev, fold, precision, metric, N, M, D, sigma):
    from vtc_b200 import ops

    T, V = make_retrieval_pair(min(N, M), M, D, sigma=sigma, seed=N + M)
    T = T[:N].contiguous()
    rank0, gts = ops.sim_rank(T.to(cuda_dev), V.to(cuda_dev), metric=metric, precision=precision)
    hits, medr = ops.rank_finalize(rank0, gts, M, [1, 5, 10])
    want = _oracle_ranks(T, V, metric, precision)
    np.testing.assert_array_equal(_np(rank0), want)
    np.testing.assert_array_equal(_np(hits), [np.sum(want < k) for k in (1, 5, 10)])
    assert _np(medr)[0] == O.medr(want)


@pytest.mark.parametrize("precision", ["exact", "bf16"])
def test_fold_same_ranks_as_default_path(cuda_dev, monkeypatch, precision):
    """Mixed-difficulty queries, shuffled gallery (explicit gt), non-unit gallery rows."""
    from vtc_b200 import ops

    T, V = make_retrieval_pair(1500, 2500, 512, seed=11, mixed=True)
    V = V * (0.5 + torch.rand(2500, 1, generator=torch.Generator().manual_seed(5)))  # norms 0.5..1.5
    perm = torch.randperm(2500, generator=torch.Generator().manual_seed(3))
    Vp = V[perm].contiguous()
    inv = torch.empty(2500, dtype=torch.int64)
    inv[perm] = torch.arange(2500)
    gt = inv[:1500].contiguous()
    q, g, gtd = T.to(cuda_dev), Vp.to(cuda_dev), gt.to(cuda_dev)
    monkeypatch.delenv("VTC_RANK_FOLD", raising=False)
    base, _ = ops.sim_rank(q, g, gt=gtd, precision=precision)
    monkeypatch.setenv("VTC_RANK_FOLD", "1")
    got, _ = ops.sim_rank(q, g, gt=gtd, precision=precision)
    np.testing.assert_array_equal(_np(got), _np(base))
    np.testing.assert_array_equal(_np(got), _oracle_ranks(T, Vp, "l2", precision, gt=gt.numpy()))


@pytest.mark.parametrize("precision", ["exact", "bf16"])
def test_fold_adversarial_golden(cuda_dev, fold, golden, precision):
    """Duplicated gallery rows (exact ties), a zero row, a non-unit row, a NaN query."""
    from vtc_b200 import ops

    g = golden("retrieval_small.npz")
    Q, G = torch.from_numpy(g["queries"]), torch.from_numpy(g["gallery"])
    for metric, key in (("l2", "rank0"), ("dot", "rank0_dot")):
        rank0, gts = ops.sim_rank(Q.to(cuda_dev), G.to(cuda_dev), metric=metric, precision=precision)
        ops.rank_finalize(rank0, gts, G.shape[0], [1])
        want = _oracle_ranks(Q, G, metric, precision) if precision == "bf16" else g[key]
        np.testing.assert_array_equal(_np(rank0), want)


def test_fold_duplicates_nan_gallery_row_and_chunks(cuda_dev, fold):
    from vtc_b200 import ops

    # every pair ties: everything goes through the re-check list
    row = torch.randn(1, 128)
    G = row.repeat(600, 1).contiguous()
    Q = row.repeat(400, 1).contiguous()
    for precision in ("exact", "bf16"):
        rank0, _ = ops.sim_rank(Q.to(cuda_dev), G.to(cuda_dev), precision=precision)
        np.testing.assert_array_equal(_np(rank0), np.arange(400))
    # a gallery row with a NaN / an inf element: the fold operands cannot carry it, the call must
    # fall back to the canonical brute-force kernel and still agree with the oracle
    T, V = make_retrieval_pair(300, 900, 256, sigma=3.0, seed=8)
    V[17, 5] = float("nan")
    V[400, 0] = float("inf")
    for precision in ("exact", "bf16"):
        rank0, gts = ops.sim_rank(T.to(cuda_dev), V.to(cuda_dev), precision=precision)
        ops.rank_finalize(rank0, gts, 900, [1])
        np.testing.assert_array_equal(_np(rank0), _oracle_ranks(T, V, "l2", precision))
    # additive over gallery chunks (ground truth outside the chunk, col_offset, accumulate)
    T, V = make_retrieval_pair(500, 2000, 256, sigma=4.0, seed=21)
    q, g = T.to(cuda_dev), V.to(cuda_dev)
    for precision in ("exact", "bf16"):
        whole, gts = ops.sim_rank(q, g, precision=precision)
        acc = torch.zeros(500, dtype=torch.int32, device=cuda_dev)
        bounds = [0, 700, 701, 1500, 2000]
        for s, e in zip(bounds[:-1], bounds[1:]):
            ops.sim_rank(q, g[s:e].contiguous(), col_offset=s, precision=precision, gt_score=gts,
                         rank0=acc, accumulate=True)
        np.testing.assert_array_equal(_np(acc), _np(whole))
        np.testing.assert_array_equal(_np(whole), _oracle_ranks(T, V, "l2", precision))


def test_fold_10k_and_100k_slice(cuda_dev, fold):
    from vtc_b200 import ops

    T, V = make_retrieval_pair(10000, 10000, 512, seed=1023)
    q, g = T.to(cuda_dev), V.to(cuda_dev)
    for precision in ("exact", "bf16"):
        rank0, _ = ops.sim_rank(q, g, precision=precision)
        np.testing.assert_array_equal(_np(rank0), _oracle_ranks(T, V, "l2", precision))
    N = M = 100_000
    T, V = make_retrieval_pair(N, M, 512, seed=1023)
    q, g = T.to(cuda_dev), V.to(cuda_dev)
    rank_bf16, gts = ops.sim_rank(q, g, precision="bf16")
    sl = slice(50_000, 50_256)
    want = _oracle_ranks(T[sl], V, "l2", "bf16", row_offset=50_000)
    np.testing.assert_array_equal(_np(rank_bf16[sl]), want)
    os.environ.pop("VTC_RANK_FOLD")
    base, _ = ops.sim_rank(q, g, precision="bf16")
    os.environ["VTC_RANK_FOLD"] = "1"
    np.testing.assert_array_equal(_np(rank_bf16), _np(base))


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


# VTC_FAST_THR=1 -- the guard-band thresholds from a coalesced fp32 norm of the query rows (an upper
# bound of ||q|| is all the band needs) instead of gt_score_kernel's second fp64 walk; alone and
# together with the fold epilogue, incl. the chunked (gt_score given) calls it helps most
@pytest.mark.parametrize("fold_on", [False, True])
@pytest.mark.parametrize("precision", ["exact", "bf16"])
def test_fast_thresholds_same_ranks(cuda_dev, monkeypatch, golden, precision, fold_on):
    from vtc_b200 import ops

    monkeypatch.setenv("VTC_FAST_THR", "1")
    monkeypatch.setenv("VTC_RANK_FOLD", "1" if fold_on else "0")
    for N, M, D, metric in ((1000, 1000, 512, "l2"), (333, 1201, 96, "dot"), (129, 257, 768, "l2")):
        T, V = make_retrieval_pair(N, M, D, sigma=4.0, seed=N + M)
        rank0, gts = ops.sim_rank(T.to(cuda_dev), V.to(cuda_dev), metric=metric, precision=precision)
        ops.rank_finalize(rank0, gts, M, [1])
        np.testing.assert_array_equal(_np(rank0), _oracle_ranks(T, V, metric, precision))
    g = golden("retrieval_small.npz")           # ties, zero row, non-unit row, NaN query
    Q, G = torch.from_numpy(g["queries"]), torch.from_numpy(g["gallery"])
    rank0, gts = ops.sim_rank(Q.to(cuda_dev), G.to(cuda_dev), precision=precision)
    ops.rank_finalize(rank0, gts, G.shape[0], [1])
    want = _oracle_ranks(Q, G, "l2", precision) if precision == "bf16" else g["rank0"]
    np.testing.assert_array_equal(_np(rank0), want)
    T, V = make_retrieval_pair(500, 2000, 256, sigma=4.0, seed=21)
    q, gal = T.to(cuda_dev), V.to(cuda_dev)
    whole, gts = ops.sim_rank(q, gal, precision=precision)
    acc = torch.zeros(500, dtype=torch.int32, device=cuda_dev)
    for s, e in ((0, 700), (700, 701), (701, 2000)):
        ops.sim_rank(q, gal[s:e].contiguous(), col_offset=s, precision=precision, gt_score=gts,
                     rank0=acc, accumulate=True)
    np.testing.assert_array_equal(_np(acc), _np(whole))
    np.testing.assert_array_equal(_np(whole), _oracle_ranks(T, V, "l2", precision))


# vtc_rank_prepare / vtc_sim_rank_prepared -- norms, norm bounds and ground-truth scores computed
# once per chunk and handed to every call that touches the chunk (new entry points; used by the
# host-staging pipeline and the sharded step when VTC_RANK_PREPARED=1)
@pytest.mark.parametrize("fold_on", [False, True])
@pytest.mark.parametrize("precision", ["exact", "bf16"])
def test_prepared_ranking_same_ranks(cuda_dev, monkeypatch, golden, precision, fold_on):
    from vtc_b200 import ops

    monkeypatch.setenv("VTC_RANK_FOLD", "1" if fold_on else "0")

    def canon(x):  # the rows must already be the canonical values of the mode
        return x.to(cuda_dev).bfloat16() if precision == "bf16" else x.to(cuda_dev)

    cases = [make_retrieval_pair(1000, 1000, 512, sigma=4.0, seed=3) + ("l2",),
             make_retrieval_pair(333, 1201, 96, sigma=2.0, seed=4) + ("dot",),
             make_retrieval_pair(129, 257, 768, sigma=7.0, seed=5) + ("l2",)]
    g = golden("retrieval_small.npz")           # ties, zero row, non-unit row, NaN query
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
        # whole gallery in one prepared call
        rank0 = torch.full((N,), -5, dtype=torch.int32, device=cuda_dev)
        ops.sim_rank(q, gal, metric=metric, precision=precision, gt_score=gts, rank0=rank0,
                     accumulate=False, sq64=sq64, qq=qq)
        ops.rank_finalize(rank0, gts, M, [1])
        np.testing.assert_array_equal(_np(rank0), want)
        # ... and accumulated over ragged gallery chunks with slices of the same arrays
        acc = torch.zeros(N, dtype=torch.int32, device=cuda_dev)
        bounds = [0, M // 3, M // 3 + 1, M]
        for s, e in zip(bounds[:-1], bounds[1:]):
            ops.sim_rank(q, gal[s:e].contiguous(), col_offset=s, metric=metric, precision=precision,
                         gt_score=gts, rank0=acc, accumulate=True, sq64=sq64[s:e].contiguous(), qq=qq)
        ops.rank_finalize(acc, gts, M, [1])
        np.testing.assert_array_equal(_np(acc), want)


@pytest.mark.parametrize("schedule", ["equal", "balanced"])
def test_prepared_host_staging_pipeline(cuda_dev, monkeypatch, schedule):
    """RecallAtK.compute from host arrays with VTC_RANK_PREPARED=1 (and the balanced chunk schedule):
    same recalls and ranks as the oracle."""
    from vtc_b200.model.metric import RecallAtK

    monkeypatch.setenv("VTC_RANK_PREPARED", "1")
    monkeypatch.setenv("VTC_PIPELINE_SCHEDULE", schedule)
    N, D = 40_000, 256                           # 41 MB per side: above PIPELINE_MIN_BYTES
    T, V = make_retrieval_pair(N, N, D, sigma=5.0, seed=12)
    for precision in ("bf16", "exact"):
        m = RecallAtK("v", "t", [1, 5, 10], precision=precision)
        full = m.compute_full(V.numpy(), T.numpy())
        sl = slice(20_000, 20_200)
        want = _oracle_ranks(T[sl], V, "l2", precision, row_offset=20_000)
        np.testing.assert_array_equal(_np(full["rank0"][sl]), want)
        monkeypatch.setenv("VTC_RANK_PREPARED", "0")
        base = m.compute_full(V.numpy(), T.numpy())
        monkeypatch.setenv("VTC_RANK_PREPARED", "1")
        np.testing.assert_array_equal(_np(full["rank0"]), _np(base["rank0"]))
        np.testing.assert_array_equal(_np(full["hits"]), _np(base["hits"]))


@pytest.mark.parametrize("precision", ["brute", "exact", "bf16"])
def test_default_path_infinite_rows(cuda_dev, monkeypatch, precision):
    """Not a variant: an edge case of the DEFAULT path that no committed fixture holds yet -- gallery
    rows of -inf (the padding rows evaluation/retrieval_evaluation.py:238-252 creates for missing
    captions) and +inf, also as ground truth.  Canonical arithmetic gives such a column an infinite or
    NaN score (never closer than a finite one; NaN comparisons false); kept here until it has passed
    on a GPU once, then it moves to tests/test_gpu_parity.py."""
    from vtc_b200 import ops

    for v in ("VTC_RANK_FOLD", "VTC_FAST_THR", "VTC_RANK_PREPARED"):
        monkeypatch.delenv(v, raising=False)
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
