"""CPU check of the arithmetic behind the opt-in EPI_RANK_FOLD pass (vtc_b200/csrc/fold.cu,
RankFoldEpi in csrc/sim_tc_kernel.cuh): the fold operands, the sign convention and the guard band
are restated in numpy exactly as the CUDA code builds them, the accumulator is modelled as an fp32
value within the tensor-core error bound, and every decision the epilogue would take WITHOUT the
fp64 re-check is compared with the canonical fp64-sequential comparison d(t,j) < d(t,gt) that
replaces the faiss search + hit loop of model/metric.py:137-161.  It validates the math, not the
kernel: the kernel's own parity tests are tests/test_gpu_experimental.py."""
import numpy as np
import pytest
import torch

from oracle import vtc_oracle as O
from vtc_b200.synthetic import make_retrieval_pair


def _bf16(x64: np.ndarray) -> np.ndarray:
    """RN-even bf16 rounding of (float)x, returned as float64 (what __float2bfloat16_rn does)."""
    return torch.from_numpy(x64.astype(np.float32)).bfloat16().double().numpy()


def _split3(v: np.ndarray):
    rem = v.astype(np.float64).copy()
    pieces = []
    for _ in range(3):
        p = _bf16(rem)
        pieces.append(p)
        rem = rem - p
    return pieces, rem


def _guard_rel(kp: int) -> float:  # api.cu::guard_rel_for, VTC_PREC_BF16
    return (kp // 16 + 8) * 2.0 ** -24


@pytest.mark.parametrize("metric", ["l2", "dot"])
@pytest.mark.parametrize("scale_rows", [False, True])
def test_fold_decisions_agree_with_canonical_comparison(metric, scale_rows):
    N, M, D = 300, 1200, 128
    T, V = make_retrieval_pair(N, M, D, sigma=3.0, seed=5, mixed=True)
    if scale_rows:  # non-unit gallery rows (mean of unit vectors in the reference's eval tail)
        V = V * (0.4 + 1.2 * torch.rand(M, 1, generator=torch.Generator().manual_seed(1)))
    Q, G = O.bf16_round(T), O.bf16_round(V)  # canonical values of the bf16 mode
    met = O.METRIC_L2 if metric == "l2" else O.METRIC_DOT
    d = O.scores64(Q, G, met)                # canonical d(t, j), fp64-sequential
    sq = O.sqnorm64(G)
    gt = np.arange(N)
    d0 = d[gt, gt]
    Q64, G64 = np.asarray(Q, dtype=np.float64), np.asarray(G, dtype=np.float64)
    dot = Q64 @ G64.T
    qn = np.sqrt((Q64 * Q64).sum(1))
    gmax_sq = float(np.float32(sq.max()))
    gn = np.sqrt(gmax_sq)
    g = _guard_rel(-(-D // 64) * 64)

    # exact.cu::gt_score_kernel: the (lo, hi) thresholds of the default epilogue, rounded outwards
    if metric == "l2":
        delta = 2.0 * g * qn * gn + 2.4e-7 * (gmax_sq + 2.0 * qn * gn)
    else:
        delta = g * qn * gn + 1.2e-7 * qn * gn
    lo = np.nextafter((d0 - delta).astype(np.float32), np.float32(-np.inf))
    hi = np.nextafter((d0 + delta).astype(np.float32), np.float32(np.inf))

    # fold.cu::fold_q_kernel / fold_g_kernel
    dl = 0.5 * (hi.astype(np.float64) - lo.astype(np.float64))
    assert (dl >= delta).all()
    if metric == "l2":
        m = 0.5 * d0
        w = 0.5 * dl + 4.8e-7 * (dl / (2.0 * g) + 0.5 * gmax_sq + np.abs(m))
        h = -0.5 * sq
    else:
        m = d0.copy()
        w = dl + 4.8e-7 * (dl / g + np.abs(m))
        h = np.zeros(M)
    w = (w * (1.0 + 1e-6)).astype(np.float32)
    mp, mrem = _split3(m)
    hp, hrem = _split3(h)
    assert np.abs(mrem).max() <= 2.0 ** -24 * max(np.abs(m).max(), 1e-30)
    assert np.abs(hrem).max() <= 2.0 ** -24 * max(np.abs(h).max(), 1e-30)

    # the accumulator after the fold step: q.x + (m0+m1+m2)*1 + 1*(h0+h1+h2), as an fp32 value that
    # carries the worst-case tensor-core error the guard band is built for, in both directions
    fold_term = (mp[0] + mp[1] + mp[2])[:, None] + (hp[0] + hp[1] + hp[2])[None, :]
    err = (g * qn * gn)[:, None] + 8 * 2.0 ** -24 * (np.abs(dot) + np.abs(fold_term))
    closer = (d < d0[:, None]) | ((d == d0[:, None]) & (np.arange(M)[None, :] < gt[:, None]))
    notgt = np.arange(M)[None, :] != gt[:, None]
    in_band_total = 0
    for sign in (-1.0, 1.0):
        acc = (dot + fold_term + sign * err).astype(np.float32)
        decided = np.abs(acc) > w[:, None]
        # RankFoldEpi: outside the band a column counts iff its sign bit is clear
        assert ((acc > 0)[decided & notgt] == closer[decided & notgt]).all()
        # the ground truth's own column is inside the band by construction
        assert (np.abs(acc[gt, gt]) <= w).all()
        in_band_total += int((~decided & notgt).sum())
    # ... and the band is narrow: what goes to the fp64 re-check is a sliver of the matrix
    assert in_band_total < 2e-3 * N * M
    # padding columns (h = -1e30) and NaN rows (m' = -1e30) are never closer and never in the band
    far = np.float32(-1.0e30)
    pieces, _ = _split3(np.array([-1.0e30]))
    assert np.isfinite(sum(pieces)).all() and np.float32(sum(pieces)[0] + 2.0) < -w.max()
    assert far < 0


@pytest.mark.parametrize("metric", ["l2", "dot"])
def test_fold_epilogue_algorithm_end_to_end(metric):
    """The whole EPI_RANK_FOLD algorithm transcribed to numpy -- fold operands incl. the -1e30
    padding columns, fp32 accumulator, 8-column groups counted by sign bits, the rare path
    (group decided iff its only in-band column is the ground truth, else handed to the canonical
    re-check), duplicated gallery rows (exact ties) and a NaN query -- must reproduce the oracle's
    ranks exactly.  Mirrors RankFoldEpi::chunk / slow_group and exact.cu::recheck_kernel."""
    N, M, D = 96, 1000, 64   # M is not a multiple of 8 * anything nice: 1000 = 3 full tiles + 232
    T, V = make_retrieval_pair(N, M, D, sigma=2.5, seed=17, mixed=True)
    V[500] = V[3]            # exact duplicates of two ground-truth rows, one on either side
    V[1] = V[40]
    T[7, 5] = float("nan")
    Q, G = O.bf16_round(T), O.bf16_round(V)
    met = O.METRIC_L2 if metric == "l2" else O.METRIC_DOT
    want = O.rank0_exact(Q, G, metric=met)
    d = O.scores64(Q, G, met)
    sq = O.sqnorm64(G)
    gt = np.arange(N)
    d0 = d[gt, gt]
    Q64, G64 = np.asarray(Q, dtype=np.float64), np.asarray(G, dtype=np.float64)
    qn = np.sqrt(np.nansum(Q64 * Q64, 1))
    gmax_sq = float(np.float32(sq.max()))
    gn = np.sqrt(gmax_sq)
    g = _guard_rel(64)
    delta = (2.0 * g * qn * gn + 2.4e-7 * (gmax_sq + 2.0 * qn * gn)) if metric == "l2" \
        else (g * qn * gn + 1.2e-7 * qn * gn)
    lo = np.nextafter((d0 - delta).astype(np.float32), np.float32(-np.inf)).astype(np.float64)
    hi = np.nextafter((d0 + delta).astype(np.float32), np.float32(np.inf)).astype(np.float64)
    dl = 0.5 * (hi - lo)
    ok = np.isfinite(d0)
    m = np.where(ok, 0.5 * d0 if metric == "l2" else d0, -1.0e30)
    if metric == "l2":
        w = 0.5 * dl + 4.8e-7 * (dl / (2.0 * g) + 0.5 * gmax_sq + np.abs(m))
        h = -0.5 * sq
    else:
        w = dl + 4.8e-7 * (dl / g + np.abs(m))
        h = np.zeros(M)
    w = np.where(ok, (w * (1.0 + 1e-6)), -1.0).astype(np.float32)
    Mpad = -(-M // 256) * 256
    hpad = np.concatenate([h, np.full(Mpad - M, -1.0e30)])
    mp, _ = _split3(m)
    hp, _ = _split3(hpad)
    Gpad = np.concatenate([G64, np.zeros((Mpad - M, D))])       # TMA zero-fills the padding rows
    with np.errstate(invalid="ignore", over="ignore"):
        acc = (Q64 @ Gpad.T + (mp[0] + mp[1] + mp[2])[:, None]
               + (hp[0] + hp[1] + hp[2])[None, :]).astype(np.float32)
    neg = np.signbit(acc)                                        # x >> 31
    rank = np.zeros(N, dtype=np.int64)
    pushed = 0
    for t in range(N):
        for j0 in range(0, Mpad, 8):
            x = acc[t, j0:j0 + 8]
            with np.errstate(invalid="ignore"):
                mn = np.fmin.reduce(np.abs(x))                   # FMNMX ignores NaN operands
            if not (mn <= w[t]):                                 # common path (and all-NaN rows)
                rank[t] += 8 - int(neg[t, j0:j0 + 8].sum())
                continue
            in_band = np.abs(x) <= w[t]
            gi = t - j0
            if 0 <= gi < 8 and in_band.sum() == 1 and in_band[gi]:
                rank[t] += int((~neg[t, j0:j0 + 8]).sum()) - int(not neg[t, j0 + gi])
                continue
            pushed += 1                                          # recheck_kernel, canonical fp64
            for jl in range(j0, min(j0 + 8, M)):
                if jl != t and (d[t, jl] < d0[t] or (d[t, jl] == d0[t] and jl < t)):
                    rank[t] += 1
    rank[~ok] = M                                                # vtc_rank_finalize: NaN score
    np.testing.assert_array_equal(rank, want)
    assert pushed < 0.01 * N * Mpad / 8
