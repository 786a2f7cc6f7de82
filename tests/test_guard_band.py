"""The guard band of the exact mode (csrc/api.cu::guard_rel_for), checked on the CPU.

VTC_PREC_EXACT feeds the tensor cores x = hi + lo (hi = bf16(x), lo = bf16(x - hi)) and issues
hi*hi + hi*lo + lo*hi.  What that drops must stay inside the band, or a tied / nearly tied column is
counted definitively instead of going to the fp64 re-check (ranks would no longer be bit-exact against
the fp32 search of model/metric.py:140-146).  bf16 has 8 significand bits, so its unit roundoff is
2^-8 and the dropped part is bounded by 3 * 2^-16 (1 + 2^-7) < 4.62e-5 of |q||x| -- NOT by 3 * 2^-18,
the constant of round 1.  The inputs below are the structured ones on which the old constant failed:
constant-magnitude rows whose elements sit just below a bf16 rounding midpoint.
"""
import numpy as np
import pytest

from oracle import vtc_oracle as O

SPLIT_TERM = 4.62e-5          # csrc/api.cu::guard_rel_for, VTC_PREC_EXACT
OLD_SPLIT_TERM = 1.2e-5       # round 1 (understated)


def split3_dot(q, x):
    """hi*hi + hi*lo + lo*hi in exact arithmetic (what the bf16x3 operands [hi|hi|lo].[hi|lo|hi]
    multiply; the accumulation error of the fp32 accumulator is the other term of the band)."""
    qh = O.bf16_round(q)
    xh = O.bf16_round(x)
    ql = O.bf16_round(q - qh)
    xl = O.bf16_round(x - xh)
    qh, ql, xh, xl = (a.astype(np.float64) for a in (qh, ql, xh, xl))
    return qh @ xh.T + qh @ xl.T + ql @ xh.T


def rel_err(q, x):
    exact = q.astype(np.float64) @ x.astype(np.float64).T
    norms = np.linalg.norm(q.astype(np.float64), axis=1)[:, None] * \
        np.linalg.norm(x.astype(np.float64), axis=1)[None, :]
    return np.abs(split3_dot(q, x) - exact) / norms


def adversarial_rows(rng, rows, D, value, base=None):
    """+-value with mostly agreeing signs between rows (cos ~ 0.64 .. 1): every product errs the
    same way."""
    if base is None:
        base = np.sign(rng.standard_normal(D)).astype(np.float32)
    flip = rng.random((rows, D)) < 0.1
    return (np.where(flip, -base, base) * np.float32(value)).astype(np.float32)


@pytest.mark.parametrize("D", [64, 512, 768])
def test_split_error_is_inside_the_new_band_and_outside_the_old_one(D):
    rng = np.random.default_rng(D)
    worst = 0.0
    # elements just below the midpoint between two bf16 values (spacing 2^-7 at 1.0): the residual
    # x - hi is as large as it gets, and so is what bf16(x - hi) drops
    for value in (1.00385, 1.0038, 1.00389, 1.0116, 0.50192, 1.99):
        base = np.sign(rng.standard_normal(D)).astype(np.float32)
        q = adversarial_rows(rng, 24, D, value, base)
        x = adversarial_rows(rng, 24, D, value, base)
        x[0] = q[0]
        worst = max(worst, rel_err(q, x).max())
    print(f"\n[guard band] D={D}: worst bf16x3 split error {worst:.3e} of |q||x| "
          f"(band {SPLIT_TERM:.2e}, round-1 band {OLD_SPLIT_TERM:.1e})")
    assert worst <= SPLIT_TERM
    assert worst > OLD_SPLIT_TERM, "these inputs are meant to break the understated constant"


def test_split_error_random_search_never_exceeds_the_band():
    rng = np.random.default_rng(7)
    worst = 0.0
    for _ in range(40):
        D = int(rng.choice([32, 96, 256]))
        scale = np.float32(2.0 ** rng.integers(-6, 6))
        kind = rng.integers(0, 3)
        if kind == 0:      # Gaussian rows
            q = rng.standard_normal((16, D)).astype(np.float32) * scale
            x = rng.standard_normal((16, D)).astype(np.float32) * scale
        elif kind == 1:    # quantised magnitudes near midpoints
            mags = np.float32(1.0) + np.float32(2.0 ** -8) * rng.uniform(0.9, 1.0, (16, D)).astype(np.float32)
            q = mags * np.sign(rng.standard_normal((16, D))).astype(np.float32) * scale
            x = mags[::-1] * np.sign(rng.standard_normal((16, D))).astype(np.float32) * scale
        else:              # near-duplicate rows
            q = adversarial_rows(rng, 16, D, 1.00385 * scale)
            x = q + np.float32(1e-4) * scale * rng.standard_normal((16, D)).astype(np.float32)
        worst = max(worst, rel_err(q, x).max())
    assert worst <= SPLIT_TERM


# ---- the per-row bound the rank path uses since round 2 (csrc/sim_tc.cuh::rank_split_bound) -------
def split_pieces(x):
    """hi = bf16(x), lo = bf16(x - hi), e = x - hi - lo (each difference is exact in fp32)."""
    hi = O.bf16_round(x)
    r = (x - hi).astype(np.float32)
    lo = O.bf16_round(r)
    e = (r - lo).astype(np.float32)
    return hi, lo, e


def row_bound(q, x):
    """|q| max|ex| + |lq| max|lx| + |eq| (max|x| + max|ex|) for every query row, in float64 from the
    pieces themselves (the kernel computes the same norms in fp32 and rounds them up)."""
    _, lq, eq = split_pieces(q)
    _, lx, ex = split_pieces(x)
    n = lambda a: np.linalg.norm(a.astype(np.float64), axis=1)  # noqa: E731
    return n(q) * n(ex).max() + n(lq) * n(lx).max() + n(eq) * (n(x).max() + n(ex).max())


@pytest.mark.parametrize("D", [64, 512, 768])
def test_per_row_split_bound_holds_on_adversarial_and_random_rows(D):
    rng = np.random.default_rng(100 + D)
    cases = []
    for value in (1.00385, 1.0038, 1.00389, 1.0116, 0.50192, 1.99):
        base = np.sign(rng.standard_normal(D)).astype(np.float32)
        q = adversarial_rows(rng, 24, D, value, base)
        x = adversarial_rows(rng, 24, D, value, base)
        x[0] = q[0]
        cases.append((q, x))
    for scale in (2.0 ** -5, 1.0, 37.0):
        q = (rng.standard_normal((24, D)) * scale).astype(np.float32)
        x = (rng.standard_normal((24, D)) * scale).astype(np.float32)
        cases.append((q / np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32), x))
    tightest = np.inf
    for q, x in cases:
        exact = q.astype(np.float64) @ x.astype(np.float64).T
        err = np.abs(split3_dot(q, x) - exact)
        bound = row_bound(q, x)[:, None]
        assert (err <= bound * (1 + 1e-12)).all()
        norms = np.linalg.norm(q.astype(np.float64), axis=1)[:, None] * \
            np.linalg.norm(x.astype(np.float64), axis=1).max()
        tightest = min(tightest, (bound / norms).min())
    # on ordinary rows the bound is an order of magnitude below the worst-case constant: that is
    # what cuts the fp64 re-check of the exact mode
    print(f"\n[guard band] D={D}: smallest per-row split bound {tightest:.2e} of |q| max|x| "
          f"(constant {SPLIT_TERM:.2e})")
    assert tightest < SPLIT_TERM / 5
