"""Seeded synthetic embeddings for the retrieval hot path.

The reference ships no data (`.MISSING_LARGE_BLOBS`) and there is no network, so every test and
benchmark uses this recipe: a unit-norm gallery and queries that are noisy copies of the first N
gallery rows, so that ground truth is "same index" (model/metric.py:149-160) and R@K is neither
0 nor 1.  Seed 1023 is the reference's default (`train.py:34`).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

DEFAULT_SEED = 1023


def _normalize(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(dim=-1, keepdim=True)


def default_sigma(D: int) -> float:
    """sigma=6 for D=512, sigma=7 for D=768 (SURVEY.md §8d calibration)."""
    return 6.0 if D <= 512 else 7.0


def make_retrieval_pair(N: int, M: int, D: int, sigma: Optional[float] = None, seed: int = DEFAULT_SEED,
                        mixed: bool = False, chunk: int = 65536) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (queries [N,D], gallery [M,D]) fp32 CPU tensors, both unit-norm.

    gallery  V = normalize(randn(M, D))
    queries  T = normalize(V[:N] + sigma * randn(N, D) / sqrt(D))      (N <= M)
    ``mixed`` draws a per-query sigma_i ~ U(2, 12) so ranks spread from 0 to thousands.
    """
    assert N <= M, "queries are noisy copies of the first N gallery rows"
    g = torch.Generator().manual_seed(seed)
    V = torch.empty(M, D)
    for s in range(0, M, chunk):
        e = min(M, s + chunk)
        V[s:e] = _normalize(torch.randn(e - s, D, generator=g))
    sig = default_sigma(D) if sigma is None else float(sigma)
    T = torch.empty(N, D)
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        noise = torch.randn(e - s, D, generator=g) / math.sqrt(D)
        if mixed:
            si = 2.0 + 10.0 * torch.rand(e - s, 1, generator=g)
        else:
            si = sig
        T[s:e] = _normalize(V[s:e] + si * noise)
    return T, V


def make_batch_pair(b: int, D: int, seed: int = DEFAULT_SEED, sigma: float = 8.0
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Unit-norm (vis, text) batch for InfoNCE: text is a noisy copy of vis."""
    T, V = make_retrieval_pair(b, b, D, sigma=sigma, seed=seed)
    return V, T


def make_cam_inputs(b: int, nc: int, D: int, seed: int = DEFAULT_SEED
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(main [b,D], aux [nc,b,D]) plain randn, un-normalised."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(b, D, generator=g), torch.randn(nc, b, D, generator=g)
