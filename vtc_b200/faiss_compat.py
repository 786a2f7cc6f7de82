"""The three faiss names the reference's retrieval metric calls, on the B200 library.

`model/metric.py` of the reference builds its index in exactly four lines:

    self.knn_config = faiss.GpuIndexFlatConfig(); .useFloat16 = False            (:112-113)
    self.knn_config.device = ...                                                 (:127)
    index = faiss.GpuIndexFlatL2(faiss.StandardGpuResources(), num_dims, cfg)    (:139-142)
    index.add(features_a); _, I = index.search(features_b, max(k_vals) + 1)      (:143-146)

`vtc_b200.model.metric.RecallAtK` replaces that whole method (fused similarity + rank, no top-k list
at all).  This module is the other way to switch: keep the reference's `model/metric.py` UNMODIFIED
and bind its `faiss` import to this module -- `import vtc_b200.faiss_compat as faiss`, or
`sys.modules["faiss"] = vtc_b200.faiss_compat` before `model.metric` is imported.  `search` is the
fused tcgen05 similarity + streaming top-k kernel (`vtc_sim_topk`, include/vtc_b200.h): the N x M
distance matrix is never formed, and the k best are exact -- `useFloat16 = False` selects
VTC_PREC_EXACT (fp32 inputs ranked like faiss' exact fp32 search: ascending squared L2, ties by the
lower index), `useFloat16 = True` VTC_PREC_BF16 (exact on the bf16-rounded rows).

Only what the reference touches is provided (flat L2 index: add / search / reset / ntotal / d); there
is no CPU fallback -- without a CUDA device the calls raise.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import ops

MAX_K = 16  # vtc_sim_topk keeps at most 16 candidates per row (the reference asks for max(k_vals)+1 = 11)


class GpuIndexFlatConfig:
    """Attribute bag like faiss.GpuIndexFlatConfig: the reference sets `useFloat16` and `device`."""

    def __init__(self):
        self.useFloat16 = False
        self.device = 0


class StandardGpuResources:
    """faiss' scratch-memory owner.  The library's workspace is owned by the caller (vtc_b200.ops
    keeps one per device), so there is nothing to hold."""

    def noTempMemory(self) -> None:  # noqa: N802  (faiss spelling)
        pass

    def setTempMemory(self, nbytes: int) -> None:  # noqa: N802
        pass


def _as_rows(x, d: int, name: str) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor) or x.dim() != 2 or x.shape[1] != d:
        raise ValueError(f"{name} must be a 2-D [rows, {d}] array")
    return x


class GpuIndexFlatL2:
    """Exact brute-force squared-L2 index on one GPU (faiss.GpuIndexFlatL2's add / search)."""

    def __init__(self, resources: Optional[StandardGpuResources], dims: int,
                 config: Optional[GpuIndexFlatConfig] = None):
        if int(dims) <= 0:
            raise ValueError("dims must be positive")
        config = config or GpuIndexFlatConfig()
        if not torch.cuda.is_available():
            raise ops.VtcError("GpuIndexFlatL2 needs a CUDA device: vtc_b200 has no CPU path")
        self.d = int(dims)
        self.precision = "bf16" if getattr(config, "useFloat16", False) else "exact"
        self.device = torch.device("cuda", int(getattr(config, "device", 0) or 0))
        self.is_trained = True
        self._rows = []        # device tensors in insertion order
        self._gallery = None   # their concatenation, built by the first search after an add

    @property
    def ntotal(self) -> int:
        return sum(int(r.shape[0]) for r in self._rows)

    def _stage(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(self.device, non_blocking=True)
        if self.precision == "bf16":
            # the bf16 mode ranks the RN-even bf16 roundings: round once here; bf16 rows of whole
            # 128-byte swizzle atoms are the tensor-core operands in place (no per-search re-prep)
            return x.to(torch.bfloat16).contiguous()
        return x.float().contiguous()

    def add(self, x) -> None:
        x = _as_rows(x, self.d, "x")
        if x.shape[0]:
            self._rows.append(self._stage(x))
            self._gallery = None

    def reset(self) -> None:
        self._rows, self._gallery = [], None

    def search(self, x, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """(D fp32 [n, k] ascending squared L2, I int64 [n, k]); -1 / +inf beyond ntotal, like faiss.
        numpy in -> numpy out (what the reference passes, model/metric.py:170-171); device tensors in
        -> device tensors out, nothing synchronises."""
        k = int(k)
        if not 1 <= k <= MAX_K:
            raise ValueError(f"k must be in [1, {MAX_K}] (got {k})")
        as_numpy = isinstance(x, np.ndarray)
        q = self._stage(_as_rows(x, self.d, "x"))
        if self._gallery is None:
            self._gallery = (torch.cat(self._rows) if len(self._rows) != 1 else self._rows[0]) \
                if self._rows else torch.empty((0, self.d), dtype=q.dtype, device=self.device)
        vals, idx = ops.sim_topk(q, self._gallery, k, metric="l2", precision=self.precision)
        if as_numpy:
            return vals.cpu().numpy(), idx.cpu().numpy()
        return vals, idx


__all__ = ["GpuIndexFlatConfig", "StandardGpuResources", "GpuIndexFlatL2", "MAX_K"]
