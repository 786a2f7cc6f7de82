"""Drop-in for the reference's model/loss.py (only `clip_loss` is on the hot path; the other
three losses there are unused by every shipped config -- SURVEY.md §2 row 1).

`clip_loss(input, meta)` keeps the reference signature (model/loss.py:18-22): it reads only
`input[2]`.  When that is a :class:`LazySim` (what this package's model forward returns) the
N x N logit matrix is never formed: one fused tcgen05 pass per direction yields the row / column
log-sum-exp and the diagonal (csrc/sim_tc.cu, EPI_LSE).  When it is a materialised CUDA tensor the
same reductions run over that tensor (csrc/infonce_dense.cu).  Both are differentiable.
"""
from __future__ import annotations

import torch

from .. import ops
from .lazy import LazySim

__all__ = ["clip_loss", "LazySim"]


class _FusedClipLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats_a, feats_b, scale, precision):
        a = feats_a.detach()
        b = feats_b.detach()
        loss, row, col, _diag = ops.infonce_fwd(a, b, scale, precision)
        ctx.save_for_backward(a, b, scale.detach() if isinstance(scale, torch.Tensor) else
                              torch.tensor(float(scale), device=a.device), row, col)
        ctx.scale_is_tensor = isinstance(scale, torch.Tensor)
        ctx.precision = precision
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        a, b, scale, row, col = ctx.saved_tensors
        dA, dB, ds = ops.infonce_bwd(a, b, scale, row, col, grad_out, ctx.precision)
        dscale = ds.reshape(scale.shape).to(scale.dtype) if ctx.scale_is_tensor else None
        return dA.to(a.dtype), dB.to(b.dtype), dscale, None


class _DenseClipLoss(torch.autograd.Function):
    """clip_loss of a materialised sim: three reductions of the matrix, elementwise backward."""

    @staticmethod
    def forward(ctx, sim):
        s = sim.detach().float().contiguous()
        loss, row, col, _diag = ops.infonce_dense_fwd(s)
        ctx.save_for_backward(s, row, col)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        s, row, col = ctx.saved_tensors
        return ops.infonce_dense_bwd(s, row, col, grad_out)


def clip_loss(input, meta=None, precision: str = None, **loss_args):
    """0.5 * (CE(sim, arange) + CE(sim.t(), arange)) -- model/loss.py:18-22."""
    sim = input[2]
    if isinstance(sim, LazySim):
        prec = precision or sim.precision
        if sim.feats_a.shape != sim.feats_b.shape:
            raise ValueError("clip_loss requires a square similarity matrix")
        return _FusedClipLoss.apply(sim.feats_a, sim.feats_b, sim.scale, prec)
    if not isinstance(sim, torch.Tensor):
        raise TypeError("input[2] must be a LazySim or a tensor")
    if not sim.is_cuda:
        raise ops.VtcError("clip_loss: vtc_b200 has no CPU path; move `sim` to a CUDA device")
    if sim.dim() != 2 or sim.shape[0] != sim.shape[1]:
        raise ValueError("clip_loss requires a square similarity matrix")
    # a materialised sim (what the reference's forward hands over): reduce the matrix itself
    return _DenseClipLoss.apply(sim).to(sim.dtype)
