"""LazySim: the `sim` a model forward returns without forming the N x M logit matrix.

The reference's forward returns `(feats_vis, feats_text, sim)` with
`sim = logit_scale.exp() * feats_vis @ feats_text.t()` (model/model.py:369,478,504,621) and
`clip_loss` reads only that third element (model/loss.py:19).  Returning a LazySim keeps the call
sites unchanged while letting `clip_loss` run the fused kernel; anything else that touches it as a
tensor (torch functions, attributes, indexing) transparently materialises it on the GPU through
vtc_sim_matrix.
"""
from __future__ import annotations

import torch

from .. import ops


class _SimMatrixFunction(torch.autograd.Function):
    """sim = (s * A) @ B^T with gradients to A, B and s (dA = s g B, dB = s g^T A,
    ds = sum(g * A B^T)), every product on the library's tensor-core GEMM."""

    @staticmethod
    def forward(ctx, a, b, scale, precision):
        sim = ops.sim_matrix(a, b, scale, precision)
        ctx.save_for_backward(a, b, scale if isinstance(scale, torch.Tensor) else torch.tensor(float(scale)))
        ctx.precision = precision
        ctx.sim = sim
        return sim

    @staticmethod
    def backward(ctx, g):
        a, b, scale = ctx.saved_tensors
        s = scale.detach().to(device=g.device, dtype=torch.float32).reshape(())
        g = g.contiguous().float()
        da = db = ds = None
        if ctx.needs_input_grad[0]:
            da = (s * ops.linear(g, b.detach().float().t().contiguous(), None, None, 0, ctx.precision)
                  ).to(a.dtype)
        if ctx.needs_input_grad[1]:
            db = (s * ops.linear(g.t().contiguous(), a.detach().float().t().contiguous(), None, None, 0,
                                 ctx.precision)).to(b.dtype)
        if ctx.needs_input_grad[2]:
            ds = ((g * ctx.sim).sum() / s).reshape(scale.shape).to(scale.dtype)
        return da, db, ds, None


class LazySim:
    def __init__(self, feats_a: torch.Tensor, feats_b: torch.Tensor, scale, precision: str = "exact"):
        self.feats_a = feats_a
        self.feats_b = feats_b
        self.scale = scale
        self.precision = precision
        self._dense = None

    # cheap metadata without materialising
    @property
    def shape(self):
        return torch.Size((self.feats_a.shape[0], self.feats_b.shape[0]))

    @property
    def device(self):
        return self.feats_a.device

    @property
    def dtype(self):
        return torch.float32

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 2

    def materialize(self) -> torch.Tensor:
        """(scale * A) @ B.t() as a real fp32 CUDA tensor.  Like the reference's `sim`
        (model/model.py:369) it is differentiable: when autograd is recording and the features or
        the scale require grad, the tensor carries a graph back to them."""
        wants_grad = torch.is_grad_enabled() and any(
            isinstance(t, torch.Tensor) and t.requires_grad
            for t in (self.feats_a, self.feats_b, self.scale))
        if wants_grad:
            if self._dense is None or not self._dense.requires_grad:
                self._dense = _SimMatrixFunction.apply(self.feats_a, self.feats_b, self.scale,
                                                       self.precision)
            return self._dense
        if self._dense is None:
            self._dense = ops.sim_matrix(self.feats_a.detach(), self.feats_b.detach(), self.scale,
                                         self.precision)
        return self._dense

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        conv = lambda x: x.materialize() if isinstance(x, LazySim) else x  # noqa: E731
        args = tuple(conv(a) for a in args)
        kwargs = {k: conv(v) for k, v in kwargs.items()}
        return func(*args, **kwargs)

    def __getattr__(self, name):
        # only reached for attributes LazySim itself lacks: defer to the dense tensor
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __getitem__(self, item):
        return self.materialize()[item]

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f"LazySim(shape={tuple(self.shape)}, device={self.device}, precision={self.precision!r})"
