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
        """(scale * A) @ B.t() as a real fp32 CUDA tensor (no autograd graph)."""
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
