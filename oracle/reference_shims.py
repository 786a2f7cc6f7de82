"""Run the REFERENCE'S OWN function bodies in this container.  TEST INFRASTRUCTURE ONLY.

/root/reference is pure Python but two of its imports are absent here (``clip`` -- OpenAI CLIP,
un-pinned git dependency, environment.yml:30; ``faiss`` -- un-pinned conda faiss-gpu,
environment.yml:12) and a few lines assume Python < 3.10.  This module installs in-memory
``sys.modules`` stand-ins (nothing is written to disk, nothing is copied from the reference) so
that these reference symbols import and execute unmodified:

    model/loss.py::clip_loss
    model/metric.py::RecallAtK (.compute / .update / .result)
    model/model.py::PretrainedCLIPBase._adapt_feature / ._encode_with_comments / normalize
    evaluation/retrieval_evaluation.py::compute_recall

It is used by ``tests/golden/generate_golden.py`` (to produce the committed fixtures) and by
``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is absent, e.g. on the GPU
box).  The recipe follows SURVEY.md Appendix A.
"""
from __future__ import annotations

import collections
import collections.abc
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("VTC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "loss.py"))


# ----------------------------------------------------------------------------- clip stand-in
class _LayerNorm(nn.LayerNorm):
    def forward(self, x):
        return super().forward(x.float()).to(x.dtype)


class _QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class _ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head, attn_mask=None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = _LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", _QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = _LayerNorm(d_model)
        self.attn_mask = attn_mask

    def forward(self, x):
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False, attn_mask=self.attn_mask)[0]
        return x + self.mlp(self.ln_2(x))


class _Transformer(nn.Module):
    """Stand-in for clip.model.Transformer(width, layers, heads, attn_mask=None); attribute names
    (`resblocks[i].attn/.ln_1/.mlp.c_fc/.mlp.c_proj/.ln_2`) are the ones model/model.py:440-450
    and checkpoints address."""

    def __init__(self, width, layers, heads, attn_mask=None):
        super().__init__()
        self.width = width
        self.layers = layers
        self.resblocks = nn.Sequential(
            *[_ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


# ----------------------------------------------------------------------------- faiss stand-in
class _GpuIndexFlatConfig:
    useFloat16 = False
    device = 0


class _StandardGpuResources:
    pass


def _make_faiss_module():
    from oracle.vtc_oracle import FlatL2Index

    class GpuIndexFlatL2(FlatL2Index):
        def __init__(self, res, d, cfg=None):
            super().__init__(d)

    m = types.ModuleType("faiss")
    m.GpuIndexFlatConfig = _GpuIndexFlatConfig
    m.StandardGpuResources = _StandardGpuResources
    m.GpuIndexFlatL2 = GpuIndexFlatL2
    return m


_installed = False


def install() -> None:
    """Install the stand-ins and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    # model/metric.py:106 uses collections.Iterable (removed in Python 3.10)
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable

    clip = types.ModuleType("clip")
    clip_model = types.ModuleType("clip.model")
    clip_model.Transformer = _Transformer
    clip_model.LayerNorm = _LayerNorm
    clip_model.QuickGELU = _QuickGELU
    clip.model = clip_model

    def _no_clip(*a, **k):
        raise RuntimeError("clip.load/tokenize are out of scope for the oracle shim")

    clip.load = _no_clip
    clip.tokenize = _no_clip
    sys.modules.setdefault("clip", clip)
    sys.modules.setdefault("clip.model", clip_model)
    sys.modules.setdefault("faiss", _make_faiss_module())

    # evaluation/retrieval_evaluation.py:16 imports the dataset loaders (rake_nltk, ffmpeg, ...)
    # at module level; compute_recall never touches them.
    dl = types.ModuleType("dataset_loaders")
    dl.__path__ = []
    dl_dl = types.ModuleType("dataset_loaders.dataset_loaders")
    dl.dataset_loaders = dl_dl
    sys.modules.setdefault("dataset_loaders", dl)
    sys.modules.setdefault("dataset_loaders.dataset_loaders", dl_dl)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def ref_loss_module():
    install()
    import importlib

    return importlib.import_module("model.loss")


def ref_metric_module():
    install()
    import importlib

    return importlib.import_module("model.metric")


def ref_model_module():
    install()
    import importlib

    return importlib.import_module("model.model")


def ref_retrieval_evaluation_module():
    install()
    import importlib

    # evaluation/retrieval_evaluation.py:110 annotates `model: torch.nn.module` (sic), which old
    # torch tolerated; give the name a meaning so the module imports unmodified.
    if not hasattr(torch.nn, "module"):
        torch.nn.module = torch.nn.Module

    return importlib.import_module("evaluation.retrieval_evaluation")


def make_ref_cam(width: int, layers: int, heads: int, params=None, init_from_avg=True,
                 residual_activation=None, final_linear_weight=None, mask_embedding=None,
                 branch_to_adapt_val="text"):
    """A light subclass of the reference's PretrainedCLIPBase carrying only what
    `_adapt_feature` / `_encode_with_comments` read."""
    mm = ref_model_module()

    class _Cam(mm.PretrainedCLIPBase):
        def __init__(self):
            nn.Module.__init__(self)
            self.feature_dim = width
            self.final_transformer = _Transformer(width, layers, heads)
            self.final_linear = nn.Linear(width, width, bias=False)
            self.mask_embedding = nn.Parameter(torch.randn(1, width))
            self.init_from_avg = init_from_avg
            self.residual_activation = residual_activation
            self.random_skip_adapter = True
            self.random_comment_masking = False
            self.branch_to_adapt = "text"
            self.branch_to_adapt_val = branch_to_adapt_val
            self.branch_to_freeze = ""
            self.init_audio_model = False

    cam = _Cam()
    if params is not None:
        missing, unexpected = cam.final_transformer.load_state_dict(params, strict=True)
    if final_linear_weight is not None:
        with torch.no_grad():
            cam.final_linear.weight.copy_(final_linear_weight)
    if mask_embedding is not None:
        with torch.no_grad():
            cam.mask_embedding.copy_(mask_embedding)
    cam.eval()
    return cam
