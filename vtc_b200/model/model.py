"""Drop-in for the hot-path part of the reference's model/model.py: `normalize` (:26-27), the
Context Adapter Module `PretrainedCLIPBase._adapt_feature` (:141-205) with its transformer
(`clip.model.Transformer`, constructed at :396-398), `_encode_with_comments` (:216-266), the
averaging fusion (:356-362) and the `(feats_vis, feats_text, sim)` forward contract (:326-371,
:458-480).

The CLIP / TimeSformer backbones are out of scope: the model classes here
take precomputed features (`len(shape) == 2` branch of the reference forward, :328-330,:460-462)
or any `backbone` object exposing `encode_image` / `encode_text`.  Parameter names match the
reference (`final_transformer.resblocks.{i}.attn.in_proj_weight`, ..., `final_linear.weight`,
`mask_embedding`) so its checkpoints load and `train.py:105`'s substring grouping keeps working.

The CAM forward runs on the CUDA kernels of csrc/cam.cu + csrc/sim_tc.cu (dense projections on
the tensor cores) as one C call (vtc_cam_forward).  Under autograd (`_adapt_feature` with
parameters or inputs that require grad) it goes through `_CamAdaptFunction`, whose backward runs
on csrc/cam_bwd.cu + the same tensor-core GEMMs; gradients are checked against torch autograd
through the oracle (tests/test_gpu_parity.py::test_cam_backward_*).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import List, Optional, Sequence, Union

import torch
import torch.nn as nn

from .. import _ffi, ops
from .lazy import LazySim

__all__ = [
    "normalize",
    "CAMTransformer",
    "PretrainedCLIPBase",
    "PretrainedCLIP",
    "PretrainedCLIP_finaltf",
    "PretrainedCLIP_TimeSformer",
    "PretrainedCLIP_TimeSformer_finaltf",
    "RESIDUAL_ACTIVATIONS",
]


class _Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x.detach())
        return ops.normalize(x.detach())

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        return ops.normalize_bwd(x, dy).to(x.dtype)


class _UniformReadout(torch.autograd.Function):
    """normalize(mean_l T_l): the averaging fusion of model/model.py:356-366."""

    @staticmethod
    def forward(ctx, stacked):
        ctx.save_for_backward(stacked.detach())
        return ops.cam_readout(stacked.detach(), None, _ffi.CAM_READOUT_UNIFORM)

    @staticmethod
    def backward(ctx, dout):
        (stacked,) = ctx.saved_tensors
        return ops.cam_readout_bwd(stacked, None, dout, _ffi.CAM_READOUT_UNIFORM)[0]


def normalize(x: torch.Tensor) -> torch.Tensor:
    """x / x.norm(dim=-1, keepdim=True) -- model/model.py:26-27 (no eps; zero row -> NaN).
    Differentiable (vtc_normalize / vtc_normalize_bwd)."""
    if torch.is_grad_enabled() and x.requires_grad:
        return _Normalize.apply(x)
    return ops.normalize(x)


# residual activations (model/model.py:30-77) -> (VTC_RESACT_*, scale); "sub_mean" / "bn" use the
# BatchNorm1d running statistics (their frozen form; training-mode batch statistics are out of scope,
# no shipped config uses them)
RESIDUAL_ACTIVATIONS = {
    None: (_ffi.RESACT_NONE, 1.0), "none": (_ffi.RESACT_NONE, 1.0),
    "normalize": (_ffi.RESACT_NORMALIZE_EPS, 1.0),
    "squash": (_ffi.RESACT_SQUASH, 1.0), "squash10": (_ffi.RESACT_SQUASH, 10.0),
    "squash1p2": (_ffi.RESACT_SQUASH, 1.2), "squash1p5": (_ffi.RESACT_SQUASH, 1.5),
    "squash1p8": (_ffi.RESACT_SQUASH, 1.8),
    "tanh": (_ffi.RESACT_TANH, 1.0),
    "sub_mean": (_ffi.RESACT_AFFINE, 1.0), "bn": (_ffi.RESACT_AFFINE, 1.0),
}
NEEDS_STATE = ["sub_mean", "bn"]


class _Attn(nn.Module):
    """Parameter container with nn.MultiheadAttention's names (in_proj_weight, in_proj_bias,
    out_proj.weight, out_proj.bias) and default initialisation."""

    def __init__(self, width: int, heads: int):
        super().__init__()
        self.embed_dim = width
        self.num_heads = heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * width, width))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * width))
        self.out_proj = nn.Linear(width, width)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)


class _ResBlock(nn.Module):
    def __init__(self, width: int, heads: int):
        super().__init__()
        self.attn = _Attn(width, heads)
        self.ln_1 = nn.LayerNorm(width)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(width, width * 4)),
            ("gelu", nn.Identity()),  # QuickGELU is fused into the c_fc GEMM epilogue
            ("c_proj", nn.Linear(width * 4, width)),
        ]))
        self.ln_2 = nn.LayerNorm(width)


class CAMTransformer(nn.Module):
    """clip.model.Transformer(width, layers, heads) for the CAM (model/model.py:396-398): per block
    x += MHA(LN1(x)); x += c_proj(QuickGELU(c_fc(LN2(x)))) on sequence-first [L, b, D], no mask
    (block layout: model/timesformer_clip_alt.py:112-124)."""

    def __init__(self, width: int, layers: int, heads: int, precision: str = "exact"):
        super().__init__()
        self.width = width
        self.layers = layers
        self.heads = heads
        self.precision = precision
        self.resblocks = nn.Sequential(*[_ResBlock(width, heads) for _ in range(layers)])
        self._prepared_cache = None
        self.register_load_state_dict_post_hook(lambda module, _keys: module.invalidate_prepared())

    # ---- prepared weights: bf16 tensor-core operands + padded biases, rebuilt only when a
    # parameter changes.  torch bumps `_version` on every in-place update / optimizer step and a
    # `.to()` / `.cuda()` moves the storage, so (data_ptr, _version) catches those; writes through
    # `.data` (p.data.copy_(), EMA weight swaps) bump nothing, so load_state_dict() and train() /
    # eval() drop the cache as well, and invalidate_prepared() is there for everything else.
    def _param_key(self, extra=()):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + tuple(extra)

    def invalidate_prepared(self) -> None:
        """Forget the prepared tensor-core operands (call after writing weights through `.data`)."""
        self._prepared_cache = None

    def train(self, mode: bool = True):
        self.invalidate_prepared()
        return super().train(mode)

    def prepared(self):
        """(ctypes array of vtc_cam_layer, keep-alive list) for vtc_cam_forward."""
        key = self._param_key((self.precision,))
        cache = getattr(self, "_prepared_cache", None)
        if cache is not None and cache[0] == key:
            return cache[1], cache[2]
        keep = []
        arr = (_ffi.CamLayer * len(self.resblocks))()
        for i, blk in enumerate(self.resblocks):
            def prep(w, b):
                buf = ops.linear_prepare(w, b, self.precision)
                keep.append(buf)
                return buf.data_ptr()

            def f32(t):
                t = t.detach().float().contiguous()
                keep.append(t)
                return t.data_ptr()

            arr[i].ln1_g, arr[i].ln1_b = f32(blk.ln_1.weight), f32(blk.ln_1.bias)
            arr[i].ln2_g, arr[i].ln2_b = f32(blk.ln_2.weight), f32(blk.ln_2.bias)
            arr[i].qkv = prep(blk.attn.in_proj_weight, blk.attn.in_proj_bias)
            arr[i].out = prep(blk.attn.out_proj.weight, blk.attn.out_proj.bias)
            arr[i].fc = prep(blk.mlp.c_fc.weight, blk.mlp.c_fc.bias)
            arr[i].proj = prep(blk.mlp.c_proj.weight, blk.mlp.c_proj.bias)
        self._prepared_cache = (key, arr, keep)
        return arr, keep

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "calling the bare CAMTransformer under autograd is not supported: training goes "
                "through PretrainedCLIPBase._adapt_feature (differentiable, _CamAdaptFunction); "
                "wrap a stand-alone call in torch.no_grad()")
        L, b, D = x.shape
        prec = self.precision
        x2 = x.float().contiguous().reshape(L * b, D)
        for blk in self.resblocks:
            h = ops.layernorm(x2, blk.ln_1.weight, blk.ln_1.bias, blk.ln_1.eps)
            qkv = ops.linear(h, blk.attn.in_proj_weight, blk.attn.in_proj_bias, precision=prec)
            a = ops.cam_attn_core(qkv.reshape(L, b, 3 * D), self.heads).reshape(L * b, D)
            x2 = ops.linear(a, blk.attn.out_proj.weight, blk.attn.out_proj.bias, residual=x2,
                            precision=prec)
            h = ops.layernorm(x2, blk.ln_2.weight, blk.ln_2.bias, blk.ln_2.eps)
            f = ops.linear(h, blk.mlp.c_fc.weight, blk.mlp.c_fc.bias, act=1, precision=prec)
            x2 = ops.linear(f, blk.mlp.c_proj.weight, blk.mlp.c_proj.bias, residual=x2,
                            precision=prec)
        return x2.reshape(L, b, D)


_PARAMS_PER_LAYER = 12


def _layer_params(blk) -> list:
    return [blk.attn.in_proj_weight, blk.attn.in_proj_bias, blk.attn.out_proj.weight,
            blk.attn.out_proj.bias, blk.ln_1.weight, blk.ln_1.bias, blk.mlp.c_fc.weight,
            blk.mlp.c_fc.bias, blk.mlp.c_proj.weight, blk.mlp.c_proj.bias, blk.ln_2.weight,
            blk.ln_2.bias]


class _CamAdaptFunction(torch.autograd.Function):
    """Differentiable `_adapt_feature` (model/model.py:141-205) for training: the forward runs the
    CUDA ops one by one and keeps the activations; the backward is ONE C call (vtc_cam_backward): the
    kernels of csrc/cam_bwd.cu plus tensor-core products through the prepared transposed weights
    (dX = dY W) and on transposed operand copies (dW = dY^T X)."""

    @staticmethod
    def forward(ctx, main, aux, skip_mask, cfg, flw, *params):
        layers, heads, avg, prec, res_act = cfg
        main, aux = main.detach().float().contiguous(), aux.detach().float().contiguous()
        L, (b, D) = aux.shape[0] + 1, main.shape
        X = ops.cam_stack_normalize(main, aux).reshape(L * b, D)
        saved = []
        for i in range(layers):
            wqkv, bqkv, wo, bo, g1, b1, wfc, bfc, wpr, bpr, g2, b2 = (
                t.detach() for t in params[_PARAMS_PER_LAYER * i:_PARAMS_PER_LAYER * (i + 1)])
            H1 = ops.layernorm(X, g1, b1)
            QKV = ops.linear(H1, wqkv, bqkv, precision=prec)
            A = ops.cam_attn_core(QKV.reshape(L, b, 3 * D), heads).reshape(L * b, D)
            X2 = ops.linear(A, wo, bo, residual=X, precision=prec)
            H2 = ops.layernorm(X2, g2, b2)
            U = ops.linear(H2, wfc, bfc, precision=prec)
            Fa = ops.bias_act(U, act=1)
            Xn = ops.linear(Fa, wpr, bpr, residual=X2, precision=prec)
            saved += [X, H1, QKV, A, X2, H2, U, Fa]
            X = Xn
        T = X.reshape(L, b, D)
        res = None
        if avg:
            out = ops.cam_readout(T, main, _ffi.CAM_READOUT_AVG, skip_mask=skip_mask,
                                  res_act=res_act)
        else:
            res = ops.linear(T[0], flw.detach(), precision=prec)
            out = ops.cam_readout(None, main, _ffi.CAM_READOUT_RESIDUAL_ONLY, res_in=res,
                                  skip_mask=skip_mask, res_act=res_act)
        ctx.cfg = cfg
        ctx.skip_mask = skip_mask
        ctx.dims = (L, b, D)
        ctx.has_flw = flw is not None
        ctx.save_for_backward(main, aux, T, res if res is not None else main,
                              flw.detach() if flw is not None else main,
                              *[p.detach() for p in params], *saved)
        return out

    @staticmethod
    def backward(ctx, dout):
        layers, heads, avg, prec, res_act = ctx.cfg
        L, b, D = ctx.dims
        sv = ctx.saved_tensors
        main, aux, T, res, flw = sv[:5]
        params = sv[5:5 + _PARAMS_PER_LAYER * layers]
        acts = sv[5 + _PARAMS_PER_LAYER * layers:]
        dev = main.device
        keep = []

        def wt(w):  # the weight transposed, as a prepared gallery-side operand: dX = dY W
            buf = ops.linear_prepare(w.t().contiguous(), None, prec)
            keep.append(buf)
            return buf.data_ptr()

        grads = []
        arr = (_ffi.CamLayerBwd * layers)()
        for i in range(layers):
            wqkv, bqkv, wo, bo, g1, b1, wfc, bfc, wpr, bpr, g2, b2 = params[
                _PARAMS_PER_LAYER * i:_PARAMS_PER_LAYER * (i + 1)]
            X, H1, QKV, A, X2, H2, U, Fa = acts[8 * i:8 * (i + 1)]
            a = arr[i]
            a.X, a.H1, a.QKV, a.A = X.data_ptr(), H1.data_ptr(), QKV.data_ptr(), A.data_ptr()
            a.X2, a.H2, a.U, a.Fa = X2.data_ptr(), H2.data_ptr(), U.data_ptr(), Fa.data_ptr()
            g1c, g2c = g1.float().contiguous(), g2.float().contiguous()
            keep += [g1c, g2c]
            a.ln1_g, a.ln2_g = g1c.data_ptr(), g2c.data_ptr()
            a.qkv_t, a.out_t, a.fc_t, a.proj_t = wt(wqkv), wt(wo), wt(wfc), wt(wpr)
            gl = [torch.empty(p.shape, dtype=torch.float32, device=dev)
                  for p in (wqkv, bqkv, wo, bo, g1, b1, wfc, bfc, wpr, bpr, g2, b2)]
            (a.dWqkv, a.dbqkv, a.dWo, a.dbo, a.dg1, a.db1, a.dWfc, a.dbfc, a.dWpr, a.dbpr, a.dg2,
             a.db2) = (t.data_ptr() for t in gl)
            grads += gl
        flw_t = None
        if not avg:
            flw_t = ops.linear_prepare(flw.t().contiguous(), None, prec)
        dmain, daux, dflw = ops.cam_backward(
            dout, main, aux, T, None if avg else res, arr, heads,
            _ffi.CAM_READOUT_AVG if avg else _ffi.CAM_READOUT_RESIDUAL_ONLY, flw_t, ctx.skip_mask,
            prec, res_act, want_dflw=not avg)
        del keep
        return (dmain, daux, None, None, dflw if ctx.has_flw else None, *grads)


class PretrainedCLIPBase(nn.Module):
    """model/model.py:132-305 (hot-path methods only)."""

    # attributes the methods read; subclasses set them in __init__
    feature_dim: int
    residual_activation = None
    init_from_avg = True
    random_skip_adapter = True
    random_comment_masking = False
    branch_to_adapt = "text"
    branch_to_adapt_val = "text"
    precision = "exact"

    # embedding width of the CLIP checkpoints `clip.load` knows (model_type strings of the
    # reference's configs, e.g. configs/pretrained_clip_comments_attention.jsonc:9)
    CLIP_FEATURE_DIMS = {"ViT-B/32": 512, "ViT-B/16": 512, "ViT-L/14": 768, "ViT-L/14@336px": 768,
                         "RN50": 1024, "RN101": 512, "RN50x4": 640, "RN50x16": 768, "RN50x64": 1024}

    _final_prepared = None

    def invalidate_prepared(self) -> None:
        """Forget every prepared tensor-core operand of the CAM (final_linear's and the
        transformer's).  train() / eval() and load_state_dict() do this on their own; call it after
        writing weights through `.data`."""
        self._final_prepared = None
        tfm = getattr(self, "final_transformer", None)
        if tfm is not None and hasattr(tfm, "invalidate_prepared"):
            tfm.invalidate_prepared()

    def train(self, mode: bool = True):
        self.invalidate_prepared()
        return super().train(mode)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_prepared()
        return out

    @classmethod
    def _resolve_feature_dim(cls, model_type, feature_dim, backbone) -> int:
        """The reference reads the width off the loaded CLIP model (model/model.py:320,395); here
        the backbone is optional (precomputed features), so the width comes from, in order: the
        backbone's `ln_final`, an explicit `feature_dim`, an int passed where the reference takes
        `model_type`, or the model_type string."""
        if backbone is not None and hasattr(backbone, "ln_final"):
            return int(backbone.ln_final.normalized_shape[0])
        if feature_dim is not None:
            return int(feature_dim)
        if isinstance(model_type, int):
            return int(model_type)
        if model_type in cls.CLIP_FEATURE_DIMS:
            return cls.CLIP_FEATURE_DIMS[model_type]
        raise ValueError(f"unknown model_type {model_type!r}: pass feature_dim=")

    def _freeze(self, branch_to_freeze):
        """model/model.py:268-305.  Without a backbone there is nothing to freeze for "visual" /
        "text" / "all"; "finaltf" freezes the CAM exactly as the reference does."""
        self.branch_to_freeze = branch_to_freeze
        if branch_to_freeze is False or branch_to_freeze == "none":
            return
        did_freeze = False
        backbone = getattr(self, "model", None)
        if "visual" in branch_to_freeze:
            did_freeze = True
            if backbone is not None and hasattr(backbone, "visual"):
                for param in backbone.visual.parameters():
                    param.requires_grad = False
        if "text" in branch_to_freeze:
            did_freeze = True
            if backbone is not None and hasattr(backbone, "transformer"):
                for param in backbone.transformer.parameters():
                    param.requires_grad = False
        if "all" in branch_to_freeze:
            did_freeze = True
            if backbone is not None:
                for param in backbone.parameters():
                    param.requires_grad = False
        if "finaltf" in branch_to_freeze:
            did_freeze = True
            if hasattr(self, "final_transformer"):
                for param in self.final_transformer.parameters():
                    param.requires_grad = False
                for param in self.final_linear.parameters():
                    param.requires_grad = False
                self.mask_embedding.requires_grad = False
            else:
                import warnings

                warnings.warn("Tried to freeze finaltf but model has no final transformer, ignoring!")
        if not did_freeze:
            raise Exception("Unknown branch_to_freeze")

    def _common_init(self):
        """model/model.py:133-139: a BatchNorm1d holds the running stats of sub_mean / bn."""
        if getattr(self, "residual_activation", None) in NEEDS_STATE:
            self.mean_center_bn = nn.BatchNorm1d(self.feature_dim, affine=False, momentum=0.2)

    def _res_act(self):
        name = self.residual_activation
        if name not in RESIDUAL_ACTIVATIONS:
            raise KeyError(f"unknown residual_activation {name!r}")
        act, scale = RESIDUAL_ACTIVATIONS[name]
        if name in NEEDS_STATE:
            if self.training and "finaltf" not in (getattr(self, "branch_to_freeze", "") or ""):
                raise NotImplementedError(
                    f"residual_activation={name!r} in training mode updates batch statistics "
                    "(model/model.py:41-60); only the eval-mode form is built")
            bn = self.mean_center_bn
            mul = None if name == "sub_mean" else torch.rsqrt(bn.running_var + bn.eps)
            return act, scale, bn.running_mean, mul
        return act, scale, None, None

    def _adapt_feature(self, feature_main: torch.Tensor, features_aux) -> torch.Tensor:
        """CAM: model/model.py:141-205.  feature_main [b, D]; features_aux [nc, b, D] tensor or a
        list of nc [b, D] tensors.  Returns the adapted, unit-norm feature [b, D]."""
        res_act = self._res_act()
        assert len(feature_main.shape) == 2
        b = feature_main.shape[0]
        if not isinstance(features_aux, torch.Tensor):
            features_aux = torch.stack(list(features_aux), dim=0)
        assert features_aux.shape[1] == b
        tfm = self.final_transformer
        # the reference draws one number from the global CPU RNG here for a 5 % debug print
        # (:163); keep the draw so RNG streams stay aligned with it, drop the print.
        torch.rand([])
        skip_mask = None
        if self.training and self.random_skip_adapter:
            skip_mask = torch.rand(b) > 0.5                                      # :199-201
        needs_grad = torch.is_grad_enabled() and (
            feature_main.requires_grad or features_aux.requires_grad
            or any(p.requires_grad for p in tfm.parameters())
            or (not self.init_from_avg and self.final_linear.weight.requires_grad))
        if needs_grad:
            params = [p for blk in tfm.resblocks for p in _layer_params(blk)]
            act = None if res_act[0] == _ffi.RESACT_NONE else tuple(
                t.detach() if isinstance(t, torch.Tensor) else t for t in res_act)
            cfg = (len(tfm.resblocks), tfm.heads, bool(self.init_from_avg), self.precision, act)
            sm = None if skip_mask is None else skip_mask.to(feature_main.device)
            flw = None if self.init_from_avg else self.final_linear.weight
            return _CamAdaptFunction.apply(feature_main, features_aux, sm, cfg, flw, *params)
        layers, _keep = tfm.prepared()
        final = None
        if not self.init_from_avg:                                               # :161
            w = self.final_linear.weight
            key = (w.data_ptr(), w._version, self.precision)
            cache = getattr(self, "_final_prepared", None)
            if cache is None or cache[0] != key:
                cache = (key, ops.linear_prepare(w, None, self.precision))
                self._final_prepared = cache
            final = cache[1]
        # stack + normalize (:150-151) -> transformer (:155) -> read-out (:156-161) -> skip
        # (:199-201) -> normalize(normalize(main) + res) (:203): 2 + 7 * layers launches
        return ops.cam_forward(feature_main, features_aux, layers, tfm.heads,
                               _ffi.CAM_READOUT_AVG if self.init_from_avg
                               else _ffi.CAM_READOUT_RESIDUAL_ONLY,
                               final_linear=final, skip_mask=skip_mask, precision=self.precision,
                               res_act=res_act)

    def _comment_embeddings(self, comments):
        """-> (comment embeddings [b, nc, D] fp32, empty-string mask [b, nc] or None).  `comments`
        is either precomputed comment embeddings [b, nc, D] (optionally a tuple (embeddings,
        empty_mask)) or token ids [b, nc, ntoks] when a backbone with `encode_text` is attached."""
        empty_mask = None
        if isinstance(comments, (tuple, list)):
            comments, empty_mask = comments
        if comments.dtype in (torch.int32, torch.int64):
            if getattr(self, "model", None) is None:
                raise ops.VtcError("token-id comments need a backbone with encode_text")
            empty_mask = comments[..., 1] == 49407                               # :208
            b, ncomms, ntoks = comments.shape
            feats_comm = self.model.encode_text(comments.reshape(b * ncomms, ntoks))
            feats_comm = feats_comm.reshape(b, ncomms, self.feature_dim).float()
        else:
            feats_comm = comments.float().clone()
        return feats_comm, empty_mask

    def _load_comment_features(self, comments) -> torch.Tensor:
        """model/model.py:207-214: comment embeddings with `mask_embedding` in place of empty
        comments, sequence-first [nc, b, D]."""
        feats_comm, empty_mask = self._comment_embeddings(comments)
        if empty_mask is not None:
            # :212 -- an autograd-visible assignment: mask_embedding is a trained parameter
            feats_comm[empty_mask] = self.mask_embedding.to(feats_comm.dtype)
        return feats_comm.permute(1, 0, 2)                                       # :213

    def _encode_with_comments(self, feats_vis, feats_title, comments):
        """model/model.py:216-266 (audio branch out of scope)."""
        feats_comm = self._load_comment_features(comments)
        bs = feats_title.shape[0]
        if self.training:
            if self.random_comment_masking:
                comm_masks = [torch.randint(low=0, high=2, size=(bs, 1), device=comm.device)
                              for comm in feats_comm]                             # :237-240
            else:
                comm_masks = torch.ones(len(feats_comm), device=feats_comm.device)  # :241-242
            feats_comm = [comm * mask + self.mask_embedding * (1 - mask)
                          for comm, mask in zip(feats_comm, comm_masks)]          # :243-246
            branch_to_adapt = self.branch_to_adapt
        else:
            branch_to_adapt = self.branch_to_adapt_val

        if branch_to_adapt == "text":
            feats_vis_out = feats_vis
            feats_text_out = self._adapt_feature(feats_title, feats_comm)
        elif branch_to_adapt == "image":
            feats_vis_out = self._adapt_feature(feats_vis, feats_comm)
            feats_text_out = feats_title
        elif branch_to_adapt == "skip":
            feats_vis_out = feats_vis
            feats_text_out = feats_title
        else:
            raise Exception("Unknown branch_to_adapt")

        return normalize(feats_vis_out.float()), normalize(feats_text_out.float())  # :263-264

    # ---- shared forward helpers -------------------------------------------------------------
    def _features(self, vis, title):
        shp = vis.shape
        if len(shp) == 2 and shp[1] == self.feature_dim:
            feats_vis = vis                                                      # precomputed, :328-330
        else:
            if getattr(self, "model", None) is None:
                raise ops.VtcError("raw frames need a backbone with encode_image")
            if len(shp) == 4:
                feats_vis = self.model.encode_image(vis).float()
            else:                                                                # :335-340
                feats_vis = self.model.encode_image(
                    vis.reshape(shp[0] * shp[1], shp[2], shp[3], shp[4])).float()
                feats_vis = feats_vis.reshape(shp[0], shp[1], -1).mean(1)
        if title.dim() == 2 and title.dtype.is_floating_point and title.shape[1] == self.feature_dim:
            feats_title = title
        else:
            if getattr(self, "model", None) is None:
                raise ops.VtcError("token-id titles need a backbone with encode_text")
            feats_title = self.model.encode_text(title)
        return feats_vis, feats_title

    def _logit_scale_exp(self, device):
        if getattr(self, "model", None) is not None and hasattr(self.model, "logit_scale"):
            return self.model.logit_scale.exp().to(device)
        return self.logit_scale.exp().to(device)


class PretrainedCLIP(PretrainedCLIPBase):
    """model/model.py:308-371 with the backbone made pluggable."""

    def __init__(self, model_type="ViT-B/32", freeze=False, residual_activation=None,
                 comment_fusion=None, *, feature_dim: Optional[int] = None, backbone=None,
                 logit_scale_init: float = math.log(1 / 0.07), precision: str = "exact",
                 lazy_sim: bool = True):
        """Positional / keyword arguments as the reference's (model/model.py:309-324), so that
        `config.init_obj("arch", module_arch)` works with its configs; `model_type` may also be an
        int = the embedding width.  Keyword-only extras: `backbone` (an object with encode_image /
        encode_text / logit_scale; None = precomputed features in), `feature_dim`, `precision`."""
        super().__init__()
        self.model = backbone
        self.feature_dim = self._resolve_feature_dim(model_type, feature_dim, backbone)
        self.residual_activation = residual_activation
        self.comment_fusion = comment_fusion
        self.precision = precision
        self.lazy_sim = lazy_sim
        self._common_init()
        if backbone is None or not hasattr(backbone, "logit_scale"):
            self.logit_scale = nn.Parameter(torch.ones([]) * logit_scale_init)
        self._freeze(freeze)

    def forward(self, vis, title, comments=None):
        feats_vis, feats_title = self._features(vis, title)
        if comments is None or self.comment_fusion is None or self.comment_fusion == "None":
            feats_text = normalize(feats_title.float())
        elif self.comment_fusion == "averaging":
            # :346-351 -- plain comment embeddings: this class has no mask_embedding, empty
            # comments are averaged in as the encoder sees them
            feats_comm = self._comment_embeddings(comments)[0].permute(1, 0, 2)  # [nc, b, D]
            stacked = torch.cat([feats_title.float().unsqueeze(0), feats_comm], 0)
            feats_text = _UniformReadout.apply(stacked)                          # :356-362,:366
        else:
            raise ValueError("Comment fusion method not specified.")
        feats_vis = normalize(feats_vis.float())                                 # :367
        sim = LazySim(feats_vis, feats_text, self._logit_scale_exp(feats_vis.device), self.precision)
        return feats_vis, feats_text, (sim if self.lazy_sim else sim.materialize())


class PretrainedCLIP_finaltf(PretrainedCLIPBase):
    """model/model.py:374-480 with the backbone made pluggable (audio branch out of scope)."""

    def __init__(self, model_type="ViT-B/32", freeze=False, branch_to_adapt="text",
                 branch_to_adapt_val="text", residual_activation=None, n_layers=2, n_heads=8,
                 init_from_avg=True, random_comment_masking=False, random_skip_adapter=True,
                 init_audio_model=False, audio_model_ckpt=None, clip_audio_ckpt=None, *,
                 feature_dim: Optional[int] = None, backbone=None,
                 logit_scale_init: float = math.log(1 / 0.07), precision: str = "exact",
                 lazy_sim: bool = True):
        """Positional / keyword arguments as the reference's (model/model.py:375-390), so that its
        configs construct this class unchanged (configs/pretrained_clip_comments_attention.jsonc:
        7-17); `model_type` may also be an int = the embedding width.  Keyword-only extras as in
        PretrainedCLIP."""
        super().__init__()
        if init_audio_model or audio_model_ckpt or clip_audio_ckpt:
            raise NotImplementedError("the audio branch (model/model.py:405-437) is out of scope")
        self.model = backbone
        feature_dim = self._resolve_feature_dim(model_type, feature_dim, backbone)
        self.feature_dim = feature_dim
        self.final_transformer = CAMTransformer(feature_dim, int(n_layers), int(n_heads), precision)
        self.final_linear = nn.Linear(feature_dim, feature_dim, bias=False)
        self.mask_embedding = nn.Parameter(torch.randn(1, feature_dim))
        self.branch_to_adapt = branch_to_adapt
        self.branch_to_adapt_val = branch_to_adapt_val
        self.residual_activation = residual_activation
        self.init_from_avg = init_from_avg
        self.random_comment_masking = random_comment_masking
        self.random_skip_adapter = random_skip_adapter
        self.precision = precision
        self.lazy_sim = lazy_sim
        self._common_init()
        if backbone is None or not hasattr(backbone, "logit_scale"):
            self.logit_scale = nn.Parameter(torch.ones([]) * logit_scale_init)
        if self.init_from_avg:                                                   # :440-450
            for blk in self.final_transformer.resblocks:
                blk.mlp.c_proj.weight.data.zero_()
                blk.mlp.c_proj.bias.data.zero_()
                blk.attn.out_proj.weight.data.zero_()
        nn.init.constant_(self.final_linear.weight, 0.0)                         # :452
        self._freeze(freeze)                                                     # :454

    def forward(self, vis, title, comments):
        feats_vis, feats_title = self._features(vis, title)
        feats_vis, feats_text = self._encode_with_comments(feats_vis.float(), feats_title.float(),
                                                           comments)
        sim = LazySim(feats_vis, feats_text, self._logit_scale_exp(feats_vis.device), self.precision)
        return feats_vis, feats_text, (sim if self.lazy_sim else sim.materialize())


class _TimeSformerFeatures:
    """Feature extraction of the reference's TimeSformer variants (model/model.py:494-498,
    598-613): the visual tower is `self.model.visual` applied to the whole [b, t, c, h, w] clip
    (not per-frame `encode_image`), optionally living on its own device.  The tower itself
    (model/timesformer_clip_alt.py) is a backbone and out of scope: it comes in as `backbone`."""

    def _features(self, vis, title):
        shp = vis.shape
        if len(shp) == 2 and shp[1] == self.feature_dim:
            feats_vis = vis                                                      # precomputed
        else:
            if getattr(self, "model", None) is None:
                raise ops.VtcError("raw clips need a backbone with a `visual` tower")
            vdev = getattr(self, "visual_device", None)
            if vdev is not None:                                                 # :598-609
                if getattr(self, "text_device", None) is None:
                    self.text_device = title.device
                    self.model.visual.to(vdev)
                feats_vis = self.model.visual(vis.to(vdev)).to(self.text_device)
            else:
                feats_vis = self.model.visual(vis)
        if title.dim() == 2 and title.dtype.is_floating_point and title.shape[1] == self.feature_dim:
            feats_title = title
        else:
            if getattr(self, "model", None) is None:
                raise ops.VtcError("token-id titles need a backbone with encode_text")
            feats_title = self.model.encode_text(title)
        return feats_vis, feats_title


class PretrainedCLIP_TimeSformer(_TimeSformerFeatures, PretrainedCLIP):
    """model/model.py:483-507: normalise both towers' features, similarity -- the hot path of
    PretrainedCLIP without comment fusion."""

    def __init__(self, model_type="ViT-B/32", freeze=False, residual_activation=None, *,
                 feature_dim: Optional[int] = None, backbone=None,
                 logit_scale_init: float = math.log(1 / 0.07), precision: str = "exact",
                 lazy_sim: bool = True):
        super().__init__(model_type, freeze, residual_activation, None, feature_dim=feature_dim,
                         backbone=backbone, logit_scale_init=logit_scale_init, precision=precision,
                         lazy_sim=lazy_sim)

    def forward(self, im, text, comments=None):
        return super().forward(im, text, None)           # comments are ignored (:494-507)


class PretrainedCLIP_TimeSformer_finaltf(_TimeSformerFeatures, PretrainedCLIP_finaltf):
    """model/model.py:537-621: the CAM over a TimeSformer visual tower; `visual_device` places
    that tower on a second GPU as the reference does (:588-609)."""

    def __init__(self, model_type="ViT-B/32", freeze=False, branch_to_adapt="text",
                 branch_to_adapt_val="text", residual_activation=None, visual_device=None,
                 n_layers=2, n_heads=8, init_from_avg=True, random_comment_masking=False,
                 random_skip_adapter=True, *, feature_dim: Optional[int] = None, backbone=None,
                 logit_scale_init: float = math.log(1 / 0.07), precision: str = "exact",
                 lazy_sim: bool = True):
        super().__init__(model_type, freeze, branch_to_adapt, branch_to_adapt_val,
                         residual_activation, n_layers, n_heads, init_from_avg,
                         random_comment_masking, random_skip_adapter, feature_dim=feature_dim,
                         backbone=backbone, logit_scale_init=logit_scale_init, precision=precision,
                         lazy_sim=lazy_sim)
        self.multigpu = visual_device is not None
        if self.multigpu:
            self.visual_device = torch.device(visual_device)
            self.text_device = None
