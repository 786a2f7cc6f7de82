"""Drop-in for the hot-path part of the reference's evaluation/retrieval_evaluation.py:
`compute_recall` (:23-47) and the tail of `retrieval_evaluation` (:238-260).

The per-video forward loop (:148-236), `load_model` and the CLI are backbone / data-loading code
and out of scope.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from .. import ops
from ..model.metric import RecallAtK, _default_device, _to_device

__all__ = ["compute_recall", "compute_recall_full", "eval_tail"]


def _squeeze_text(tensor_t: torch.Tensor) -> torch.Tensor:
    t = tensor_t.squeeze()  # evaluation/retrieval_evaluation.py:31,34
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.dim() != 2:
        raise ValueError(
            "compute_recall supports one text per video ([N, 1, D] or [N, D]); multi-caption "
            f"tensors of shape {tuple(tensor_t.shape)} cannot be ranked by the reference either "
            "(faiss rejects 3-D input, SURVEY.md §3.3)")
    return t


def compute_recall_full(tensor_v: torch.Tensor, tensor_t: torch.Tensor, precision: str = "exact"
                        ) -> Dict[str, Dict[str, object]]:
    """Both retrieval directions with ranks / hits / MedR left on the device."""
    t2 = _squeeze_text(tensor_t)
    recall_range = [1, 5, 10]
    # gallery = videos, queries = texts -> text-to-video; then the roles swap
    t2v = RecallAtK("videos", "titles", recall_range, precision=precision).compute_full(tensor_v, t2)
    v2t = RecallAtK("titles", "videos", recall_range, precision=precision).compute_full(t2, tensor_v)
    return {"t2v": t2v, "v2t": v2t}


def compute_recall(tensor_v: torch.Tensor, tensor_t: torch.Tensor, split: str = "full-test",
                   dataset_name: str = "MSRVTT", precision: str = "exact") -> pd.DataFrame:
    """evaluation/retrieval_evaluation.py:23-47: R@1/5/10 (x100) in both directions as a
    DataFrame with the reference's exact index / column strings."""
    recall_range = [1, 5, 10]
    full = compute_recall_full(tensor_v, tensor_t, precision)
    n = full["t2v"]["num_samples"]
    hits = torch.stack([full["v2t"]["hits"], full["t2v"]["hits"]]).cpu().numpy()  # one D2H read
    tvr = hits[0].astype(np.float64) / n * 100.0
    vtr = hits[1].astype(np.float64) / n * 100.0
    df = pd.DataFrame(
        {
            f"{dataset_name} {split} split Video to Text": tvr,
            f"{dataset_name} {split} split Text to Video": vtr,
        },
        index=[f"R@{i}" for i in recall_range],
    )
    logging.info(df)
    return df


def _gather_rows(items: List[torch.Tensor], device: torch.device) -> torch.Tensor:
    """All rows of a list of [n_i, D] tensors as one fp32 [sum n_i, D] tensor on `device`: a single
    concatenation and (for host inputs) a single transfer instead of one per video."""
    if all(t.device == items[0].device for t in items):
        return torch.cat(items).to(device=device, dtype=torch.float32)
    return torch.cat([t.to(device=device, dtype=torch.float32) for t in items])


def _eval_tail_on(video_joint_embeddings: List[torch.Tensor],
                  caption_joint_embeddings: List[torch.Tensor],
                  device: torch.device) -> Tuple[torch.Tensor, torch.Tensor]:
    """The arithmetic of eval_tail on an explicit device (device-agnostic so that the host logic is
    testable without a GPU): O(1) launches instead of O(#videos)."""
    n = len(caption_joint_embeddings)
    D = caption_joint_embeddings[0].shape[1]
    lens_c = np.array([k.shape[0] for k in caption_joint_embeddings], dtype=np.int64)
    max_length = int(lens_c.max())
    # captions (:238-252): a buffer of -inf rows, one indexed copy of every caption row into it
    caps = torch.full((n * max_length, D), float("-inf"), dtype=torch.float32, device=device)
    if lens_c.sum() > 0:
        first = np.concatenate([[0], np.cumsum(lens_c)[:-1]])
        row = (np.arange(int(lens_c.sum())) - np.repeat(first, lens_c)
               + np.repeat(np.arange(n, dtype=np.int64) * max_length, lens_c))
        caps.index_copy_(0, torch.from_numpy(row).to(device),
                         _gather_rows([k.reshape(-1, D) for k in caption_joint_embeddings], device))
    caps = caps.view(n, max_length, D)
    # videos (:254-259): mean over each video's frame / chunk features, NOT renormalised
    lens_v = torch.tensor([k.shape[0] for k in video_joint_embeddings], dtype=torch.int64)
    vids = None
    if int(lens_v.min()) > 0:
        try:
            vids = torch.segment_reduce(_gather_rows(list(video_joint_embeddings), device), "mean",
                                        lengths=lens_v.to(device), axis=0)
        except (RuntimeError, NotImplementedError):
            vids = None  # same arithmetic, one reduction per video
    if vids is None:
        vids = torch.cat([k.to(device=device, dtype=torch.float32).mean(dim=0, keepdim=True)
                          for k in video_joint_embeddings])
    return vids, caps


def eval_tail(video_joint_embeddings: List[torch.Tensor], caption_joint_embeddings: List[torch.Tensor],
              device: Optional[torch.device] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """evaluation/retrieval_evaluation.py:238-260 on the device: -inf padding of caption lists to
    the max count, per-video mean over frame / chunk features (NOT renormalised), stack."""
    device = device or _default_device()
    return _eval_tail_on(video_joint_embeddings, caption_joint_embeddings, device)
