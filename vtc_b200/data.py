"""Cached-embedding files: the reference's on-disk input for the retrieval hot path.

`scripts/get_clip_vit_embeddings.py:72-78` writes `{"reddit_ids": int64[N], "embeddings":
float32[N, D]}` with torch.save, and `dataset_loaders/dataset_loaders.py:162-184::load_features`
reads it back by reddit id (asserting those dtypes).  These helpers load such files straight into
pinned host memory (so the H2D staging of `RecallAtK.compute` can overlap the ranking) and align two
files by id so that ground truth is "same row" as `model/metric.py:149-160` assumes.
SURVEY.md §8f row 3.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

__all__ = ["load_cached_embeddings", "align_by_id", "recall_from_cached"]


def load_cached_embeddings(path: str, pin: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (ids int64 [N], embeddings float32 [N, D]); dtype checks follow load_features :176-177."""
    stored = torch.load(path, map_location="cpu")
    if "reddit_id_to_comment_id" in stored:
        raise ValueError("comment-embedding files (lists of per-video tensors, "
                         "dataset_loaders.py:165-174) are not a flat gallery; load them with the "
                         "dataset loader and pass the tensors to the model instead")
    ids, emb = stored["reddit_ids"], stored["embeddings"]
    assert ids.dtype is torch.int64, "reddit_ids must be int64"
    assert emb.dtype is torch.float32, "embeddings must be float32"
    assert emb.dim() == 2 and ids.shape == (emb.shape[0],)
    emb = emb.contiguous()
    if pin and torch.cuda.is_available():
        emb = emb.pin_memory()
    return ids, emb


def align_by_id(ids_a: torch.Tensor, emb_a: torch.Tensor, ids_b: torch.Tensor, emb_b: torch.Tensor,
                order: Optional[Sequence[int]] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Rows of both files restricted to the ids they share, in the same order (that of `order`
    when given -- e.g. `df.reddit_id` -- else of file a).  -> (ids, emb_a', emb_b')."""
    look_b: Dict[int, int] = {int(v): i for i, v in enumerate(ids_b.tolist())}
    look_a: Dict[int, int] = {int(v): i for i, v in enumerate(ids_a.tolist())}
    wanted = [int(v) for v in (order if order is not None else ids_a.tolist())]
    keep = [v for v in wanted if v in look_a and v in look_b]
    if order is not None and len(keep) != len(wanted):
        missing = [v for v in wanted if v not in look_a or v not in look_b][:5]
        raise KeyError(f"{len(wanted) - len(keep)} requested ids are missing, e.g. {missing}")
    sa = torch.tensor([look_a[v] for v in keep], dtype=torch.int64)
    sb = torch.tensor([look_b[v] for v in keep], dtype=torch.int64)
    return torch.tensor(keep, dtype=torch.int64), emb_a[sa].contiguous(), emb_b[sb].contiguous()


def recall_from_cached(path_video: str, path_text: str, split: str = "full-test",
                       dataset_name: str = "cached", precision: str = "exact",
                       normalise: bool = True):
    """R@1/5/10 both ways (the reference's compute_recall DataFrame) from two cached files."""
    from . import ops
    from .evaluation.retrieval_evaluation import compute_recall

    ids_v, v = load_cached_embeddings(path_video)
    ids_t, t = load_cached_embeddings(path_text)
    _, v, t = align_by_id(ids_v, v, ids_t, t)
    if normalise:  # raw CLIP features: normalise like the model forward does (model.py:366-367)
        dev = torch.device("cuda", torch.cuda.current_device())
        v = ops.normalize(v.to(dev))
        t = ops.normalize(t.to(dev))
    return compute_recall(v, t.unsqueeze(1), split=split, dataset_name=dataset_name,
                          precision=precision)
