"""Second consumer of RecallAtK: the VTC test-split evaluation (evaluation/eval.py:97-138).

The reference runs the model over a loader, moves every batch of features to the host, stacks
them with numpy and calls `RecallAtK.compute` both ways, producing six floats keyed
`R{1,5,10}_{title_from_im,im_from_title}`.  Here the features stay on the device between the
forward passes and the two ranking calls; the keys, the gallery/query roles
and the returned Python floats are the reference's.  Config / checkpoint / dataset plumbing
(evaluation/eval.py:30-95) is out of scope.
"""
from __future__ import annotations

import json
from typing import Callable, Dict, Iterable, Optional

import torch

from ..model.metric import RecallAtK

K_VALS = [1, 5, 10]


def recall_summary(res_vis, res_text, precision: str = "exact") -> Dict[str, float]:
    """evaluation/eval.py:118-135: title_from_im = compute(gallery=vis, queries=text) with
    RecallAtK("images", "titles"); im_from_title = compute(gallery=text, queries=vis)."""
    title_from_im = RecallAtK("images", "titles", K_VALS, precision=precision).compute(res_vis, res_text)
    im_from_title = RecallAtK("titles", "images", K_VALS, precision=precision).compute(res_text, res_vis)
    out = {}
    for (k, r) in title_from_im:
        out[f"R{k}_title_from_im"] = r
    for (k, r) in im_from_title:
        out[f"R{k}_im_from_title"] = r
    return out


def evaluate(model: Callable, data_loader: Iterable, device: torch.device,
             save_path: Optional[str] = None, precision: str = "exact",
             transform_comments: Optional[Callable] = None) -> Dict[str, float]:
    """evaluation/eval.py:97-138.  `data_loader` yields (vis, title, comments, meta);
    `model(vis, title, comments)` returns (feats_vis, feats_text, sim).  `transform_comments`
    stands in for the reference's add_irrelevant_comms hook (`:104-108`)."""
    res_vis, res_text = [], []
    for vis, title, comments, _meta in data_loader:
        with torch.no_grad():
            if transform_comments is not None:
                comments = transform_comments(comments)
            feats_vis, feats_text, _ = model(torch.squeeze(vis).to(device),
                                             torch.squeeze(title).to(device), comments.to(device))
        res_vis.append(feats_vis.detach())      # stays on the device (the reference: .cpu().numpy())
        res_text.append(feats_text.detach())
    out = recall_summary(torch.cat(res_vis), torch.cat(res_text), precision)
    if save_path:
        with open(save_path, "w") as f:
            json.dump(out, f)
    return out
