#!/usr/bin/env python
"""Small end-to-end pass over every kernel family, meant to run under
`compute-sanitizer --tool memcheck|racecheck|synccheck`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vtc_oracle as O  # noqa: E402
from vtc_b200 import ops  # noqa: E402
from vtc_b200.model import LazySim, PretrainedCLIP_finaltf, clip_loss  # noqa: E402
from vtc_b200.synthetic import make_batch_pair, make_cam_inputs, make_retrieval_pair  # noqa: E402

dev = torch.device("cuda:0")
T, V = make_retrieval_pair(300, 700, 64, sigma=2.0, seed=3)
q, g = T.to(dev), V.to(dev)
for prec in ("brute", "exact", "bf16"):
    r, gs = ops.sim_rank(q, g, precision=prec)
    hits, medr = ops.rank_finalize(r, gs, 700, [1, 5, 10])
    Tq, Vq = (O.bf16_round(T), O.bf16_round(V)) if prec == "bf16" else (T, V)
    assert np.array_equal(r.cpu().numpy(), O.rank0_exact(Tq, Vq)), prec
    v, i = ops.sim_topk(q, g, 11, precision=prec)
    assert np.array_equal(i.cpu().numpy(), O.topk_exact(Tq, Vq, 11)[1]), prec
# the one-call evaluation (prologue, tensor-core pass, re-check chain, hit counts, median)
for prec in ("exact", "bf16"):
    full = ops.rank_eval(q[:300], g[:300], [1, 5, 10], precision=prec)
    Tq, Vq = (O.bf16_round(T), O.bf16_round(V)) if prec == "bf16" else (T, V)
    assert np.array_equal(full["rank0"].cpu().numpy(), O.rank0_exact(Tq, Vq[:300])), prec
# every pair ties -> every 8-column group is listed with a full mask: the 16-pairs-per-step re-check
row = torch.randn(1, 64)
Qd, Gd = row.repeat(200, 1).contiguous(), row.repeat(304, 1).contiguous()
for prec in ("exact", "bf16"):
    r, _ = ops.sim_rank(Qd.to(dev), Gd.to(dev), precision=prec)
    assert np.array_equal(r.cpu().numpy(), np.arange(200)), prec
# a gallery long enough for the top-k sample pass (dense scores + per-row threshold kernel)
T2, V2 = make_retrieval_pair(200, 33000, 64, sigma=2.0, seed=4)
v, i = ops.sim_topk(T2.to(dev), V2.to(dev), 11, precision="bf16")
assert np.array_equal(i.cpu().numpy(), O.topk_exact(O.bf16_round(T2), O.bf16_round(V2), 11)[1])
# streamed (K' > 512) CTA-pair kernel
T3, V3 = make_retrieval_pair(300, 600, 768, sigma=2.0, seed=5)
r, gs = ops.sim_rank(T3.to(dev), V3.to(dev), precision="bf16")
assert np.array_equal(r.cpu().numpy(), O.rank0_exact(O.bf16_round(T3), O.bf16_round(V3)))
vis, txt = make_batch_pair(96, 64, seed=1)
a, t = vis.to(dev).requires_grad_(True), txt.to(dev).requires_grad_(True)
for force in ("", "1"):
    if force:
        os.environ["VTC_INFONCE_FORCE_TC"] = "1"
    loss = clip_loss((a, t, LazySim(a, t, torch.tensor(20.0, device=dev), "exact")), {})
    loss.backward()
    want = O.clip_loss(O.sim_matrix(vis, txt, torch.tensor(20.0))).item()
    assert abs(loss.item() - want) < 1e-4 * abs(want)
# clip_loss on a materialised sim (csrc/infonce_dense.cu)
sim = (20.0 * vis @ txt.t()).to(dev).requires_grad_(True)
clip_loss((None, None, sim), {}).backward()
m = PretrainedCLIP_finaltf(64, n_layers=2, n_heads=2).to(dev)
main, aux = make_cam_inputs(16, 3, 64, seed=2)
with torch.no_grad():
    m.eval()
    m._adapt_feature(main.to(dev), aux.to(dev))
m.train()
out = m._adapt_feature(main.to(dev).requires_grad_(True), aux.to(dev))
out.sum().backward()
torch.cuda.synchronize()
print("sanitize_small ok")
