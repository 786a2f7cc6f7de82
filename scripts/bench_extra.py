#!/usr/bin/env python
"""Secondary measurements (BASELINE configs 2, 3, 5 and the same-box stock-torch bars).

Prints one JSON line per measurement; results are summarised under profiles/.  Stock torch
(cuBLAS `A @ B.T` + compare/topk, `F.cross_entropy`, eager CAM) is the "existing Blackwell kernel"
bar of SURVEY.md §2c / BASELINE.md §2, timed on the same GPU in the same process.
"""
from __future__ import annotations

import json
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vtc_b200 import _ffi, ops  # noqa: E402
from vtc_b200.model import LazySim, PretrainedCLIP_finaltf, clip_loss  # noqa: E402
from vtc_b200.synthetic import make_batch_pair, make_cam_inputs, make_retrieval_pair  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _ffi.launch_count()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3, (_ffi.launch_count() - l0) / iters  # us, launches


def graphed(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


def emit(**kw):
    print(json.dumps(kw), flush=True)


def config2():
    b, D, nc = 256, 512, 5
    vis, txt = make_batch_pair(b, D, seed=1023)
    a, t = vis.to(dev), txt.to(dev)
    scale = torch.tensor(100.0, device=dev)
    for prec in ("exact", "bf16"):
        us, nl = timeit(lambda: clip_loss((a, t, LazySim(a, t, scale, prec)), {}))
        emit(bench="c2_infonce_fwd", precision=prec, us=us, launches=nl)
        try:
            rep = graphed(lambda: ops.infonce_fwd(a, t, scale, prec))
            us_g, _ = timeit(rep)
            emit(bench="c2_infonce_fwd_cudagraph", precision=prec, us=us_g)
        except Exception as e:  # noqa: BLE001
            emit(bench="c2_infonce_fwd_cudagraph", precision=prec, error=str(e)[:200])
    ar, tr = a.clone().requires_grad_(True), t.clone().requires_grad_(True)

    def fwd_bwd():
        loss = clip_loss((ar, tr, LazySim(ar, tr, scale, "exact")), {})
        loss.backward()
        ar.grad = None
        tr.grad = None

    us, nl = timeit(fwd_bwd)
    emit(bench="c2_infonce_fwd_bwd", precision="exact", us=us, launches=nl)

    def torch_loss():
        sim = scale * a @ t.t()
        labels = torch.arange(b, device=dev)
        return 0.5 * (F.cross_entropy(sim, labels) + F.cross_entropy(sim.t(), labels))

    us, _ = timeit(torch_loss)
    emit(bench="c2_infonce_fwd_torch_eager_fp32", us=us)
    us, _ = timeit(graphed(torch_loss))
    emit(bench="c2_infonce_fwd_torch_cudagraph_fp32", us=us)

    # CAM
    main, aux = make_cam_inputs(b, nc, D, seed=1023)
    m, x = main.to(dev), aux.to(dev)
    for prec in ("exact", "bf16"):
        cam = PretrainedCLIP_finaltf(D, precision=prec).to(dev).eval()
        for blk in cam.final_transformer.resblocks:  # make the transformer non-trivial
            torch.nn.init.normal_(blk.mlp.c_proj.weight, std=0.02)
            torch.nn.init.normal_(blk.attn.out_proj.weight, std=0.02)
        with torch.no_grad():
            us, nl = timeit(lambda: cam._adapt_feature(m, x), iters=30)
            emit(bench="c2_cam_adapt_feature", precision=prec, us=us, launches=nl)
            try:
                us_g, _ = timeit(graphed(lambda: cam._adapt_feature(m, x)), iters=30)
                emit(bench="c2_cam_adapt_feature_cudagraph", precision=prec, us=us_g)
            except Exception as e:  # noqa: BLE001
                emit(bench="c2_cam_adapt_feature_cudagraph", precision=prec, error=str(e)[:200])

    # stock torch CAM (nn.MultiheadAttention blocks), eager fp32 and bf16 autocast
    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.attn = torch.nn.MultiheadAttention(D, 8)
            self.ln_1 = torch.nn.LayerNorm(D)
            self.c_fc = torch.nn.Linear(D, 4 * D)
            self.c_proj = torch.nn.Linear(4 * D, D)
            self.ln_2 = torch.nn.LayerNorm(D)

        def forward(self, h):
            y = self.ln_1(h)
            h = h + self.attn(y, y, y, need_weights=False)[0]
            y = self.c_fc(self.ln_2(h))
            return h + self.c_proj(y * torch.sigmoid(1.702 * y))

    blocks = torch.nn.Sequential(Block(), Block()).to(dev).eval()

    def nrm(v):
        return v / v.norm(dim=-1, keepdim=True)

    def torch_cam():
        with torch.no_grad():
            c = nrm(torch.stack([m, *x], 0))
            tf = blocks(c)
            res = nrm(torch.mean(torch.stack([nrm(s) for s in tf], 0), 0))
            return nrm(nrm(m) + res)

    us, _ = timeit(torch_cam, iters=30)
    emit(bench="c2_cam_torch_eager_fp32", us=us)
    try:
        us, _ = timeit(graphed(torch_cam), iters=30)
        emit(bench="c2_cam_torch_cudagraph_fp32", us=us)
    except Exception as e:  # noqa: BLE001
        emit(bench="c2_cam_torch_cudagraph_fp32", error=str(e)[:200])


def config2_train():
    """One training step of the hot path at config-2 size: CAM forward + InfoNCE + backward through
    both (precomputed CLIP features in, as in the reference's cached-feature branch)."""
    b, D, nc = 256, 512, 5
    vis, txt = make_batch_pair(b, D, seed=1023)
    main, aux = make_cam_inputs(b, nc, D, seed=1023)
    v, x = vis.to(dev), aux.to(dev).permute(1, 0, 2).contiguous()  # comments [b, nc, D]
    for prec in ("exact", "bf16"):
        model = PretrainedCLIP_finaltf(D, precision=prec).to(dev).train()
        model.random_skip_adapter = False
        for blk in model.final_transformer.resblocks:
            torch.nn.init.normal_(blk.mlp.c_proj.weight, std=0.02)
            torch.nn.init.normal_(blk.attn.out_proj.weight, std=0.02)
        title = txt.to(dev).requires_grad_(True)

        def step():
            out = model(v, title, x)
            loss = clip_loss(out, {})
            loss.backward()
            model.zero_grad(set_to_none=True)
            title.grad = None

        us, nl = timeit(step, iters=20)
        emit(bench="c2_train_step_cam+infonce_fwd_bwd", precision=prec, us=us, launches=nl)
        # the library never synchronises, so the whole step (forward, loss, backward) can be
        # captured and replayed as one CUDA graph: GPU time without the Python launch overhead
        try:
            us, _ = timeit(graphed(step), iters=50)
            emit(bench="c2_train_step_cam+infonce_fwd_bwd_cudagraph", precision=prec, us=us)
        except Exception as e:  # noqa: BLE001
            emit(bench="c2_train_step_cam+infonce_fwd_bwd_cudagraph", precision=prec, error=str(e)[:200])

    # stock torch: same computation, eager fp32
    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.attn = torch.nn.MultiheadAttention(D, 8)
            self.ln_1 = torch.nn.LayerNorm(D)
            self.c_fc = torch.nn.Linear(D, 4 * D)
            self.c_proj = torch.nn.Linear(4 * D, D)
            self.ln_2 = torch.nn.LayerNorm(D)

        def forward(self, h):
            y = self.ln_1(h)
            h = h + self.attn(y, y, y, need_weights=False)[0]
            y = self.c_fc(self.ln_2(h))
            return h + self.c_proj(y * torch.sigmoid(1.702 * y))

    blocks = torch.nn.Sequential(Block(), Block()).to(dev).train()
    scale = torch.nn.Parameter(torch.tensor(math.log(1 / 0.07), device=dev))
    title = txt.to(dev).requires_grad_(True)
    comm = x.permute(1, 0, 2).contiguous()

    def nrm(t):
        return t / t.norm(dim=-1, keepdim=True)

    def torch_step():
        c = nrm(torch.stack([title, *comm], 0))
        tf = blocks(c)
        res = nrm(torch.mean(torch.stack([nrm(s_) for s_ in tf], 0), 0))
        ft = nrm(nrm(nrm(title) + res))
        fv = nrm(v)
        sim = scale.exp() * fv @ ft.t()
        labels = torch.arange(b, device=dev)
        loss = 0.5 * (F.cross_entropy(sim, labels) + F.cross_entropy(sim.t(), labels))
        loss.backward()
        blocks.zero_grad(set_to_none=True)
        title.grad = None
        scale.grad = None

    us, _ = timeit(torch_step, iters=20)
    emit(bench="c2_train_step_torch_eager_fp32", us=us)
    try:
        us, _ = timeit(graphed(torch_step), iters=50)
        emit(bench="c2_train_step_torch_cudagraph_fp32", us=us)
    except Exception as e:  # noqa: BLE001
        emit(bench="c2_train_step_torch_cudagraph_fp32", error=str(e)[:200])


def torch_rank(q, g, tile=8192):
    """Stock torch on the same GPU: bf16 cuBLAS GEMM tile + compare-count against the gt score."""
    qb, gb = q.bfloat16(), g.bfloat16()
    sq = (gb.float() ** 2).sum(1)
    N = q.shape[0]
    rank = torch.empty(N, dtype=torch.int64, device=q.device)
    idx = torch.arange(N, device=q.device)
    for s in range(0, N, tile):
        e = min(N, s + tile)
        d = sq[None, :] - 2.0 * (qb[s:e] @ gb.t()).float()
        d0 = d[torch.arange(e - s, device=q.device), idx[s:e]]
        rank[s:e] = (d < d0[:, None]).sum(1)
    return rank


def config3_and_5():
    T, V = make_retrieval_pair(10000, 10000, 512, seed=1023)
    q, g = T.to(dev), V.to(dev)
    for prec in ("bf16", "exact"):
        us, nl = timeit(lambda: ops.rank_eval(q, g, [1, 5, 10], precision=prec), iters=20)
        emit(bench="c3_rank_10kx10k", precision=prec, us=us, launches=nl, pairs_per_s=1e8 / (us * 1e-6))
        # the same step replayed as one CUDA graph (17 dependent launches around a ~70 us
        # tensor-core kernel: the gaps between them are a third of the eager step)
        try:
            from vtc_b200.parallel import GraphedRankEval

            ge = GraphedRankEval(q, g, 10000, 10000, [1, 5, 10], "l2", prec)
            us, _ = timeit(lambda: ge(), iters=50)
            emit(bench="c3_rank_10kx10k_cudagraph", precision=prec, us=us,
                 launches=ge.launches_per_replay, pairs_per_s=1e8 / (us * 1e-6))
            ge.close()
        except Exception as e:  # noqa: BLE001
            emit(bench="c3_rank_10kx10k_cudagraph", precision=prec, error=str(e)[:200])
    us, _ = timeit(lambda: torch_rank(q, g), iters=10)
    emit(bench="c3_rank_10kx10k_torch_bf16_eager", us=us, pairs_per_s=1e8 / (us * 1e-6))

    # config 5, one GPU's share: 10k queries x 125k gallery rows (1M / 8), k = 11
    T, V = make_retrieval_pair(10000, 125000, 512, seed=1023)
    q, g = T.to(dev), V.to(dev)
    for prec in ("bf16", "exact"):
        us, nl = timeit(lambda: ops.sim_topk(q, g, 11, precision=prec), iters=10)
        emit(bench="c5_topk_10kx125k_k11", precision=prec, us=us, launches=nl,
             pairs_per_s=1.25e9 / (us * 1e-6), tflops=2 * 1.25e9 * 512 / (us * 1e-6) / 1e12)

    def torch_topk():
        qb, gb = q.bfloat16(), g.bfloat16()
        sq = (gb.float() ** 2).sum(1)
        out = []
        for s in range(0, 10000, 2048):
            d = sq[None, :] - 2.0 * (qb[s:s + 2048] @ gb.t()).float()
            out.append(torch.topk(d, 11, dim=1, largest=False))
        return out

    us, _ = timeit(torch_topk, iters=5)
    emit(bench="c5_topk_10kx125k_k11_torch_bf16_eager", us=us, pairs_per_s=1.25e9 / (us * 1e-6))

    # the headline size with stock torch (same-box bar for bench.py's value)
    T, V = make_retrieval_pair(100000, 100000, 512, seed=1023)
    q, g = T.to(dev), V.to(dev)
    us, _ = timeit(lambda: torch_rank(q, g), iters=3, warm=1)
    emit(bench="headline_rank_100kx100k_torch_bf16_eager", us=us, pairs_per_s=1e10 / (us * 1e-6))


if __name__ == "__main__":
    _ffi.load()
    which = sys.argv[1:] or ["c2", "c35"]
    if "c2" in which:
        config2()
    if "c2" in which or "train" in which:
        config2_train()
    if "c35" in which:
        config3_and_5()
