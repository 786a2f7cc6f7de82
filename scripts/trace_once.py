#!/usr/bin/env python
"""Per-launch times of the library's paths as they run back to back in one stream (profiling aid).

Uses the library's launch trace (vtc_trace_begin / vtc_trace_end: one CUDA event after every launch;
include/vtc_b200.h).  Unlike ncu's launch list (serialised, cold caches) these are the times the
launches take inside the warm step, so they add up to the step's own device time (+ ~1 us per event).

    python scripts/trace_once.py [cam] [c3] [topk] [nce] [rank]     (default: all but rank)
One JSON line per traced call: {"what", "total_us", "launches": [[where, us], ...]} (median of 5).
"""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from vtc_b200 import _ffi, ops  # noqa: E402


def traced(what, fn, reps=5, warm=3, **meta):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    runs = []
    for _ in range(reps):
        _ffi.trace_begin(torch.cuda.current_stream().cuda_stream)
        fn()
        runs.append(_ffi.trace_end())
    n = len(runs[0])
    rows = []
    for i in range(n):
        us = [r[i][1] for r in runs if len(r) == n]
        rows.append([runs[0][i][0], round(statistics.median(us), 2)])
    print(json.dumps({"what": what, **meta, "total_us": round(sum(r[1] for r in rows), 1),
                      "n_launches": n, "launches": rows}), flush=True)


def main():
    which = set(sys.argv[1:]) or {"cam", "c3", "topk", "nce"}
    dev = torch.device("cuda", 0)
    _ffi.load()
    gen = torch.Generator(device=dev).manual_seed(1023)

    def pair(n, m, d, sigma=6.0):
        V = torch.nn.functional.normalize(torch.randn(m, d, generator=gen, device=dev), dim=1)
        T = torch.nn.functional.normalize(
            V[:n] + sigma * torch.randn(n, d, generator=gen, device=dev) / d ** 0.5, dim=1)
        return T, V

    if "cam" in which:
        from vtc_b200.model import PretrainedCLIP_finaltf
        from vtc_b200.synthetic import make_cam_inputs
        main_f, aux = make_cam_inputs(256, 5, 512, seed=1023)
        m, x = main_f.to(dev), aux.to(dev)
        for prec in ("bf16", "exact"):
            cam = PretrainedCLIP_finaltf(512, precision=prec).to(dev).eval()
            for blk in cam.final_transformer.resblocks:  # make the transformer non-trivial
                torch.nn.init.normal_(blk.mlp.c_proj.weight, std=0.02)
                torch.nn.init.normal_(blk.attn.out_proj.weight, std=0.02)
            with torch.no_grad():
                traced("c2_cam_adapt_feature", lambda: cam._adapt_feature(m, x), precision=prec)
    if "c3" in which:
        T, V = pair(10_000, 10_000, 512)
        for prec in ("bf16", "exact"):
            traced("c3_rank_10kx10k", lambda: ops.rank_eval(T, V, [1, 5, 10], precision=prec),
                   precision=prec)
    if "rank" in which:
        T, V = pair(100_000, 100_000, 512)
        for prec in ("bf16", "exact"):
            traced("rank_100kx100k", lambda: ops.rank_eval(T, V, [1, 5, 10], precision=prec),
                   precision=prec)
    if "topk" in which:
        T, V = pair(10_000, 125_000, 512)
        for prec in ("bf16", "exact"):
            traced("c5_topk_10kx125k_k11", lambda: ops.sim_topk(T, V, 11, precision=prec),
                   precision=prec)
    if "nce" in which:
        for n in (256, 4096, 16384):
            a = torch.nn.functional.normalize(torch.randn(n, 512, generator=gen, device=dev), dim=1)
            b = torch.nn.functional.normalize(a + 0.5 * torch.randn(n, 512, generator=gen, device=dev), dim=1)
            for prec in ("bf16", "exact"):
                traced("infonce_fwd", lambda: ops.infonce_fwd(a, b, 100.0, precision=prec), n=n,
                       precision=prec)


if __name__ == "__main__":
    main()
