#!/usr/bin/env python
"""Where a small dense tensor-core product (the CAM's linears) spends its time (profiling aid):
VTC_DBG_PROF counters of the MMA issuer / first epilogue warp per CTA next to the launch's warm time.
    python scripts/linear_prof.py"""
import json
import os
import statistics
import sys

os.environ["VTC_DBG_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vtc_b200 import _ffi, ops  # noqa: E402

dev = torch.device("cuda", 0)
_ffi.load()
g = torch.Generator(device=dev).manual_seed(0)
for name, rows, fin, fout, act in (("qkv", 1536, 512, 1536, 0), ("out", 1536, 512, 512, 0),
                                   ("fc", 1536, 512, 2048, 1), ("proj", 1536, 2048, 512, 0)):
    for prec in ("bf16", "exact"):
        x = torch.randn(rows, fin, generator=g, device=dev)
        w = torch.randn(fout, fin, generator=g, device=dev) * 0.02
        b = torch.randn(fout, generator=g, device=dev)
        for _ in range(5):
            ops.linear(x, w, b, None, act, prec)
        torch.cuda.synchronize()
        _ffi.debug_prof_read()
        _ffi.trace_begin(torch.cuda.current_stream().cuda_stream)
        ops.linear(x, w, b, None, act, prec)
        tr = _ffi.trace_end()
        prof = [p for p in _ffi.debug_prof_read(256) if p[4] > 0]
        med = lambda i: statistics.median(p[i] for p in prof)  # noqa: E731
        print(json.dumps({"linear": name, "precision": prec, "shape": [rows, fin, fout],
                          "gemm_us_trace": tr[-1][1], "ctas": len(prof), "tiles_per_cta": med(4),
                          "issuer_ticks": med(0), "issuer_ns": med(1), "wait_acc": med(2), "wait_ld": med(3),
                          "epi_wait": med(5), "epi_ticks": med(6)}), flush=True)
