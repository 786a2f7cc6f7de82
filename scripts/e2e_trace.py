#!/usr/bin/env python
"""Launch trace of one end-to-end RecallAtK.compute from pinned host memory (profiling aid):
where the main stream waits for the copy stream and what each library call costs.
    python scripts/e2e_trace.py [chunks]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vtc_b200 import _ffi  # noqa: E402
from vtc_b200.model.metric import RecallAtK  # noqa: E402
from vtc_b200.synthetic import make_retrieval_pair  # noqa: E402

T, V = make_retrieval_pair(100_000, 100_000, 512, seed=1023)
Tp, Vp = T.pin_memory(), V.pin_memory()
for c in ([int(x) for x in sys.argv[1:]] or [RecallAtK.PIPELINE_CHUNKS_2D]):
    RecallAtK.PIPELINE_CHUNKS_2D = c
    m = RecallAtK("videos", "titles", [1, 5, 10], precision="bf16")
    for _ in range(2):
        m.compute(Vp, Tp)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        m.compute(Vp, Tp)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    _ffi.trace_begin(torch.cuda.current_stream().cuda_stream)
    m.compute(Vp, Tp)
    tr = _ffi.trace_end()
    by = {}
    for w, us in tr:
        by[w] = by.get(w, 0.0) + us
    print(json.dumps({"chunks": c, "e2e_ms": round(ms, 3), "traced_ms": round(sum(u for _, u in tr) / 1e3, 3),
                      "n_launches": len(tr), "by_site_us": {k: round(v, 1) for k, v in by.items()},
                      "launches": [[w, round(u, 1)] for w, u in tr]}), flush=True)
