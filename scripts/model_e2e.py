#!/usr/bin/env python
"""Analytical model of the end-to-end number (RecallAtK.compute from pinned host memory)
next to the measured chunk sweep in profiles/r01_e2e_chunks.jsonl.  No GPU needed.

A query can be ranked once its own ground-truth gallery row has landed, so with gallery and query
chunks copied interleaved, after a fraction f of the transfer only f^2 of the pairs are rankable:
    finish >= max_i [ f_i * T + (1 - f_{i-1}^2) * W ] + calls_on_the_critical_path * overhead
T = transfer time of both sides, W = ranking time of the whole matrix, 2c - 1 library calls.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bound(fr, T, W):
    return max(fr[i] * T + (1.0 - fr[i - 1] ** 2) * W for i in range(1, len(fr)))


def main():
    T = 409.6e6 / 55e9 * 1e3      # ms: 409.6 MB over PCIe 5 x16 at ~55 GB/s
    W = 7.5                       # ms: tensor-core kernel + re-check of the whole 100k x 100k job
    per_call = 0.09               # ms of preparation per library call on the critical path (fitted)
    rows = [json.loads(ln) for ln in open(os.path.join(ROOT, "profiles", "r01_e2e_chunks.jsonl"))]
    print(f"T = {T:.2f} ms, W = {W:.2f} ms, continuous bound 0.5 T + 0.75 W = {0.5 * T + 0.75 * W:.2f} ms")
    print("chunks  calls  bound(equal)  +overhead  measured   bound(balanced)  +overhead")
    os.environ["VTC_PIPELINE_SCHEDULE"] = "balanced"
    from vtc_b200.model.metric import RecallAtK

    for r in rows:
        c = r["chunks_per_side"]
        eq = [i / c for i in range(c + 1)]
        bal = [x / 1e6 for x in RecallAtK._pipeline_bounds_2d(1_000_000, c)]
        calls = 2 * c - 1
        print(f"{c:6d} {calls:6d} {bound(eq, T, W):13.2f} {bound(eq, T, W) + calls * per_call:10.2f} "
              f"{r['e2e']['ms_per_step']:9.2f} {bound(bal, T, W):17.2f} "
              f"{bound(bal, T, W) + calls * per_call:10.2f}")


if __name__ == "__main__":
    main()
