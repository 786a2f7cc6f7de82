#!/usr/bin/env python
"""BASELINE config 5 on N GPUs: 10k queries x 1M gallery x 512, k = 11, gallery-sharded streaming
top-k (`vtc_b200.parallel.sharded_topk`): every rank scans its M/N gallery rows, the [N, k]
candidates are all-gathered and merged.  Run under torchrun; rank 0 prints one JSON line."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtc_b200 import _ffi  # noqa: E402
from vtc_b200.parallel import shard_bounds, sharded_topk  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000)
    ap.add_argument("--m", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=512)
    ap.add_argument("--k", type=int, default=11)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _ffi.load()
    gs, ge = shard_bounds(a.m, world, rank)
    # synthetic unit-norm gallery shard (seeded per rank) and replicated noisy-copy queries of the
    # first rows of shard 0, generated on the device: 1M x 512 on the host would take minutes
    g = torch.Generator(device=dev).manual_seed(1023 + rank)
    G = torch.nn.functional.normalize(torch.randn(ge - gs, a.d, generator=g, device=dev), dim=1)
    base = G[:a.n].clone() if rank == 0 else torch.empty(a.n, a.d, device=dev)
    if world > 1:
        dist.broadcast(base, 0)  # queries = noisy copies of the first rows of shard 0
    gq = torch.Generator(device=dev).manual_seed(7)
    Q = torch.nn.functional.normalize(
        base + 6.0 * torch.randn(a.n, a.d, generator=gq, device=dev) / a.d ** 0.5, dim=1)

    def step():
        return sharded_topk(Q, G, a.m, a.k, precision=a.precision)

    for _ in range(3):
        vals, idx = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        vals, idx = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # sanity: R@1 of the noisy copies against the 1M gallery (the source row of query t is row t)
    top1_self = int((idx[:, 0] == torch.arange(a.n, device=dev)).sum()) if rank == 0 else 0
    if rank == 0:
        print(json.dumps({"bench": "c5_topk_sharded", "n_gpus": world, "N": a.n, "M": a.m, "D": a.d,
                          "k": a.k, "precision": a.precision, "ms_per_step": ms.item(),
                          "pairs_per_s": a.n * a.m / (ms.item() * 1e-3),
                          "tflops": 2.0 * a.n * a.m * a.d / (ms.item() * 1e-3) / 1e12,
                          "r_at_1": top1_self / a.n}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
