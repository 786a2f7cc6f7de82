#!/usr/bin/env python
"""Per-launch times of one sharded retrieval-evaluation step on rank 0 and on a middle rank (profiling aid; run under
torchrun on a multi-GPU box).  The library's launch trace records one CUDA event after each of its
launches, so every interval also contains whatever ran in the stream before that launch (casts,
NCCL kernels): the sum is the step's device time.

    python -m torch.distributed.run --nproc-per-node 8 ... scripts/dist_trace.py [--d 512] [--precision bf16]
"""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from vtc_b200 import _ffi  # noqa: E402
from vtc_b200.parallel import shard_bounds, sharded_rank_eval  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--d", type=int, default=512)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    gen = torch.Generator(device=dev).manual_seed(1023)
    V = torch.nn.functional.normalize(torch.randn(a.n, a.d, generator=gen, device=dev), dim=1)
    T = torch.nn.functional.normalize(V + 6.0 * torch.randn(a.n, a.d, generator=gen, device=dev) / a.d ** 0.5, dim=1)
    qs, qe = shard_bounds(a.n, world, rank)
    q, g = T[qs:qe].contiguous(), V[qs:qe].contiguous()
    del T, V
    _ffi.load()
    for _ in range(3):
        sharded_rank_eval(q, g, a.n, a.n, (1, 5, 10), "l2", a.precision)
    torch.cuda.synchronize()
    runs = []
    for _ in range(5):
        dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        _ffi.trace_begin(torch.cuda.current_stream().cuda_stream)
        sharded_rank_eval(q, g, a.n, a.n, (1, 5, 10), "l2", a.precision)
        ev1.record()
        tr = _ffi.trace_end()
        torch.cuda.synchronize()
        runs.append((tr, ev0.elapsed_time(ev1) * 1e3))
    for who in sorted({0, world // 2}):   # the first rank and a middle one (remote rows on both sides)
        dist.barrier()
        if rank != who:
            continue
        n = len(runs[0][0])
        rows = [[runs[0][0][i][0], round(statistics.median(r[0][i][1] for r in runs if len(r[0]) == n), 2)]
                for i in range(n)]
        print(json.dumps({"what": "sharded_rank_eval eager", "rank": rank, "world": world, "d": a.d,
                          "precision": a.precision, "step_us": round(statistics.median(r[1] for r in runs), 1),
                          "traced_us": round(sum(r[1] for r in rows), 1), "launches": rows}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
