#!/usr/bin/env python
"""Multi-GPU parity check, launched under torchrun by tests/test_multigpu.py (one rank per GPU,
NCCL): row-sharded rank eval and gallery-sharded top-k against the CPU oracle.  Lives under tests/
because it is a checker: nothing outside tests/, smoke() and bench.py's CPU baseline touches oracle/."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import vtc_oracle as O  # noqa: E402
from vtc_b200.parallel import (GraphedRankEval, PipelinedRankEval, shard_bounds,  # noqa: E402
                               sharded_rank_eval, sharded_topk)
from vtc_b200.synthetic import make_retrieval_pair  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for (N, M, D, prec) in ((3001, 3001, 512, "exact"), (2000, 5003, 256, "bf16"), (4096, 4096, 768, "bf16")):
        T, V = make_retrieval_pair(N, M, D, sigma=4.0, seed=N)
        qs, qe = shard_bounds(N, world, rank)
        gs, ge = shard_bounds(M, world, rank)
        res = sharded_rank_eval(T[qs:qe].contiguous().to(dev), V[gs:ge].contiguous().to(dev), N, M,
                                precision=prec)
        Tq, Vq = (O.bf16_round(T), O.bf16_round(V)) if prec == "bf16" else (T, V)
        want = O.rank0_exact(Tq, Vq)
        got = res["rank0_local"].cpu().numpy()
        hits = res["hits"].cpu().numpy()
        medr = float(res["medr"].cpu()[0])
        good = (np.array_equal(got, want[qs:qe]) and
                list(hits) == [int((want < k).sum()) for k in (1, 5, 10)] and medr == O.medr(want))
        tv, ti = sharded_topk(T[:64].contiguous().to(dev), V[gs:ge].contiguous().to(dev), M, 11,
                              precision=prec)
        wi = O.topk_exact(Tq[:64], Vq, 11)[1]
        good = good and np.array_equal(ti.cpu().numpy(), wi)
        # the same step captured in a CUDA graph, replayed on NEW inputs (row-permuted problem)
        ev = GraphedRankEval(T[qs:qe].contiguous().to(dev), V[gs:ge].contiguous().to(dev), N, M,
                             precision=prec)
        T2, V2 = make_retrieval_pair(N, M, D, sigma=4.0, seed=N + 1)
        for _ in range(2):
            res2 = ev(T2[qs:qe].contiguous().to(dev), V2[gs:ge].contiguous().to(dev))
        Tq2, Vq2 = (O.bf16_round(T2), O.bf16_round(V2)) if prec == "bf16" else (T2, V2)
        want2 = O.rank0_exact(Tq2, Vq2)
        good = good and (np.array_equal(res2["rank0_local"].cpu().numpy(), want2[qs:qe]) and
                         list(res2["hits"].cpu().numpy()) == [int((want2 < k).sum()) for k in (1, 5, 10)]
                         and float(res2["medr"].cpu()[0]) == O.medr(want2))
        ev.close()
        # a stream of evaluations from pinned host shards, copies overlapped with the previous replay:
        # jobs alternate between the two problems, every result must be its own job's
        pipe = PipelinedRankEval(T[qs:qe].contiguous().to(dev), V[gs:ge].contiguous().to(dev), N, M,
                                 precision=prec)
        jobs = [(T, V, want), (T2, V2, want2), (T2, V2, want2), (T, V, want), (T, V, want)]
        pinned = {id(a): a[lo:hi].contiguous().pin_memory()
                  for a, lo, hi in ((T, qs, qe), (T2, qs, qe), (V, gs, ge), (V2, gs, ge))}
        results = [pipe.submit(pinned[id(a)], pinned[id(b)]) for a, b, _ in jobs] + [pipe.flush()]
        good = good and results[0] is None
        for (_, _, w), r in zip(jobs, results[1:]):
            good = good and (list(r["hits"].numpy()) == [int((w < k).sum()) for k in (1, 5, 10)]
                             and r["medr"] == O.medr(w))
        pipe.close()
        print(f"[rank {rank}/{world}] N={N} M={M} D={D} {prec}: {'OK' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    import faulthandler
    import traceback

    faulthandler.dump_traceback_later(int(os.environ.get("VTC_DIST_CHECK_TIMEOUT", "240")), exit=True)
    try:
        main()
    except SystemExit:
        raise
    except BaseException:
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)  # a failed rank must not wait in the NCCL destructors for its peers
