"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in vtc_b200/parallel.py.  The CUDA
backend is swapped for an oracle-backed one so the sharding / exchange / reduction code runs here
exactly as it does over NCCL on the box."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """Same interface as vtc_b200.parallel.CudaBackend, computed by the CPU oracle."""

    def __init__(self):
        from oracle import vtc_oracle as O

        self.O = O

    def _m(self, metric):
        return self.O.METRIC_L2 if metric == "l2" else self.O.METRIC_DOT

    def gt_scores(self, q, g, row_offset, col_offset, metric, precision, gt=None):
        full = self.O.scores64(q, g, self._m(metric)) if g.shape[0] else np.zeros((q.shape[0], 0))
        out = np.full(q.shape[0], np.nan)
        for t in range(q.shape[0]):
            j = (int(gt[t]) if gt is not None else t + row_offset) - col_offset
            if 0 <= j < g.shape[0]:
                out[t] = full[t, j]
        return torch.from_numpy(out)

    def rank_prepare(self, x, precision, want_sq64, want_qq, sq64_out=None, qq_out=None):
        sq = self.O.sqnorm64(x.numpy()) if x.shape[0] else np.zeros(0)
        if sq64_out is not None:
            sq64_out.copy_(torch.from_numpy(sq))
        if qq_out is not None:
            qq_out.copy_(torch.from_numpy((sq * (1 + 1e-4)).astype(np.float32)))
        return (sq64_out if sq64_out is not None else (torch.from_numpy(sq) if want_sq64 else None),
                qq_out if qq_out is not None else
                (torch.from_numpy((sq * (1 + 1e-4)).astype(np.float32)) if want_qq else None))

    def sim_rank(self, q, g, row_offset, col_offset, metric, precision, gt_score, rank0,
                 sq64=None, qq=None, sq64_out=None, qq_out=None, gt=None):
        self.rank_calls = getattr(self, "rank_calls", 0) + 1
        if sq64_out is not None:  # the owner's call computes and stores the cached quantities
            sq64_out.copy_(torch.from_numpy(self.O.sqnorm64(g.numpy()) if g.shape[0] else np.zeros(0)))
        if qq_out is not None:
            qq_out.copy_(torch.from_numpy((self.O.sqnorm64(q.numpy()) * (1 + 1e-4)).astype(np.float32)))
        if sq64 is not None:
            # cached quantities arrive as slices of per-buffer arrays: they must belong to THESE rows
            self.prepared_calls = getattr(self, "prepared_calls", 0) + 1
            assert qq is not None or qq_out is not None   # (the norm bounds: handed in or asked for)
            fin = np.isfinite(sq64.numpy())
            np.testing.assert_array_equal(sq64.numpy()[fin], self.O.sqnorm64(g.numpy())[fin])
            if qq is not None:
                assert qq.shape == (q.shape[0],) and (qq.numpy() >= self.O.sqnorm64(q.numpy())).all()
        if gt_score is None:
            gt_score = self.gt_scores(q, g, row_offset, col_offset, metric, precision, gt=gt)
        full = self.O.scores64(q, g, self._m(metric))
        if sq64 is not None and metric == "l2":   # the library scores with the norms it is handed
            full[:, ~np.isfinite(sq64.numpy())] = np.inf
        d0 = gt_score.numpy()
        gts = gt
        for t in range(q.shape[0]):
            gt = int(gts[t]) if gts is not None else t + row_offset
            jg = np.arange(g.shape[0]) + col_offset
            # like the tensor-core epilogue: a column that scores strictly below d(t,gt) counts
            # whatever its id (the true ground-truth column never does); ids only break exact ties
            c = (full[t] < d0[t]).sum() + ((full[t] == d0[t]) & (jg < gt)).sum()
            rank0[t] += int(c)
        return gt_score

    def rank_finalize(self, rank0, gt_score, M_total, k_vals, want_medr):
        if gt_score is not None:
            rank0[torch.isnan(gt_score)] = M_total
        hits = torch.tensor([int((rank0 < k).sum()) for k in k_vals], dtype=torch.int64)
        medr = torch.tensor([self.O.medr(rank0.numpy())]) if want_medr else None
        return hits, medr

    def sim_topk(self, q, g, k, metric, precision, col_offset):
        v, i = self.O.topk_exact(q, g, k, metric=self._m(metric), col_offset=col_offset, threads=1)
        return torch.from_numpy(v).float(), torch.from_numpy(i)

    def topk_merge(self, vals, idx):
        parts, N, k = vals.shape
        ov = torch.empty(N, k)
        oi = torch.empty(N, k, dtype=torch.int64)
        for t in range(N):
            cand = [(float(vals[p, t, j]), int(idx[p, t, j])) for p in range(parts) for j in range(k)
                    if idx[p, t, j] >= 0]
            cand.sort()
            cand = cand[:k] + [(float("inf"), -1)] * (k - len(cand[:k]))
            ov[t] = torch.tensor([c[0] for c in cand])
            oi[t] = torch.tensor([c[1] for c in cand])
        return ov, oi


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _with_duplicates(V, dups):
    """Exact ties across shards: copies of gallery rows below and above the original's shard."""
    if dups:
        M = V.shape[0]
        for src, dst in ((2 * M // 5, 3), (2 * M // 5, M - 10), (2 * M // 5, 2 * M // 5 + 1),
                         (M - 20, 10), (5, M // 2)):
            V[dst] = V[src]
    return V


def _worker(rank, world, port, N, M, out_q, dups=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vtc_b200.parallel import shard_bounds, sharded_rank_eval, sharded_topk
        from vtc_b200.synthetic import make_retrieval_pair

        T, V = make_retrieval_pair(N, M, 64, sigma=2.0, seed=13)
        V = _with_duplicates(V, dups)
        qs, qe = shard_bounds(N, world, rank)
        gs, ge = shard_bounds(M, world, rank)
        be = OracleBackend()
        res = sharded_rank_eval(T[qs:qe].contiguous(), V[gs:ge].contiguous(), N, M, backend=be)
        # every call against gathered rows gets the owners' norms (none when a rank has no queries)
        assert getattr(be, "prepared_calls", 0) >= (1 if qe > qs and M > ge - gs else 0)
        tv, ti = sharded_topk(T[:16].contiguous(), V[gs:ge].contiguous(), M, 5, backend=be)
        out_q.put((rank, res["hits"].numpy(), float(res["medr"][0]), res["rank0_local"].numpy(),
                   ti.numpy()))
    finally:
        dist.destroy_process_group()


def _run(N, M, world=2, dups=False):
    from oracle import vtc_oracle as O
    from vtc_b200.parallel import shard_bounds
    from vtc_b200.synthetic import make_retrieval_pair

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, M, q, dups)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    T, V = make_retrieval_pair(N, M, 64, sigma=2.0, seed=13)
    V = _with_duplicates(V, dups)
    want = O.rank0_exact(T, V)
    want_hits = [int((want < k).sum()) for k in (1, 5, 10)]
    _, want_topk = O.topk_exact(T[:16], V, 5)
    for rank, hits, medr, local, ti in got:
        qs, qe = shard_bounds(N, world, rank)
        np.testing.assert_array_equal(local, want[qs:qe])
        np.testing.assert_array_equal(hits, want_hits)
        assert medr == O.medr(want)
        np.testing.assert_array_equal(ti, want_topk)


def test_row_sharded_eval_square_world2():
    _run(101, 101)


def test_row_sharded_eval_rectangular_world2():
    """N != M: some ground truths live in the other rank's gallery shard."""
    _run(90, 151)


def test_row_sharded_eval_equal_shards_world2():
    """Equal shards: the gathered gallery is used as two contiguous remote ranges."""
    _run(100, 100)


def test_row_sharded_eval_unequal_and_tiny_world2():
    """Unequal shards (per-shard slices of the gathered norms, padding rows in between) and problems
    smaller than the world: a rank without query rows / without gallery rows still takes part in
    every collective."""
    _run(101, 101)
    _run(1, 1)
    _run(1, 3)


def test_row_sharded_eval_world4_middle_ranks_and_cross_shard_ties():
    """Middle ranks: the remote rows lie on both sides of the own shard (two ranges, world - 1 with
    unequal shards) -- equal shards, unequal shards, ground truths in other ranks' shards, and exact
    ties with rows of lower and of higher shards (the tie-break by column id must come out as in the
    global gallery)."""
    _run(100, 100, world=4)
    _run(100, 100, world=4, dups=True)
    _run(103, 103, world=4, dups=True)
    _run(90, 151, world=4, dups=True)
    _run(3, 5, world=4)
