"""Multi-GPU retrieval evaluation: one process per GPU, torch.distributed (NCCL over NVLink 5 /
NVSwitch on the box, gloo in the CPU tests) for the plumbing.  SURVEY.md §8e.

The reference has no multi-GPU evaluation (its only parallelism is a nominal, shape-inconsistent
DataParallel, train.py:76-82); this is the explicit design that replaces it:

* row-sharded rank eval (BASELINE config 4, N ~ M): rank r owns query rows and gallery rows
  [r*n/g, (r+1)*n/g).  The gallery shards are all-gathered; every rank ranks ITS queries against
  each gallery chunk with `vtc_sim_rank(accumulate=1, col_offset=chunk start)` -- rank counts are
  additive over gallery chunks -- starting with its own chunk while the gather is in flight.
  The only other exchange is the reduction of the R@K hit counts and a gather of the int32 ranks
  for the median.
* gallery-sharded top-k (config 5, M >> N): queries are replicated, each rank scans its gallery
  shard with `vtc_sim_topk(col_offset=shard start)`, the [N, k] candidates are all-gathered and
  merged with `vtc_topk_merge`.  The gallery never moves.

The compute backend is injectable so that the host logic is testable on CPU with gloo
(tests/test_dist_gloo.py plugs the oracle in); the default backend is the CUDA library.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class _Phases:
    """Optional CUDA-event phase timing (VTC_PHASE_TIMING=1): where a sharded step spends its time."""

    def __init__(self, dev):
        self.on = bool(os.environ.get("VTC_PHASE_TIMING")) and dev.type == "cuda"
        self.marks = []
        self.dev = dev
        self.mark("start")

    def mark(self, name):
        if self.on:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.dev))
            self.marks.append((name, ev))

    def result(self):
        if not self.on:
            return None
        torch.cuda.synchronize(self.dev)
        return {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(self.marks[:-1], self.marks[1:])}


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced split of range(n): the first n % world shards get one extra row."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class CudaBackend:
    """The product backend: libvtc_b200.so through vtc_b200.ops."""

    def __init__(self):
        from . import ops

        self.ops = ops

    def gt_scores(self, q, g, row_offset, col_offset, metric, precision):
        return self.ops.gt_scores(q, g, None, row_offset, col_offset, metric, precision)

    def sim_rank(self, q, g, row_offset, col_offset, metric, precision, gt_score, rank0,
                 sq64=None, qq=None, sq64_out=None, qq_out=None):
        """Accumulates into rank0; gt_score None = every ground truth lies inside g (computed and
        returned).  sq64 / qq hand cached per-row quantities in, sq64_out / qq_out have this call
        compute and store them (vtc_sim_rank_prepared)."""
        _, gs = self.ops.sim_rank(q, g, None, row_offset, col_offset, metric, precision, gt_score,
                                  rank0, accumulate=True, sq64=sq64, qq=qq, sq64_out=sq64_out,
                                  qq_out=qq_out)
        return gs

    def rank_prepare(self, x, precision, want_sq64, want_qq, sq64_out=None, qq_out=None):
        return self.ops.rank_prepare(x, precision, want_sq64, want_qq, sq64_out, qq_out)

    def rank_finalize(self, rank0, gt_score, M_total, k_vals, want_medr):
        return self.ops.rank_finalize(rank0, gt_score, M_total, k_vals, want_medr)

    def sim_topk(self, q, g, k, metric, precision, col_offset):
        return self.ops.sim_topk(q, g, k, metric, precision, col_offset)

    def topk_merge(self, vals, idx):
        return self.ops.topk_merge(vals, idx)


def _world(group) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def _all_gather_padded(x: torch.Tensor, sizes: Sequence[int], group, async_op: bool = False):
    """all_gather of row-shards with unequal row counts (pads to the largest shard)."""
    world = len(sizes)
    mx = max(sizes)
    if x.shape[0] < mx:
        pad = torch.zeros((mx - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad])
    out = [torch.empty_like(x) for _ in range(world)]
    work = dist.all_gather(out, x.contiguous(), group=group, async_op=async_op)
    return out, work


def sharded_rank_eval(q_local: torch.Tensor, g_local: torch.Tensor, N_total: int, M_total: int,
                      k_vals: Sequence[int] = (1, 5, 10), metric: str = "l2",
                      precision: str = "exact", group=None, backend=None,
                      want_medr: bool = True) -> Dict[str, object]:
    """Row-sharded similarity + rank + R@K (+MedR).  q_local / g_local are this rank's
    shard_bounds() rows of the global query / gallery matrices; gt(t) = t (global row index).

    Per-row quantities are computed once, by the rank that owns the rows: the canonical ||x||^2 of a
    gallery shard come out of its owner's local ranking call and ride a second (800 KB) all_gather,
    so the calls against the gathered rows walk no row at all (vtc_sim_rank_prepared).

    Returns {"hits": int64 [nk] (global), "medr": float or None, "rank0_local": int32 [n_r],
             "num_queries": N_total}.  hits / medr are identical on every rank."""
    backend = backend or CudaBackend()
    world, rank = _world(group)
    qs, qe = shard_bounds(N_total, world, rank)
    assert q_local.shape[0] == qe - qs, "q_local does not match shard_bounds(N_total)"
    g_sizes = [shard_bounds(M_total, world, r)[1] - shard_bounds(M_total, world, r)[0]
               for r in range(world)]
    g_starts = [shard_bounds(M_total, world, r)[0] for r in range(world)]
    assert g_local.shape[0] == g_sizes[rank], "g_local does not match shard_bounds(M_total)"
    dev = q_local.device
    ph = _Phases(dev)
    if precision == "bf16" and q_local.dtype == torch.float32 and g_local.dtype == torch.float32:
        # the bf16 mode ranks the RN-even bf16 roundings of the inputs: round BEFORE the exchange
        # (identical results, half the NVLink bytes; bf16 rows of whole swizzle atoms are the
        # tensor-core operands in place)
        q_local = q_local.to(torch.bfloat16)
        g_local = g_local.to(torch.bfloat16)

    # gallery exchange: one all_gather into a single [world * max_shard, D] buffer, so that with
    # equal shards the remote rows form (at most) two contiguous ranges [0, gs) and [ge, M)
    work = None
    gathered = None
    mx = max(g_sizes)
    equal = all(sz == mx for sz in g_sizes)
    if world > 1:
        send = g_local
        if send.shape[0] < mx:
            pad = torch.zeros((mx - send.shape[0], send.shape[1]), dtype=send.dtype, device=dev)
            send = torch.cat([send, pad])
        gathered = torch.empty((world * mx, send.shape[1]), dtype=send.dtype, device=dev)
        work = dist.all_gather_into_tensor(gathered, send.contiguous(), group=group, async_op=True)
    ph.mark("cast+gather_issue")

    n_local = qe - qs
    gs0, ge0 = g_starts[rank], g_starts[rank] + g_sizes[rank]
    have_local = g_sizes[rank] > 0
    rank0 = torch.zeros(n_local, dtype=torch.int32, device=dev)
    if n_local == 0:
        # more ranks than query rows: take part in the collectives, rank nothing
        if work is not None:
            work.wait()
            if hasattr(backend, "rank_prepare") and precision in ("bf16", "exact"):
                _exchange_norms(backend, g_local, mx, world, precision, group, dev, have_local)
        gt_score = torch.empty(0, dtype=torch.float64, device=dev)
        return _finish_sharded(backend, rank0, gt_score, N_total, M_total, k_vals, world, group,
                               want_medr, ph)
    gt_local = (world == 1 or _gt_all_local(qs, qe, gs0, g_sizes[rank])) and have_local
    cached = hasattr(backend, "rank_prepare") and precision in ("bf16", "exact") and world > 1
    sq_local = qq = None
    if cached:
        sq_local = torch.zeros(mx, dtype=torch.float64, device=dev)   # padded like the shard
        qq = torch.empty(n_local, dtype=torch.float32, device=dev)
    local_done = False
    work_sq = sq_all = None
    if gt_local:
        # every ground truth is in our own chunk: rank against it while the gather is in flight;
        # the call also yields the ground-truth scores and the query-norm bounds.  The shard's
        # canonical norms are computed first, on their own, so that their (800 KB) all_gather runs
        # behind the local tensor-core pass instead of after it.
        if cached:
            backend.rank_prepare(g_local, precision, True, False, sq64_out=sq_local[:g_sizes[rank]])
            sq_all = torch.empty(world * mx, dtype=torch.float64, device=dev)
            work_sq = dist.all_gather_into_tensor(sq_all, sq_local, group=group, async_op=True)
            gt_score = backend.sim_rank(q_local, g_local, qs, gs0, metric, precision, None, rank0,
                                        sq64=sq_local[:g_sizes[rank]], qq_out=qq)
        else:
            gt_score = backend.sim_rank(q_local, g_local, qs, gs0, metric, precision, None, rank0)
        local_done = True
    else:
        if have_local:
            gt_score = backend.gt_scores(q_local, g_local, qs, gs0, metric, precision)
        else:  # more ranks than gallery rows: every ground truth lives elsewhere
            gt_score = torch.full((n_local,), float("nan"), dtype=torch.float64, device=dev)
        if cached:
            if have_local:
                backend.rank_prepare(g_local, precision, True, False,
                                     sq64_out=sq_local[:g_sizes[rank]])
            backend.rank_prepare(q_local, precision, False, True, qq_out=qq)
    ph.mark("gt+local_rank")
    if work is not None:
        if cached:
            # the owners' norms: [world * mx] fp64, laid out like the gathered rows
            if work_sq is not None:
                work_sq.wait()
            else:
                sq_all = torch.empty(world * mx, dtype=torch.float64, device=dev)
                dist.all_gather_into_tensor(sq_all, sq_local, group=group)
        work.wait()
        ph.mark("gather_wait")
        # remote row ranges as (start row in the global gallery, buffer start, buffer end)
        if equal:
            remote = [(0, 0, gs0), (ge0, ge0, M_total)]
        else:
            remote = [(g_starts[r], r * mx, r * mx + g_sizes[r]) for r in range(world) if r != rank]
        remote = [(st, b0, b1) for st, b0, b1 in remote if b1 > b0]
        if not gt_local:
            # ground truths that live in another rank's shard (N != M splits): fill in where still NaN
            for st, b0, b1 in remote:
                other = backend.gt_scores(q_local, gathered[b0:b1], qs, st, metric, precision)
                gt_score = torch.where(torch.isnan(gt_score), other, gt_score)
        kw = (lambda b0, b1: {"sq64": sq_all[b0:b1], "qq": qq}) if cached else (lambda b0, b1: {})
        if not local_done and have_local:
            backend.sim_rank(q_local, g_local, qs, gs0, metric, precision, gt_score, rank0,
                             **({"sq64": sq_local[:g_sizes[rank]], "qq": qq} if cached else {}))
        for st, b0, b1 in remote:
            backend.sim_rank(q_local, gathered[b0:b1], qs, st, metric, precision, gt_score, rank0,
                             **kw(b0, b1))
    elif not local_done and have_local:
        backend.sim_rank(q_local, g_local, qs, gs0, metric, precision, gt_score, rank0)
    ph.mark("remote_rank")
    if world == 1:
        hits, medr = backend.rank_finalize(rank0, gt_score, M_total, list(k_vals), want_medr)
        return {"hits": hits.clone(), "medr": medr, "rank0_local": rank0, "num_queries": N_total,
                "phases_ms": ph.result()}
    return _finish_sharded(backend, rank0, gt_score, N_total, M_total, k_vals, world, group,
                           want_medr, ph)


def _exchange_norms(backend, g_local, mx, world, precision, group, dev, have_local):
    """A rank without query rows still owns gallery rows: contribute their norms to the exchange."""
    sq_local = torch.zeros(mx, dtype=torch.float64, device=dev)
    if have_local:
        backend.rank_prepare(g_local, precision, True, False, sq64_out=sq_local[:g_local.shape[0]])
    sq_all = torch.empty(world * mx, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(sq_all, sq_local, group=group)


def _finish_sharded(backend, rank0, gt_score, N_total, M_total, k_vals, world, group, want_medr, ph):
    """The exchange at the end of a sharded step (shared by the default and the prepared path)."""
    medr = None
    if want_medr:
        backend.rank_finalize(rank0, gt_score, M_total, [], False)
        q_sizes = [shard_bounds(N_total, world, r)[1] - shard_bounds(N_total, world, r)[0]
                   for r in range(world)]
        if all(sz == q_sizes[0] for sz in q_sizes):
            # equal shards: one collective straight into the full vector (no per-rank copies, no cat)
            full = torch.empty(N_total, dtype=rank0.dtype, device=rank0.device)
            dist.all_gather_into_tensor(full, rank0.contiguous(), group=group)
        else:
            allr, _ = _all_gather_padded(rank0, q_sizes, group)
            full = torch.cat([allr[r][:q_sizes[r]] for r in range(world)])
        hits, medr = backend.rank_finalize(full, None, M_total, list(k_vals), True)
        hits = hits.clone()
    else:
        hits, _ = backend.rank_finalize(rank0, gt_score, M_total, list(k_vals), False)
        hits = hits.clone()
        dist.all_reduce(hits, op=dist.ReduceOp.SUM, group=group)
    ph.mark("finalize+collectives")
    return {"hits": hits, "medr": medr, "rank0_local": rank0, "num_queries": N_total,
            "phases_ms": ph.result()}


class GraphedRankEval:
    """One sharded_rank_eval step captured in a CUDA graph and replayed.

    At 8 GPUs a 100k x 100k evaluation is ~1 ms of GPU work per rank behind ~40 kernel launches
    and three collectives, which is launch-bound from Python; every shape is fixed, the library
    never synchronises (overflow fallbacks are device-side flags) and NCCL collectives are
    capturable, so the whole step -- gather, local + remote rank passes, rank exchange, hit counts
    and median -- becomes a single graph launch.  Inputs live in static buffers: pass new shards
    to __call__ to have them copied in (device-to-device) before the replay.

    All ranks must construct and call it collectively, and close() it before the process group is
    destroyed."""

    def __init__(self, q_local: torch.Tensor, g_local: torch.Tensor, N_total: int, M_total: int,
                 k_vals: Sequence[int] = (1, 5, 10), metric: str = "l2", precision: str = "exact",
                 group=None, want_medr: bool = True):
        if q_local.device.type != "cuda":
            raise RuntimeError("GraphedRankEval needs CUDA tensors")
        from . import ops

        self.q = q_local.detach().clone()
        self.g = g_local.detach().clone()
        self._args = (N_total, M_total, tuple(k_vals), metric, precision, group, None, want_medr)
        dev = self.q.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):  # allocator, workspace and NCCL communicator warm-up
                sharded_rank_eval(self.q, self.g, *self._args)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _ffi

        before = dict(ops._ws)
        l0 = _ffi.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = sharded_rank_eval(self.q, self.g, *self._args)
        self.launches_per_replay = _ffi.launch_count() - l0  # library kernels inside the graph
        # workspaces created while capturing belong to the graph: keep them alive here and out of
        # the shared cache, where a later, larger request would free them under the graph
        self._held = [ops._ws.pop(k) for k in list(ops._ws) if ops._ws[k] is not before.get(k)]

    def __call__(self, q_local: Optional[torch.Tensor] = None,
                 g_local: Optional[torch.Tensor] = None) -> Dict[str, object]:
        if q_local is not None and q_local.data_ptr() != self.q.data_ptr():
            self.q.copy_(q_local, non_blocking=True)
        if g_local is not None and g_local.data_ptr() != self.g.data_ptr():
            self.g.copy_(g_local, non_blocking=True)
        self.graph.replay()
        return self.out

    def close(self) -> None:
        """Destroy the graph.  NCCL keeps a communicator alive while a captured graph still refers
        to it, so call this before dist.destroy_process_group() (which otherwise waits forever)."""
        if self.graph is not None:
            torch.cuda.synchronize(self.q.device)
            self.graph.reset()
            self.graph = None
        self.out = None
        self._held = []


class PipelinedRankEval:
    """A stream of sharded evaluations from pinned HOST shards with the copies hidden: two captured
    steps (GraphedRankEval) with their own static input buffers alternate, so that the host-to-device
    copy of evaluation k + 1 (on a copy stream) runs while evaluation k is being ranked, and the
    16-byte result of evaluation k is read back while k + 1 runs.

        pipe = PipelinedRankEval(q_dev, g_dev, N, M, precision="bf16")     # collective
        for q_host, g_host in jobs:            # pinned fp32 shards of this rank
            done = pipe.submit(q_host, g_host) # result of the evaluation submitted BEFORE this one
        last = pipe.flush()
    Every rank must submit the same number of jobs.  close() before destroy_process_group()."""

    def __init__(self, q_local: torch.Tensor, g_local: torch.Tensor, N_total: int, M_total: int,
                 k_vals: Sequence[int] = (1, 5, 10), metric: str = "l2", precision: str = "exact",
                 group=None, want_medr: bool = True):
        dev = q_local.device
        self.dev = dev
        self.steps = [GraphedRankEval(q_local, g_local, N_total, M_total, k_vals, metric, precision,
                                      group, want_medr) for _ in range(2)]
        self.copy = torch.cuda.Stream(dev)
        self.ready = [None, None]      # inputs of slot i have landed (recorded on the copy stream)
        self.consumed = [None, None]   # replay of slot i has finished reading its inputs (main stream)
        self.host = [torch.empty(len(k_vals), dtype=torch.int64).pin_memory() for _ in range(2)]
        self.host_medr = [torch.empty(1, dtype=torch.float64).pin_memory() for _ in range(2)]
        self.read = [None, None]       # device-to-host read of slot i's result is complete
        self.k = 0
        self.pending = None

    def _result(self, slot):
        self.read[slot].synchronize()
        return {"hits": self.host[slot].clone(), "medr": float(self.host_medr[slot][0])}

    def submit(self, q_host: torch.Tensor, g_host: torch.Tensor) -> Optional[Dict[str, object]]:
        slot = self.k & 1
        step = self.steps[slot]
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy):
            if self.consumed[slot] is not None:
                self.copy.wait_event(self.consumed[slot])   # the replay two jobs ago is done with them
            step.q.copy_(q_host, non_blocking=True)
            step.g.copy_(g_host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy)
            self.ready[slot] = ev
        main.wait_event(self.ready[slot])
        out = step()                                         # one graph launch
        done = torch.cuda.Event()
        done.record(main)
        self.consumed[slot] = done
        self.host[slot].copy_(out["hits"], non_blocking=True)
        if out.get("medr") is not None:
            self.host_medr[slot].copy_(out["medr"], non_blocking=True)
        rd = torch.cuda.Event()
        rd.record(main)
        self.read[slot] = rd
        prev, self.pending = self.pending, slot
        self.k += 1
        return None if prev is None else self._result(prev)

    def flush(self) -> Optional[Dict[str, object]]:
        prev, self.pending = self.pending, None
        return None if prev is None else self._result(prev)

    def close(self) -> None:
        torch.cuda.synchronize(self.dev)
        for s in self.steps:
            s.close()


def _gt_all_local(qs: int, qe: int, g_start: int, g_size: int) -> bool:
    return qs >= g_start and qe <= g_start + g_size


def sharded_topk(q: torch.Tensor, g_local: torch.Tensor, M_total: int, k: int, metric: str = "l2",
                 precision: str = "bf16", group=None, backend=None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Gallery-sharded streaming top-k: every rank holds all queries `q` and its shard_bounds()
    rows of the gallery.  Returns the merged (vals [N,k], idx int64 [N,k]) on every rank."""
    backend = backend or CudaBackend()
    world, rank = _world(group)
    gs, ge = shard_bounds(M_total, world, rank)
    assert g_local.shape[0] == ge - gs, "g_local does not match shard_bounds(M_total)"
    vals, idx = backend.sim_topk(q, g_local, k, metric, precision, gs)
    if world == 1:
        return vals, idx
    av = [torch.empty_like(vals) for _ in range(world)]
    ai = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(av, vals.contiguous(), group=group)
    dist.all_gather(ai, idx.contiguous(), group=group)
    return backend.topk_merge(torch.stack(av), torch.stack(ai))
