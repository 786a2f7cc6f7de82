"""Drop-in for the reference's model/metric.py::RecallAtK (model/metric.py:103-187).

Same constructor, attributes and methods (`set_writer`, `reset`, `update`, `avg`, `result`,
`compute`), same result keys.  What changes underneath:

* `compute(features_a, features_b)` no longer builds a faiss index and a Python loop
  (model/metric.py:140-160); it calls the fused similarity + rank-of-ground-truth kernel
  (vtc_sim_rank) and counts `rank0 < k` on the device (vtc_rank_finalize).  R@k is identical to
  "gt index is inside the top-k list" because rank0 is the position of the ground truth in a
  stable ascending sort of the same exact-L2 scores.
* `update` keeps the per-batch features on the GPU instead of `.cpu()` per batch (:129-130).
* `compute_full` additionally returns per-query ranks and the median rank (MedR), which the
  reference does not define.

Inputs may be numpy arrays / CPU tensors (as the reference passes them): they are staged through
pinned host memory to the current CUDA device.  There is no CPU compute path.
"""
from __future__ import annotations

import collections.abc
import os
import time
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import ops

ArrayLike = Union[np.ndarray, torch.Tensor]

__all__ = ["BaseMetric", "ScalarPerBatchMetric", "LossMetric", "RecallAtK", "MetricTracker"]


class MetricTracker:
    """model/metric.py:10-42, with the API the trainer drives it through (trainer/trainer.py:49-54):
    `MetricTracker(*metrics)`, `add_metric`, `set_writer`, `reset`, `update`, `avg`, `result`;
    `metrics` maps metric name -> metric, in insertion order."""

    def __init__(self, *metrics):
        self.metrics = {}
        for metric in metrics:
            self.add_metric(metric)
        self.reset()

    def add_metric(self, metric):
        self.metrics[metric.name] = metric

    def _each(self):
        return list(self.metrics.values())

    def set_writer(self, writer):
        for metric in self._each():
            metric.set_writer(writer)

    def reset(self):
        for metric in self._each():
            metric.reset()

    def update(self, loss, output, meta):
        for metric in self._each():
            metric.update(loss, output, meta)

    def avg(self):
        return {metric.name: metric.avg() for metric in self._each()}

    def result(self):
        merged = {}
        for metric in self._each():
            merged.update(metric.result())
        return merged


class BaseMetric:
    """model/metric.py:45-65: name, writer, the two phase flags, and the four methods a metric has to
    provide."""

    def __init__(self, name):
        self.name, self.writer = name, None
        self.is_train = self.is_val = True

    def set_writer(self, writer):
        self.writer = writer

    def _abstract(self, what):
        raise NotImplementedError(f"{type(self).__name__}.{what}")

    def reset(self):
        self._abstract("reset")

    def update(self, loss, output, meta):
        self._abstract("update")

    def avg(self):
        self._abstract("avg")

    def result(self):
        self._abstract("result")


class ScalarPerBatchMetric(BaseMetric):
    """model/metric.py:68-95: running total / count / average of a per-batch scalar.  Host glue
    (the trainer feeds it `loss.item()`); the reference keeps the three numbers in a one-row
    DataFrame and zeroes it through `.values[:] = 0`, which pandas >= 3 rejects (copy-on-write), so
    they are plain floats here -- same values, same `avg()` / `result()`."""

    def __init__(self, name, metric_fun):
        super().__init__(name)
        self.fun = metric_fun
        self.reset()

    def reset(self):
        self.total = 0
        self.counts = 0
        self.average = 0

    def update(self, loss, output, meta, n=1):
        value = self.fun(loss, output, meta)
        if self.writer is not None:
            self.writer.add_scalar(self.name, value)
        self.total += value * n
        self.counts += n
        self.average = self.total / self.counts

    def avg(self):
        return self.average

    def result(self):
        return {self.name: self.average}


class LossMetric(ScalarPerBatchMetric):
    """model/metric.py:98-100."""

    def __init__(self):
        super().__init__("loss", lambda loss, o, m: loss)


def _to_device(x: ArrayLike, device: torch.device) -> torch.Tensor:
    """numpy / CPU tensor -> CUDA fp32 (or bf16) tensor via pinned staging; CUDA tensors pass."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"expected numpy array or tensor, got {type(x)}")
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    if x.is_cuda:
        return x
    if not x.is_pinned():
        x = x.contiguous().pin_memory()
    return x.to(device, non_blocking=True)


_copy_streams: Dict[int, "torch.cuda.Stream"] = {}


def _copy_stream(device: torch.device) -> "torch.cuda.Stream":
    """One persistent H2D staging stream per device (the caching allocator pools blocks per
    stream, so a fresh stream per call would cudaMalloc every time)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _copy_streams.get(idx)
    if st is None:
        st = torch.cuda.Stream(device)
        _copy_streams[idx] = st
    return st


def _default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise ops.VtcError("RecallAtK needs a CUDA device: vtc_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class RecallAtK(BaseMetric):
    def __init__(self, name_a, name_b, k_vals=5, precision: str = "exact", metric: str = "l2"):
        super().__init__("recall@k")
        if not isinstance(k_vals, collections.abc.Iterable):
            k_vals = [k_vals]
        self.k_vals = list(k_vals)
        self.name_a = name_a
        self.name_b = name_b
        self.is_train = False
        self.precision = precision  # "exact" (fp32 inputs ranked exactly) | "bf16" | "brute"
        self.metric = metric        # "l2" = what faiss.GpuIndexFlatL2 ranks by
        self.insert_index = 0
        self.features_a_list: List[torch.Tensor] = []
        self.features_b_list: List[torch.Tensor] = []

    def reset(self):
        self.insert_index = 0
        self.features_a_list = []
        self.features_b_list = []

    def update(self, loss, output, meta):
        fa = output[0]
        fb = output[1]
        batch_size = fa.shape[0]
        end = self.insert_index + batch_size
        # the reference moves every batch to the CPU here (model/metric.py:129-130); the features
        # stay on the device and are consumed there by result()
        self.features_a_list.append(fa.detach())
        self.features_b_list.append(fb.detach())
        self.insert_index = end

    # ------------------------------------------------------------------ the hot path
    def compute_full(self, features_a: ArrayLike, features_b: ArrayLike,
                     device: Optional[torch.device] = None) -> Dict[str, object]:
        """gallery = features_a, queries = features_b, gt(t) = t (model/metric.py:137-161).

        Returns device tensors: rank0 int32 [N], hits int64 [nk], medr fp64 [1], plus sizes."""
        if getattr(features_b, "ndim", 2) != 2 or getattr(features_a, "ndim", 2) != 2:
            raise ValueError(
                "RecallAtK.compute needs 2-D [N, D] features: one text per video "
                "(multi-caption 3-D inputs are not supported, SURVEY.md App. B #8)")
        if device is None:
            for x in (features_a, features_b):
                if isinstance(x, torch.Tensor) and x.is_cuda:
                    device = x.device
            device = device or _default_device()
        num_samples = features_a.shape[0]
        if features_b.shape[0] != num_samples:
            # gt(t) = t and the denominator is the gallery size (model/metric.py:138,154-158)
            raise AssertionError(
                f"RecallAtK assumes len(a) == len(b) (got {num_samples} vs {features_b.shape[0]})")
        host_b = not (isinstance(features_b, torch.Tensor) and features_b.is_cuda)
        if host_b and num_samples * features_b.shape[1] * 4 >= self.PIPELINE_MIN_BYTES:
            return self._compute_full_pipelined(features_a, features_b, device)
        a = _to_device(features_a, device)
        b = _to_device(features_b, device)
        if a.dtype != b.dtype:
            a, b = a.float(), b.float()
        # one library call: row prologue + tensor-core pass + epilogue chain (re-check, hit
        # counts, median)
        full = ops.rank_eval(b, a, self.k_vals, metric=self.metric, precision=self.precision)
        return {"rank0": full["rank0"], "hits": full["hits"], "medr": full["medr"],
                "num_samples": num_samples}

    # Host inputs of at least this many bytes are staged in chunks on a copy stream so that the
    # host->device transfer overlaps the ranking.  When BOTH sides come from the host, gallery and
    # queries are cut with the same bounds and copied interleaved (G0, Q0, G1, Q1, ...); when pair i
    # lands, the new query rows are ranked against all gallery rows so far (gt(t) = t puts their
    # ground truth among them) and the earlier query rows against the new gallery rows -- rank
    # counts are additive over gallery chunks -- so the tensor cores start after 1/c of the
    # transfer instead of after the whole gallery, with 2c - 1 library calls.
    PIPELINE_MIN_BYTES = 32 << 20
    PIPELINE_CHUNKS = 4      # query chunks against a device-resident gallery
    # Chunk pairs when both sides are staged from the host.  Once the last pair has landed, the
    # blocks it enables -- (2 r_last N - r_last^2) of the N^2 pairs for r_last rows -- are all that is
    # left, so the evaluation cannot end before T + that share of the ranking time W: equal chunks
    # give (2c - 1) / c^2 (c = 6: 0.31 W, c = 12: 0.16 W), and every call is memset + 4 launches
    # (rank_stage.cu), so a dozen pairs are affordable.
    PIPELINE_CHUNKS_2D = 10

    @staticmethod
    def _pipeline_bounds_2d(n: int, c: int) -> List[int]:
        """Row bounds of the c interleaved chunk pairs (equal chunks)."""
        return [n * i // c for i in range(c + 1)]

    def _compute_full_pipelined(self, features_a: ArrayLike, features_b: ArrayLike,
                                device: torch.device) -> Dict[str, object]:
        def host(x):
            if isinstance(x, np.ndarray):
                x = torch.from_numpy(np.ascontiguousarray(x))
            if x.dtype not in (torch.float32, torch.bfloat16):
                x = x.float()
            x = x.contiguous()
            return x if x.is_pinned() else x.pin_memory()

        a = _to_device(features_a, device) if (isinstance(features_a, torch.Tensor)
                                              and features_a.is_cuda) else None
        ha = None if a is not None else host(features_a)
        hb = host(features_b)
        n = hb.shape[0]
        dtype = hb.dtype if (ha is None or ha.dtype == hb.dtype) else torch.float32
        if a is not None and a.dtype != hb.dtype:
            dtype = torch.float32
        main = torch.cuda.current_stream(device)
        copy = _copy_stream(device)
        rank0 = torch.empty(n, dtype=torch.int32, device=device)
        gt_score = torch.empty(n, dtype=torch.float64, device=device)

        def stage(src, s, e):
            t = src[s:e].to(device, non_blocking=True).to(dtype)
            ev = torch.cuda.Event()
            ev.record(copy)
            return t, ev

        if a is not None:
            # gallery already on the device: stream the query chunks against all of it
            c = self.PIPELINE_CHUNKS
            bounds = [n * i // c for i in range(c + 1)]
            with torch.cuda.stream(copy):
                copy.wait_stream(main)
                staged = [stage(hb, s, e) for s, e in zip(bounds[:-1], bounds[1:])]
            a = a.to(dtype)
            for (s, e), (qc, ev) in zip(zip(bounds[:-1], bounds[1:]), staged):
                if e == s:
                    continue
                main.wait_event(ev)
                qc.record_stream(main)
                r, g = ops.sim_rank(qc, a, row_offset=s, metric=self.metric, precision=self.precision)
                rank0[s:e] = r
                gt_score[s:e] = g
            hits, medr = ops.rank_finalize(rank0, gt_score, a.shape[0], self.k_vals)
            return {"rank0": rank0, "hits": hits, "medr": medr, "num_samples": a.shape[0]}

        c = self.PIPELINE_CHUNKS_2D
        bounds = self._pipeline_bounds_2d(n, c)
        dq = torch.empty((n, hb.shape[1]), dtype=dtype, device=device)
        dg = torch.empty((n, ha.shape[1]), dtype=dtype, device=device)
        # the bf16 mode ranks the RN-even bf16 roundings of the inputs: round each chunk on the copy
        # stream as it lands (identical results); bf16 rows of whole swizzle atoms are used by the
        # library as tensor-core operands in place, so no call re-converts the rows it is handed
        to16 = self.precision == "bf16" and dtype == torch.float32
        sq, sg = dq, dg
        if to16:
            dq = torch.empty_like(sq, dtype=torch.bfloat16)
            dg = torch.empty_like(sg, dtype=torch.bfloat16)
        events = []
        with torch.cuda.stream(copy):
            copy.wait_stream(main)
            for s, e in zip(bounds[:-1], bounds[1:]):
                sg[s:e].copy_(ha[s:e], non_blocking=True)
                sq[s:e].copy_(hb[s:e], non_blocking=True)
                if to16:
                    dg[s:e].copy_(sg[s:e])
                    dq[s:e].copy_(sq[s:e])
                ev = torch.cuda.Event()
                ev.record(copy)
                events.append(ev)
        # per-row quantities are computed once, by the first call that sees the rows, and handed to
        # every later call that touches them (vtc_sim_rank_prepared): canonical ||x||^2 per gallery
        # row, the ||q||^2 bound and d(t,gt) per query row -- no call re-walks rows it was given
        cached = self.precision in ("bf16", "exact")
        sq64 = torch.empty(n, dtype=torch.float64, device=device) if cached else None
        qq = torch.empty(n, dtype=torch.float32, device=device) if cached else None
        for (s, e), ev in zip(zip(bounds[:-1], bounds[1:]), events):
            if e == s:
                continue
            main.wait_event(ev)
            if cached:
                # the earlier rows against the new gallery rows (computes the new rows' norms) ...
                if s > 0:
                    ops.sim_rank(dq[:s], dg[s:e], row_offset=0, col_offset=s, metric=self.metric,
                                 precision=self.precision, gt_score=gt_score[:s], rank0=rank0[:s],
                                 accumulate=True, sq64_out=sq64[s:e], qq=qq[:s])
                # ... and the new rows against everything that has arrived (their ground truth is
                # among it: computes their d(t,gt) and norm bounds)
                ops.sim_rank(dq[s:e], dg[:e], row_offset=s, metric=self.metric,
                             precision=self.precision, gt_score_out=gt_score[s:e], rank0=rank0[s:e],
                             accumulate=False, qq_out=qq[s:e],
                             **({"sq64": sq64[:e]} if s > 0 else {"sq64_out": sq64[:e]}))
                continue
            _, g = ops.sim_rank(dq[s:e], dg[:e], row_offset=s, metric=self.metric,
                                precision=self.precision, rank0=rank0[s:e], accumulate=False)
            gt_score[s:e] = g
            if s > 0:
                ops.sim_rank(dq[:s], dg[s:e], row_offset=0, col_offset=s, metric=self.metric,
                             precision=self.precision, gt_score=gt_score[:s], rank0=rank0[:s],
                             accumulate=True)
        for t in (dq, dg, sq, sg):
            t.record_stream(copy)
        hits, medr = ops.rank_finalize(rank0, gt_score, n, self.k_vals)
        return {"rank0": rank0, "hits": hits, "medr": medr, "num_samples": n}

    def compute(self, features_a: ArrayLike, features_b: ArrayLike) -> List[Tuple[int, float]]:
        full = self.compute_full(features_a, features_b)
        hits = full["hits"].cpu().numpy()  # the one device->host read of the result
        num_samples = full["num_samples"]
        return [(k, float(h) / num_samples) for k, h in zip(self.k_vals, hits)]

    def avg(self):
        return None

    def result(self):
        started = time.time()
        print("RecallAtK: result()...", end=" ", flush=True)
        side_a = (self.name_a, torch.cat(self.features_a_list))
        side_b = (self.name_b, torch.cat(self.features_b_list))
        assert self.insert_index == len(side_a[1])
        # both retrieval directions, keyed and ordered as model/metric.py:175-179: queries b against
        # gallery a first ("{b}_from_{a}-recall_at_{k}"), then the roles swapped
        res = {}
        for (g_name, gallery), (q_name, queries) in ((side_a, side_b), (side_b, side_a)):
            for k, recall in self.compute(gallery, queries):
                res[f"{q_name}_from_{g_name}-recall_at_{k}"] = recall
        if self.writer:
            for key, recall in res.items():
                self.writer.add_scalar(key, recall)
        print("RecallAtK: result() took %.3fs" % (time.time() - started))
        return res
