#!/usr/bin/env python
"""bench.py -- the headline benchmark of the contrastive-retrieval hot path.

Metric (BASELINE.json): query-gallery pairs/s for fused similarity + rank + R@K at 100k x 100k,
512-d, on 1/2/4/8 B200.  A "step" is one pass of the hot path over one batch of synthetic
embeddings: text->video retrieval eval of N queries against M gallery rows (vtc_sim_rank +
vtc_rank_finalize), 1e10 pairs per step at the default size.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by the driver as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
one rank per GPU over NCCL: query rows and gallery rows are sharded, gallery shards are
all-gathered over NVLink, ranks are additive over gallery chunks, the int32 ranks are gathered for
the hit counts and the median (vtc_b200/parallel.py); the whole sharded step is captured once and
replayed as a CUDA graph (--no-graph: kernel by kernel).  Total work is fixed as N grows =>
"scaling": "strong".

One JSON line is printed by rank 0 (see the contract in the task description): `value` is the
device-resident throughput, `e2e` the same metric through the reference-facing call
`RecallAtK.compute` with pinned HOST buffers (H2D + D2H inside the timed region), `roofline` the
tensor-core kernel against the measured bf16 peak, `cpu_baseline` the oracle port timed on this
box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "query-gallery pairs/s (sim+rank+R@K)"
UNIT = "pairs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=100_000, help="queries")
    ap.add_argument("--m", type=int, default=100_000, help="gallery rows")
    ap.add_argument("--d", "--dim", dest="d", type=int, default=512,
                    help="embedding width (use --dim under torchrun: its parser rejects --d as ambiguous)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "exact", "brute"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true",
                    help="N > 1: launch the sharded step kernel by kernel instead of as a CUDA graph")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU-baseline time")
    return ap.parse_args()


def workload_name(a):
    return (f"retrieval_eval_{a.n // 1000}kx{a.m // 1000}k_{a.d}d_text2video_sim+rank+R@1/5/10+MedR")


# --------------------------------------------------------------------------------- CPU baseline
def cpu_reference_pairs_per_s(T, V, target_seconds: float, repeats: int = 1):
    """The reference's algorithm for this path on the host CPU (oracle port): exact fp32 L2 search
    depth max(k)+1 (faiss stand-in: sgemm + top-k) followed by the reference's Python hit loop
    (model/metric.py:137-161), all host threads, on a bounded sample of the query rows."""
    import torch

    from oracle import vtc_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    Vn = V.numpy()
    probe = min(1024, T.shape[0])
    t0 = time.perf_counter()
    O.recall_at_k(Vn, T[:probe].numpy(), [1, 5, 10], fast=True)
    dt = time.perf_counter() - t0
    rate = probe / max(dt, 1e-6)
    rows = int(max(probe, min(T.shape[0], rate * target_seconds)))
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        rec = O.recall_at_k(Vn, T[:rows].numpy(), [1, 5, 10], fast=True)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    pairs = rows * V.shape[0]
    return {
        "value": pairs / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
        "sample": f"{rows} of {T.shape[0]} query rows x full {V.shape[0]}-row gallery, "
                  f"{best:.2f} s, torch CPU fp32 sgemm + topk(11) + the reference's Python hit loop",
        "seconds": best, "recall_sample": [r for _, r in rec],
    }


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; faiss and
    clip are absent) timed on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from vtc_b200.synthetic import make_retrieval_pair

    T, V = make_retrieval_pair(a.n, a.m, a.d, seed=1023)
    per_step = max(2.0, min(a.cpu_seconds, 120.0 / max(1, a.steps + a.warmup)))
    for _ in range(a.warmup):
        cpu_reference_pairs_per_s(T, V, per_step / 4)
    vals, last = [], None
    t_total = 0.0
    for _ in range(a.steps):
        last = cpu_reference_pairs_per_s(T, V, per_step)
        vals.append(last["value"])
        t_total += last["seconds"]
    v = statistics.median(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t_total / max(1, a.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (seed 1023, sigma=6 noisy-copy queries, SURVEY.md 8d)",
        "config": {"workload": workload_name(a), "N": a.n, "M": a.m, "D": a.d,
                   "note": "each step is a bounded sample of the workload (see cpu_baseline.sample)"},
        "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["cpu_baseline"]["value"] = v
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.lines = []   # (arrival time, csv line)
        self.proc = None
        self.t0 = None    # the timed region, in time.perf_counter() terms
        self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                 "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        return self.summarise()

    def summarise(self):
        """Samples that arrived inside the timed region; the sampler is started before the warm-up
        steps (nvidia-smi needs ~0.1 s to print its first line, the timed region of a 20-step run
        is 0.16 s), so if none fell inside, the samples taken under load since then are used."""
        lines = list(self.lines)
        inside = [ln for t, ln in lines
                  if self.t0 is not None and t >= self.t0 and (self.t1 is None or t <= self.t1 + 0.03)]
        window = "timed region"
        if not inside:
            inside = [ln for _, ln in lines]
            window = "warm-up + timed region (no sample arrived inside the timed region)"
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ our arm
def run_ours(a):
    import torch
    import torch.distributed as dist

    from vtc_b200 import _ffi, ops
    from vtc_b200.model.metric import RecallAtK
    from vtc_b200.parallel import GraphedRankEval, shard_bounds, sharded_rank_eval
    from vtc_b200.synthetic import make_retrieval_pair

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (vtc_b200 has no CPU path); "
                         "use --impl reference for the CPU baseline")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _ffi.load()

    T, V = make_retrieval_pair(a.n, a.m, a.d, seed=1023)
    k_vals = [1, 5, 10]
    try:  # opt-in kernel variant; the library reads the same variable with atoi()
        rank_fold = int(os.environ.get("VTC_RANK_FOLD", "0") or 0) != 0
    except ValueError:
        rank_fold = False
    qs, qe = shard_bounds(a.n, world, rank)
    gs, ge = shard_bounds(a.m, world, rank)
    q_local = T[qs:qe].contiguous().to(dev)
    g_local = V[gs:ge].contiguous().to(dev)

    # N > 1: ~1 ms of GPU work per rank behind ~40 launches and 3 collectives is launch-bound from
    # Python, so the whole sharded step is captured once and replayed as one CUDA graph
    graphed = None
    if world > 1 and not a.no_graph and not os.environ.get("VTC_PHASE_TIMING"):
        try:
            graphed = GraphedRankEval(q_local, g_local, a.n, a.m, k_vals, "l2", a.precision)
        except Exception as exc:  # noqa: BLE001  (capture refused: measure kernel by kernel, say so)
            print(f"[bench] CUDA-graph capture of the sharded step failed on rank {rank}: {exc!r}; "
                  "falling back to per-kernel launches", file=sys.stderr, flush=True)
            graphed = None
            torch.cuda.synchronize()
        # every rank must take the same path: the collectives inside differ otherwise
        ok = torch.tensor([1 if graphed is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0 and graphed is not None:
            graphed.close()
            graphed = None

    def eager_step():
        return sharded_rank_eval(q_local, g_local, a.n, a.m, k_vals, "l2", a.precision)

    def step():
        if world == 1:
            rank0, gts = ops.sim_rank(q_local, g_local, metric="l2", precision=a.precision)
            hits, medr = ops.rank_finalize(rank0, gts, a.m, k_vals)
            return hits, medr
        if graphed is not None:
            res = graphed()
            return res["hits"], res["medr"]
        res = eager_step()
        if res.get("phases_ms") and rank == 0:
            print("phases_ms", json.dumps(res["phases_ms"]), file=sys.stderr)
        return res["hits"], res["medr"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # before the warm-up: nvidia-smi takes ~0.1 s to deliver its first sample
    for _ in range(max(a.warmup, 3)):
        hits, medr = step()
    barrier()
    _ffi.kernel_timer_enable(True)
    _ffi.kernel_timer_read()
    launches0 = _ffi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_start()
    ev0.record()
    t_host = time.perf_counter()
    for _ in range(a.steps):
        hits, medr = step()
    host_ms_per_step = (time.perf_counter() - t_host) / a.steps * 1e3  # CPU time to enqueue a step
    ev1.record()
    barrier()
    sampler.mark_end()
    launches = _ffi.launch_count() - launches0
    tc_ms, tc_n = _ffi.kernel_timer_read()
    tc_steps = a.steps
    if graphed is not None:
        # a replayed graph has no per-kernel events: count the captured launches, and time the
        # dominant kernel on a few eager steps OUTSIDE the timed region (roofline only)
        launches = graphed.launches_per_replay * a.steps
        tc_steps = 3
        for _ in range(tc_steps):
            eager_step()
        torch.cuda.synchronize()
        tc_ms, tc_n = _ffi.kernel_timer_read()
    _ffi.kernel_timer_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    tl = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
    ms_per_step = ms.item() / a.steps
    pairs = float(a.n) * float(a.m)
    value = pairs / (ms_per_step * 1e-3)

    # ---- e2e: the reference-facing call with pinned host buffers (H2D + D2H inside the timing)
    e2e = None
    if not a.no_e2e and world == 1:
        Tp, Vp = T.pin_memory(), V.pin_memory()
        metric = RecallAtK("videos", "titles", k_vals, precision=a.precision)
        for _ in range(2):
            metric.compute(Vp, Tp)
        torch.cuda.synchronize()
        n_e2e = max(3, min(a.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            rec = metric.compute(Vp, Tp)  # returns host floats: includes the D2H read
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        e2e = {"value": pairs / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(Tp.numel() * 4 + Vp.numel() * 4),
               "d2h_bytes_per_step": 8 * len(k_vals), "ms_per_step": dt * 1e3,
               "api": "vtc_b200.model.metric.RecallAtK.compute(pinned fp32 host tensors)",
               "recall": [r for _, r in rec]}
    elif not a.no_e2e:
        # N > 1: each rank stages ITS shards from pinned host memory, then the sharded eval
        Tq, Vg = T[qs:qe].contiguous().pin_memory(), V[gs:ge].contiguous().pin_memory()
        def e2e_step():
            if graphed is not None:
                # H2D straight into the captured step's static input buffers, one graph launch
                return graphed(Tq, Vg)["hits"].cpu()
            ql = Tq.to(dev, non_blocking=True)
            gl = Vg.to(dev, non_blocking=True)
            res = sharded_rank_eval(ql, gl, a.n, a.m, k_vals, "l2", a.precision)
            return res["hits"].cpu()
        for _ in range(2):
            e2e_step()
        barrier()
        n_e2e = max(3, min(a.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": pairs / dt.item(), "unit": UNIT,
               "h2d_bytes_per_step": int(a.n * a.d * 4 + a.m * a.d * 4),
               "d2h_bytes_per_step": 8 * len(k_vals) * world, "ms_per_step": dt.item() * 1e3,
               "api": ("vtc_b200.parallel.GraphedRankEval(pinned fp32 host shards)" if graphed is not None
                       else "vtc_b200.parallel.sharded_rank_eval(pinned fp32 host shards)")}

    if graphed is not None:
        graphed.close()  # NCCL will not tear a communicator down under a live captured graph
    if rank != 0:
        if world > 1:
            _teardown(dist)
        return

    # ---- roofline of the dominant kernel (the tcgen05 similarity + rank kernel)
    peaks = {}
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        peaks = json.load(open(ppath))
    peak_tf = peaks.get("bf16_tflops")
    peak_src = "measured burst (MEASURED_PEAKS.json bf16_tflops)"
    if peak_tf is None:
        peak_tf, peak_src = 1590.0, "fallback (B200_PROFILING.md)"
    roofline = None
    if tc_n > 0 and a.precision != "brute":
        k_eff = a.d if a.precision == "bf16" else 3 * a.d
        if rank_fold:
            k_eff += 16  # the fold block: one more K16 MMA step per tile
        n_rows = qe - qs
        flops_alg = 2.0 * n_rows * a.m * a.d            # algorithmic FLOPs of this rank's launches
        launches_per_step = tc_n / tc_steps
        ms_per_launch = tc_ms / tc_n
        achieved = flops_alg / launches_per_step / (ms_per_launch * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"{a.precision}_{a.n}x{a.m}x{a.d}")
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved / peak_tf, "traffic": traffic,
                    "kernel": "vtc::tc::sim_tc_kernel<RankEpi>", "peak_source": peak_src,
                    "frac_vs_sustained": (achieved / peaks["bf16_tflops_sustained"]
                                          if peaks.get("bf16_tflops_sustained") else None),
                    "ms_per_launch": ms_per_launch, "launches_per_step": launches_per_step,
                    "kernel_share_of_step": tc_ms / tc_steps / ms_per_step,
                    "issued_tflops": achieved * k_eff / a.d,
                    "note": "achieved = algorithmic 2*N*M*D per launch / CUDA-event launch time; "
                            "the exact mode issues 3x the MMAs (bf16x3 split), see issued_tflops"}

    cpu = None
    if not a.no_cpu_baseline:
        cpu = cpu_reference_pairs_per_s(T, V, a.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if a.precision == "bf16" else ("f32 (bf16x3 tensor-core split + fp64 recheck)"
                                                       if a.precision == "exact" else "f64"),
        "data": "synthetic (seed 1023, unit-norm gallery, sigma=6 noisy-copy queries, SURVEY.md 8d)",
        "config": {"workload": workload_name(a), "N": a.n, "M": a.m, "D": a.d,
                   "precision": a.precision, "metric": "l2", "k_vals": k_vals,
                   "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU",
                   "cuda_graph": graphed is not None,
                   "rank_epilogue": "fold (VTC_RANK_FOLD=1, opt-in)" if rank_fold else "default",
                   "sharded_step": ("n/a" if world == 1 else
                                    "single pass (VTC_SHARD_SINGLE_PASS=1, opt-in)"
                                    if os.environ.get("VTC_SHARD_SINGLE_PASS", "0") not in ("", "0")
                                    else "local chunk overlapped with the gather + remote ranges"),
                   "l2_flush": "not needed: inputs + operands (>= 600 MB) exceed the 126 MB L2"},
        "hits": [int(x) for x in hits.cpu().tolist()], "medr": float(medr.cpu()[0]),
        "e2e": e2e, "gpu_launches": int(tl.item()), "host_enqueue_ms_per_step": host_ms_per_step,
        "clocks": clocks,
        "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        _teardown(dist)


def _teardown(dist):
    """destroy_process_group with a bounded wait: the measurement is already printed, a stuck
    communicator teardown must not hold the launcher."""
    t = threading.Timer(30.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
