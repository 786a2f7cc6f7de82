#!/usr/bin/env python
"""bench.py -- the headline benchmark of the contrastive-retrieval hot path.

Metric (BASELINE.json): query-gallery pairs/s for fused similarity + rank + R@K at 100k x 100k,
512-d, on 1/2/4/8 B200.  A "step" is one pass of the hot path over one batch of synthetic
embeddings: text->video retrieval eval of N queries against M gallery rows (similarity + rank of
ground truth + R@1/5/10 + MedR), 1e10 pairs per step at the default size.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by the driver as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
one rank per GPU over NCCL: query rows and gallery rows are sharded, gallery shards (and their
owners' norms) are all-gathered over NVLink, ranks are additive over gallery chunks, the int32 ranks
are gathered for the hit counts and the median (vtc_b200/parallel.py); the whole sharded step is
captured once and replayed as a CUDA graph (--no-graph: kernel by kernel).  Total work is fixed as
N grows => "scaling": "strong".

One JSON line is printed by rank 0 (see the contract in the task description):
  value / ms_per_step   the bf16 tensor-core mode, inputs resident in HBM.  K steps are timed as a
                        block (barrier + synchronize on both sides, CUDA events, max over ranks);
                        the block is repeated until >= 0.6 s have been timed so that nvidia-smi clock
                        samples land inside, and the MEDIAN block is reported (`ms_per_step_blocks`)
  e2e                   the same metric through the reference-facing call RecallAtK.compute with
                        pinned fp32 HOST tensors (H2D + D2H inside the timed region)
  roofline              the tcgen05 similarity + rank kernel against the measured bf16 peak
  exact                 the same step in the reference-identical mode (fp32 inputs, bit-exact ranks:
                        bf16x3 tensor-core split + fp64 re-check), with its own e2e and roofline
  parity_check          a slice of query rows re-ranked on the device by the fp64 brute-force kernel
                        (canonical arithmetic) and compared with both modes' ranks: "ok" | "mismatch"
  c4_d768, c5_topk      BASELINE configs 4 and 5 at this N (100k x 100k x 768 rank eval; 10k x 1M x
                        512 streaming top-k, k = 11)
  cpu_baseline          (N = 1) the oracle port of the reference algorithm timed on this box's host
                        cores on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import math
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
K_VALS = [1, 5, 10]
DATA = "synthetic (seed 1023, unit-norm gallery, sigma=6 noisy-copy queries, SURVEY.md 8d)"
MIN_TIMED_SECONDS = 0.6


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
    ap.add_argument("--precision", default="bf16", choices=["bf16", "exact", "brute"],
                    help="mode of the headline value (the other of bf16 / exact is reported nested)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the nested records (other precision, parity check, configs 4 and 5)")
    ap.add_argument("--no-graph", action="store_true",
                    help="N > 1: launch the sharded step kernel by kernel instead of as a CUDA graph")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU-baseline time")
    return ap.parse_args()


def workload_name(n, m, d):
    return f"retrieval_eval_{n // 1000}kx{m // 1000}k_{d}d_text2video_sim+rank+R@1/5/10+MedR"


def workload_config(a):
    """The keys that define the WORKLOAD -- identical in both arms (ours / reference)."""
    return {"workload": workload_name(a.n, a.m, a.d), "N": a.n, "M": a.m, "D": a.d, "metric": "l2",
            "k_vals": K_VALS,
            "l2_flush": "not needed: inputs + operands (>= 600 MB) exceed the 126 MB L2"}


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
    O.recall_at_k(Vn, T[:probe].numpy(), K_VALS, fast=True)
    dt = time.perf_counter() - t0
    rate = probe / max(dt, 1e-6)
    rows = int(max(probe, min(T.shape[0], rate * target_seconds)))
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        rec = O.recall_at_k(Vn, T[:rows].numpy(), K_VALS, fast=True)
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
    clip are absent) timed on this box's host cores.  Same `config` as our arm; each
    step is a bounded sample of the workload (see cpu_baseline.sample)."""
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
        "data": DATA, "config": workload_config(a),
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
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,"
         "enforced.power.limit")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.lines = []   # (arrival time, csv line)
        self.proc = None
        self.windows = []  # timed regions, in time.perf_counter() terms

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

    def window(self, t0, t1):
        self.windows.append((t0, t1))

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
        """Samples that arrived inside a timed block (the blocks add up to >= 0.6 s, the sampler
        period is 25 ms)."""
        lines = list(self.lines)
        inside = [ln for t, ln in lines if any(t0 <= t <= t1 + 0.03 for t0, t1 in self.windows)]
        window = "timed blocks"
        if not inside:
            inside = [ln for _, ln in lines]
            window = "warm-up + timed blocks (no sample arrived inside a timed block)"
        sm, mx, power, limit, reasons = [], [], [], [], set()
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
            try:
                limit.append(float(f[9]))
            except (IndexError, ValueError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None,
                "power_w_median": statistics.median(power) if power else None,
                "power_limit_w": max(limit) if limit else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ our arm
class Ctx:
    """Process-wide state of one bench run (one rank)."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist

        self.a = a
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (vtc_b200 has no CPU path); "
                             "use --impl reference for the CPU baseline")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.sampler = ClockSampler(self.local_rank) if self.rank == 0 else None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()

    def sum_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.item()

    def time_blocks(self, step, steps, warmup, sample_clocks=False, min_seconds=MIN_TIMED_SECONDS):
        """`warmup` untimed steps, then blocks of exactly `steps` steps, each bracketed by a barrier +
        synchronize on both sides and timed with CUDA events (max over ranks), repeated until
        `min_seconds` have been timed.  Returns (median ms/step, [ms/step per block], last result,
        host enqueue ms/step)."""
        torch = self.torch
        out = None
        for _ in range(warmup):
            out = step()
        self.barrier()
        per_block, host_ms = [], []
        total = 0.0
        while True:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.barrier()
            t0 = time.perf_counter()
            ev0.record()
            for _ in range(steps):
                out = step()
            host_ms.append((time.perf_counter() - t0) / steps * 1e3)
            ev1.record()
            self.barrier()
            t1 = time.perf_counter()
            if sample_clocks and self.sampler is not None:
                self.sampler.window(t0, t1)
            ms = self.max_over_ranks(ev0.elapsed_time(ev1))
            per_block.append(ms / steps)
            total += ms * 1e-3
            if total >= min_seconds or len(per_block) >= 64:
                break
        return statistics.median(per_block), per_block, out, statistics.median(host_ms)


def peak_tflops(timed_seconds: float = 0.0):
    """The roofline denominator: cuBLAS bf16 as measured on this pool by the driver
    (MEASURED_PEAKS.json) -- the SUSTAINED figure when the kernel was timed inside a long loop (the
    timed blocks add up to >= 0.5 s, i.e. under the same power cap the 4 s cuBLAS loop saw), the burst
    figure for a short one; the other one is reported next to it."""
    peaks = {}
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        peaks = json.load(open(ppath))
    burst, sust = peaks.get("bf16_tflops"), peaks.get("bf16_tflops_sustained")
    if burst is None:
        return 1590.0, "fallback burst (B200_PROFILING.md)", 1590.0, 1400.0
    if timed_seconds >= 0.5 and sust is not None:
        return sust, ("measured sustained (MEASURED_PEAKS.json bf16_tflops_sustained; the kernel was "
                      f"timed inside a {timed_seconds:.2f} s loop)"), burst, sust
    return burst, "measured burst (MEASURED_PEAKS.json bf16_tflops)", burst, sust


def rank_roofline(tc_ms, tc_n, tc_steps, ms_per_step, n_rows, m, d, precision, timed_seconds=0.0):
    if tc_n <= 0 or precision == "brute":
        return None
    peak_tf, peak_src, burst, sust = peak_tflops(timed_seconds)
    k_eff = d if precision == "bf16" else 3 * d
    flops_alg = 2.0 * n_rows * m * d            # algorithmic FLOPs of this rank's launches per step
    launches_per_step = tc_n / tc_steps
    ms_per_launch = tc_ms / tc_n
    achieved = flops_alg / launches_per_step / (ms_per_launch * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(f"{precision}_{n_rows}x{m}x{d}")
    return {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved / peak_tf, "traffic": traffic,
            "kernel": "vtc::tc::sim_tc_kernel<RankEpi>", "peak_source": peak_src,
            "frac_vs_burst": achieved / burst if burst else None,
            "frac_vs_sustained": achieved / sust if sust else None,
            "ms_per_launch": ms_per_launch, "launches_per_step": launches_per_step,
            "kernel_share_of_step": tc_ms / tc_steps / ms_per_step,
            "issued_tflops": achieved * k_eff / d,
            "note": "achieved = algorithmic 2*N*M*D per launch / CUDA-event launch time; "
                    "the exact mode issues 3x the MMAs (bf16x3 split), so its ceiling is 1/3 of the "
                    "bf16 peak in algorithmic FLOPs (see issued_tflops)"}


def measure_rank(cx: Ctx, q_local, g_local, n, m, d, precision, steps, warmup, sample_clocks,
                 use_graph, want_state=False):
    """Device-resident retrieval evaluation (the `value` of a record) at this world size."""
    torch = cx.torch
    from vtc_b200 import _ffi, ops
    from vtc_b200.parallel import GraphedRankEval, sharded_rank_eval

    graphed = None
    if cx.world > 1 and use_graph and not os.environ.get("VTC_PHASE_TIMING"):
        try:
            graphed = GraphedRankEval(q_local, g_local, n, m, K_VALS, "l2", precision)
        except Exception as exc:  # noqa: BLE001  (capture refused: measure kernel by kernel, say so)
            print(f"[bench] CUDA-graph capture of the sharded step failed on rank {cx.rank}: {exc!r}; "
                  "falling back to per-kernel launches", file=sys.stderr, flush=True)
            graphed = None
            torch.cuda.synchronize()
        # every rank must take the same path: the collectives inside differ otherwise
        ok = torch.tensor([1 if graphed is not None else 0], device=cx.dev)
        cx.dist.all_reduce(ok, op=cx.dist.ReduceOp.MIN)
        if ok.item() == 0 and graphed is not None:
            graphed.close()
            graphed = None

    def eager_step():
        return sharded_rank_eval(q_local, g_local, n, m, K_VALS, "l2", precision)

    def step():
        if cx.world == 1:
            # one library call: prologue + tensor-core pass + cooperative epilogue (+ memset)
            return ops.rank_eval(q_local, g_local, K_VALS, metric="l2", precision=precision)
        if graphed is not None:
            return graphed()
        res = eager_step()
        if res.get("phases_ms") and cx.rank == 0:
            print("phases_ms", json.dumps(res["phases_ms"]), file=sys.stderr)
        return res

    for _ in range(warmup):
        step()
    cx.barrier()
    _ffi.kernel_timer_enable(True)
    _ffi.kernel_timer_read()
    launches0 = _ffi.launch_count()
    ms_per_step, blocks, res, host_ms = cx.time_blocks(step, steps, 0, sample_clocks)
    launches = (_ffi.launch_count() - launches0) / (steps * len(blocks))
    tc_ms, tc_n = _ffi.kernel_timer_read()
    tc_steps = steps * len(blocks)
    if graphed is not None:
        # a replayed graph has no per-kernel events: count the captured launches, and time the
        # dominant kernel on a few eager steps OUTSIDE the timed region (roofline only)
        launches = graphed.launches_per_replay
        tc_steps = 3
        for _ in range(tc_steps):
            eager_step()
        torch.cuda.synchronize()
        tc_ms, tc_n = _ffi.kernel_timer_read()
    _ffi.kernel_timer_enable(False)
    pairs = float(n) * float(m)
    rec = {"value": pairs / (ms_per_step * 1e-3), "unit": UNIT, "ms_per_step": ms_per_step,
           "ms_per_step_blocks": [round(x, 5) for x in blocks], "timed_blocks": len(blocks),
           "precision": precision, "cuda_graph": graphed is not None,
           "hits": [int(x) for x in res["hits"].cpu().tolist()],
           "medr": float(res["medr"].cpu()[0]),
           "gpu_launches_per_step": cx.sum_over_ranks(float(launches)),
           "host_enqueue_ms_per_step": host_ms,
           "roofline": rank_roofline(tc_ms, tc_n, tc_steps, ms_per_step, q_local.shape[0], m, d,
                                     precision, timed_seconds=ms_per_step * 1e-3 * steps * len(blocks))}
    state = {"graphed": graphed, "rank0_local": res.get("rank0_local", res.get("rank0"))}
    if not want_state and graphed is not None:
        graphed.close()
        state["graphed"] = None
    return rec, state


def measure_e2e(cx: Ctx, T, V, qs, qe, gs, ge, n, m, d, precision, steps, graphed):
    """The same metric through the reference-facing call with pinned HOST tensors: H2D of the inputs
    and D2H of the hits inside the timed region, every step."""
    torch = cx.torch
    from vtc_b200.model.metric import RecallAtK
    from vtc_b200.parallel import sharded_rank_eval

    pairs = float(n) * float(m)
    n_e2e = max(3, min(steps, 10))
    if cx.world == 1:
        Tp, Vp = T.pin_memory(), V.pin_memory()
        metric = RecallAtK("videos", "titles", K_VALS, precision=precision)
        for _ in range(2):
            metric.compute(Vp, Tp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            rec = metric.compute(Vp, Tp)  # returns host floats: includes the D2H read
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        out = {"value": pairs / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(Tp.numel() * 4 + Vp.numel() * 4),
               "d2h_bytes_per_step": 8 * len(K_VALS), "ms_per_step": dt * 1e3,
               "api": "vtc_b200.model.metric.RecallAtK.compute(pinned fp32 host tensors)",
               "recall": [r for _, r in rec]}
        # the same as a STREAM of evaluations (parallel.PipelinedRankEval: two captured steps
        # alternate, the copies of evaluation k + 1 run while evaluation k is ranked); `e2e` itself
        # stays the one-evaluation-at-a-time drop-in call
        try:
            from vtc_b200.parallel import PipelinedRankEval
            pipe = PipelinedRankEval(T.to(cx.dev), V.to(cx.dev), n, m, K_VALS, "l2", precision)
            for _ in range(3):
                pipe.submit(Tp, Vp)
            pipe.flush()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                pipe.submit(Tp, Vp)
            last = pipe.flush()
            dtp = (time.perf_counter() - t0) / n_e2e
            pipe.close()
            out["pipelined"] = {"value": pairs / dtp, "unit": UNIT, "ms_per_step": dtp * 1e3,
                                "hits": [int(x) for x in last["hits"].tolist()],
                                "api": "vtc_b200.parallel.PipelinedRankEval.submit(pinned fp32 host "
                                       "tensors): copies of evaluation k+1 overlap the ranking of k"}
        except Exception as exc:  # noqa: BLE001  (an extra record: never fail the bench line on it)
            out["pipelined"] = {"error": repr(exc)[:200]}
        if precision == "bf16":
            # a caller that KEEPS its embeddings in bf16 on the host (the bf16 mode ranks the RN-even
            # bf16 roundings anyway: identical results): half the PCIe bytes, so the drop-in call is
            # bound by the ranking instead of the transfer.  An extra record; `e2e` stays fp32 input.
            try:
                Tp16, Vp16 = T.to(torch.bfloat16).pin_memory(), V.to(torch.bfloat16).pin_memory()
                for _ in range(2):
                    metric.compute(Vp16, Tp16)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(n_e2e):
                    rec16 = metric.compute(Vp16, Tp16)
                torch.cuda.synchronize()
                dt16 = (time.perf_counter() - t0) / n_e2e
                out["bf16_host"] = {"value": pairs / dt16, "unit": UNIT, "ms_per_step": dt16 * 1e3,
                                    "h2d_bytes_per_step": int(Tp16.numel() * 2 + Vp16.numel() * 2),
                                    "same_recall_as_fp32_input": [r for _, r in rec16] == out["recall"],
                                    "api": "RecallAtK.compute(pinned bf16 host tensors)"}
            except Exception as exc:  # noqa: BLE001
                out["bf16_host"] = {"error": repr(exc)[:200]}
        return out
    # N > 1: each rank stages ITS shards from pinned host memory, then the sharded eval.  With the
    # captured step the evaluations are pipelined (parallel.PipelinedRankEval): the copies of
    # evaluation k + 1 run while evaluation k is ranked; every evaluation's H2D and D2H are inside
    # the timed region.
    Tq, Vg = T[qs:qe].contiguous().pin_memory(), V[gs:ge].contiguous().pin_memory()
    n_e2e = max(n_e2e, 40)  # (a step is 1-5 ms here: time enough of them)
    pipe = None
    if graphed is not None:
        from vtc_b200.parallel import PipelinedRankEval
        pipe = PipelinedRankEval(graphed.q, graphed.g, n, m, K_VALS, "l2", precision)

    def e2e_step():
        if pipe is not None:
            return pipe.submit(Tq, Vg)
        ql = Tq.to(cx.dev, non_blocking=True)
        gl = Vg.to(cx.dev, non_blocking=True)
        return sharded_rank_eval(ql, gl, n, m, K_VALS, "l2", precision)["hits"].cpu()

    for _ in range(3):
        e2e_step()
    if pipe is not None:
        pipe.flush()
    cx.barrier()
    t0 = time.perf_counter()
    last = None
    for _ in range(n_e2e):
        last = e2e_step()
    if pipe is not None:
        last = pipe.flush()
    cx.barrier()
    dt = cx.max_over_ranks((time.perf_counter() - t0) / n_e2e)
    hits = None
    bf16_host = None
    if pipe is not None:
        hits = [int(x) for x in last["hits"].tolist()]
        pipe.close()
        if precision == "bf16":
            # the same stream from pinned bf16 host shards (a caller that keeps bf16 embeddings: the
            # mode ranks the bf16 roundings anyway): half the bytes through the host's memory.  An
            # extra record; `e2e` stays fp32 input.
            try:
                Tq16, Vg16 = Tq.to(torch.bfloat16).pin_memory(), Vg.to(torch.bfloat16).pin_memory()
                pipe16 = PipelinedRankEval(graphed.q.to(torch.bfloat16), graphed.g.to(torch.bfloat16),
                                           n, m, K_VALS, "l2", precision)
                for _ in range(3):
                    pipe16.submit(Tq16, Vg16)
                pipe16.flush()
                cx.barrier()
                t0 = time.perf_counter()
                for _ in range(n_e2e):
                    pipe16.submit(Tq16, Vg16)
                last16 = pipe16.flush()
                cx.barrier()
                dt16 = cx.max_over_ranks((time.perf_counter() - t0) / n_e2e)
                pipe16.close()
                bf16_host = {"value": pairs / dt16, "unit": UNIT, "ms_per_step": dt16 * 1e3,
                             "h2d_bytes_per_step": int(n * d * 2 + m * d * 2),
                             "same_hits_as_fp32_input": [int(x) for x in last16["hits"].tolist()] == hits,
                             "api": "PipelinedRankEval.submit(pinned bf16 host shards)"}
            except Exception as exc:  # noqa: BLE001  (an extra record: never fail the bench line on it)
                bf16_host = {"error": repr(exc)[:200]}
    return {"value": pairs / dt, "unit": UNIT, "h2d_bytes_per_step": int(n * d * 4 + m * d * 4),
            "d2h_bytes_per_step": 8 * len(K_VALS) * cx.world, "ms_per_step": dt * 1e3,
            "hits": hits,
            "api": ("vtc_b200.parallel.PipelinedRankEval.submit(pinned fp32 host shards): two captured "
                    "steps alternate, the copies of evaluation k+1 overlap the ranking of evaluation k"
                    if pipe is not None
                    else "vtc_b200.parallel.sharded_rank_eval(pinned fp32 host shards)"),
            **({"bf16_host": bf16_host} if bf16_host is not None else {})}


def parity_check(cx: Ctx, q_local, g_full, qs, ranks_by_mode, rows=256):
    """A slice of this rank's query rows re-ranked by the fp64 brute-force kernel (canonical
    arithmetic, csrc/exact.cu == oracle/vtc_oracle.c) against the FULL gallery, compared with the
    ranks the timed step produced in each mode.  bf16 mode ranks the bf16 roundings of the inputs."""
    torch = cx.torch
    from vtc_b200 import ops

    n_local = q_local.shape[0]
    ok = {}
    if n_local > 0:
        s = max(0, n_local // 2 - rows // 2)
        e = min(n_local, s + rows)
        for mode, rank0 in ranks_by_mode.items():
            if rank0 is None:
                continue
            q, g = q_local[s:e].contiguous(), g_full
            if mode == "bf16":
                q, g = q.bfloat16().float(), g.bfloat16().float()
            want, gts = ops.sim_rank(q, g, row_offset=qs + s, precision="brute")
            ops.rank_finalize(want, gts, g.shape[0], [1], want_medr=False)
            ok[mode] = bool(torch.equal(want, rank0[s:e].to(want.dtype)))
    out = {}
    for mode in ranks_by_mode:
        v = 1.0 if ok.get(mode, True) else 0.0
        allok = (-cx.max_over_ranks(-v)) > 0.5   # min over ranks
        out[mode] = "ok" if allok else "mismatch"
    out["rows_per_rank"] = rows
    out["how"] = "fp64 brute-force kernel on the device vs the timed path's rank0, same rows"
    return out


def measure_topk(cx: Ctx, steps, warmup, n=10_000, m=1_000_000, d=512, k=11, precision="bf16"):
    """BASELINE config 5: 10k queries x 1M gallery x 512, k = 11, gallery-sharded streaming top-k;
    the [N, k] candidates are all-gathered and merged.  Device-generated embeddings (1M x 512 on the
    host would take minutes)."""
    torch = cx.torch
    from vtc_b200.parallel import shard_bounds, sharded_topk

    gs, ge = shard_bounds(m, cx.world, cx.rank)
    gen = torch.Generator(device=cx.dev).manual_seed(1023 + cx.rank)
    G = torch.nn.functional.normalize(torch.randn(ge - gs, d, generator=gen, device=cx.dev), dim=1)
    base = G[:n].clone() if cx.rank == 0 else torch.empty(n, d, device=cx.dev)
    if cx.world > 1:
        cx.dist.broadcast(base, 0)  # queries = noisy copies of the first rows of shard 0
    gq = torch.Generator(device=cx.dev).manual_seed(7)
    Q = torch.nn.functional.normalize(
        base + 6.0 * torch.randn(n, d, generator=gq, device=cx.dev) / d ** 0.5, dim=1)

    def step():
        return sharded_topk(Q, G, m, k, precision=precision)

    ms, blocks, (vals, idx), _ = cx.time_blocks(step, steps, warmup)
    top1 = int((idx[:, 0] == torch.arange(n, device=cx.dev)).sum())
    peak_tf, peak_src, _, _ = peak_tflops(ms * 1e-3 * steps * len(blocks))
    tf = 2.0 * n * m * d / (ms * 1e-3) / 1e12
    return {"workload": f"streaming_topk_{n // 1000}kx{m // 1000000}M_{d}d_k{k}", "precision": precision,
            "ms_per_step": ms, "ms_per_step_blocks": [round(x, 5) for x in blocks],
            "value": float(n) * m / (ms * 1e-3), "unit": UNIT, "tflops_aggregate": tf,
            "frac_of_bf16_peak_per_gpu": tf / cx.world / peak_tf, "peak_source": peak_src,
            "r_at_1": top1 / n,
            "data": "synthetic, generated on the device (seed 1023 + rank)"}


def run_ours(a):
    cx = Ctx(a)
    torch = cx.torch
    from vtc_b200 import _ffi
    from vtc_b200.parallel import shard_bounds
    from vtc_b200.synthetic import make_retrieval_pair

    _ffi.load()
    world, rank, dev = cx.world, cx.rank, cx.dev
    warmup = max(a.warmup, 3)   # the timing rules ask for >= 3 warm-up steps
    T, V = make_retrieval_pair(a.n, a.m, a.d, seed=1023)
    qs, qe = shard_bounds(a.n, world, rank)
    gs, ge = shard_bounds(a.m, world, rank)
    q_local = T[qs:qe].contiguous().to(dev)
    g_local = V[gs:ge].contiguous().to(dev)
    if cx.sampler is not None:
        cx.sampler.start()  # before the warm-up: nvidia-smi takes ~0.1 s to deliver its first sample

    # ---- the headline record: device-resident, then end to end through the reference-facing call
    head, st = measure_rank(cx, q_local, g_local, a.n, a.m, a.d, a.precision, a.steps, warmup,
                            sample_clocks=True, use_graph=not a.no_graph, want_state=True)
    clocks = cx.sampler.stop() if cx.sampler is not None else None
    e2e = None
    if not a.no_e2e:
        e2e = measure_e2e(cx, T, V, qs, qe, gs, ge, a.n, a.m, a.d, a.precision, a.steps, st["graphed"])
    if st["graphed"] is not None:
        st["graphed"].close()  # NCCL will not tear a communicator down under a live captured graph
    ranks_by_mode = {a.precision: st["rank0_local"]}

    # ---- nested records
    other = None
    extra = {}
    if not a.no_extra and a.precision in ("bf16", "exact"):
        mode = "exact" if a.precision == "bf16" else "bf16"
        other, st2 = measure_rank(cx, q_local, g_local, a.n, a.m, a.d, mode, max(5, a.steps // 2),
                                  3, sample_clocks=False, use_graph=not a.no_graph, want_state=True)
        if not a.no_e2e:
            other["e2e"] = measure_e2e(cx, T, V, qs, qe, gs, ge, a.n, a.m, a.d, mode,
                                       max(3, a.steps // 2), st2["graphed"])
        if st2["graphed"] is not None:
            st2["graphed"].close()
        ranks_by_mode[mode] = st2["rank0_local"]
    if not a.no_extra:
        g_full = V.to(dev)
        extra["parity_check"] = parity_check(cx, q_local, g_full, qs, ranks_by_mode)
        del g_full
        # BASELINE config 4: 100k x 100k x 768 (ViT-L/14), same step, device-generated embeddings
        if (a.n, a.m, a.d) == (100_000, 100_000, 512):
            D4 = 768
            gen = torch.Generator(device=dev).manual_seed(1023)
            V4 = torch.nn.functional.normalize(torch.randn(a.m, D4, generator=gen, device=dev), dim=1)
            T4 = torch.nn.functional.normalize(
                V4[:a.n] + 7.0 * torch.randn(a.n, D4, generator=gen, device=dev) / D4 ** 0.5, dim=1)
            q4, g4 = T4[qs:qe].contiguous(), V4[gs:ge].contiguous()
            del T4, V4
            c4, _ = measure_rank(cx, q4, g4, a.n, a.m, D4, "bf16", max(5, a.steps // 2), 3,
                                 sample_clocks=False, use_graph=not a.no_graph)
            c4["workload"] = workload_name(a.n, a.m, D4)
            c4["data"] = "synthetic, generated on the device (seed 1023, sigma = 7)"
            extra["c4_d768"] = c4
            del q4, g4
            extra["c5_topk"] = measure_topk(cx, max(5, a.steps // 2), 3)

    if rank != 0:
        if world > 1:
            _teardown(cx.dist)
        return

    cpu = None
    if not a.no_cpu_baseline and world == 1:
        cpu = cpu_reference_pairs_per_s(T, V, a.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    dtype_of = {"bf16": "bf16", "exact": "f32 (bf16x3 tensor-core split + fp64 recheck)", "brute": "f64"}
    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": warmup, "warmup_requested": a.warmup, "ms_per_step": head["ms_per_step"],
        "ms_per_step_blocks": head["ms_per_step_blocks"], "timed_blocks": head["timed_blocks"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": dtype_of[a.precision], "data": DATA, "config": workload_config(a),
        "precision": a.precision,
        "impl_config": {"precision": a.precision,
                        "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU",
                        "cuda_graph": head["cuda_graph"],
                        "step": ("vtc_rank_eval: memset + row prologue + tcgen05 pass + cooperative "
                                 "epilogue" if world == 1 else
                                 "local chunk overlapped with the gallery all_gather, owners' norms "
                                 "gathered, remote ranges with cached per-row quantities, int32 rank "
                                 "all_gather + finalize"),
                        "drop_in_default": "RecallAtK(precision='exact') -- see the `exact` record"},
        "hits": head["hits"], "medr": head["medr"],
        "e2e": e2e, "gpu_launches": int(round(head["gpu_launches_per_step"] * a.steps)),
        "gpu_launches_per_step": head["gpu_launches_per_step"],
        "host_enqueue_ms_per_step": head["host_enqueue_ms_per_step"],
        "clocks": clocks, "roofline": head["roofline"], "cpu_baseline": cpu,
    }
    if other is not None:
        line[other["precision"]] = other
    line.update(extra)
    print(json.dumps(line), flush=True)
    if world > 1:
        _teardown(cx.dist)


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
