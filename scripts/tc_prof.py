#!/usr/bin/env python
"""Where the tcgen05 similarity + rank kernel spends its time (profiling aid, not a bench value).

Runs the headline step with VTC_DBG_PROF=1: the kernel's MMA issuer (one thread per CTA) and first
epilogue warp record clock64 / globaltimer and the ticks they spend waiting on each mbarrier.  Prints
one JSON line per configuration:
  sm_mhz            clock64 ticks / globaltimer ns of the kernel itself (the real SM clock under the
                    power cap -- nvidia-smi's 25 ms samples cannot resolve a 7 ms kernel)
  clk_per_tile      issuer ticks per 128 x 256 tile (floor: K'/16 * 128)
  wait_acc_frac     share of the issuer's time waiting for a FREE accumulator  -> epilogue-bound
  wait_ld_frac      share waiting for operand stages                           -> load-bound
  epi_wait_frac     share of the epilogue's time waiting for a FULL accumulator (its slack)

    python scripts/tc_prof.py [--d 512] [--precision bf16] [--n 100000]
Set VTC_DBG_SKIP_EPILOGUE=1 / VTC_PAIR=0|1 / VTC_CLUSTER=1|2|4 in the environment to compare variants.
"""
import argparse
import json
import os
import statistics
import sys

os.environ["VTC_DBG_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from vtc_b200 import _ffi, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--m", type=int, default=0)
    ap.add_argument("--d", type=int, default=512)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    m = a.m or a.n
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(1023)
    V = torch.nn.functional.normalize(torch.randn(m, a.d, generator=gen, device=dev), dim=1)
    T = torch.nn.functional.normalize(
        V[:a.n] + 6.0 * torch.randn(a.n, a.d, generator=gen, device=dev) / a.d ** 0.5, dim=1)
    _ffi.load()
    for _ in range(3):
        ops.rank_eval(T, V, [1, 5, 10], precision=a.precision)
    torch.cuda.synchronize()
    _ffi.debug_prof_read()
    rows = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(a.reps):
        ev0.record()
        ops.rank_eval(T, V, [1, 5, 10], precision=a.precision)
        ev1.record()
        torch.cuda.synchronize()
        ms.append(ev0.elapsed_time(ev1))
        prof = [p for p in _ffi.debug_prof_read(256) if p[4] > 0]
        if not prof:
            continue
        ticks = statistics.median(p[0] for p in prof)
        ns = statistics.median(p[1] for p in prof)
        tiles = sum(p[4] for p in prof)
        rows.append({
            "sm_mhz": 1e3 * ticks / ns, "kernel_ms": ns * 1e-6,
            "clk_per_tile": statistics.median(p[0] / p[4] for p in prof),
            "wait_acc_frac": statistics.median(p[2] / p[0] for p in prof),
            "wait_ld_frac": statistics.median(p[3] / p[0] for p in prof),
            "epi_wait_frac": statistics.median((p[5] / p[6]) if p[6] else 0.0 for p in prof),
            "issuing_ctas": len(prof), "tiles": tiles,
        })
    kp = a.d if a.precision == "bf16" else 3 * a.d
    kp = -(-kp // 64) * 64
    med = {k: statistics.median(r[k] for r in rows) for k in rows[0]} if rows else {}
    out = {"cfg": {"n": a.n, "m": m, "d": a.d, "precision": a.precision,
                   "skip_epilogue": os.environ.get("VTC_DBG_SKIP_EPILOGUE", "0"),
                   "pair": os.environ.get("VTC_PAIR", "auto"),
                   "cluster": os.environ.get("VTC_CLUSTER", "2")},
           "step_ms": statistics.median(ms), "floor_clk_per_tile": kp // 16 * 128, **med}
    if med:
        out["tflops_alg"] = 2.0 * a.n * m * a.d / (med["kernel_ms"] * 1e-3) / 1e12
        out["tensor_active_est"] = out["floor_clk_per_tile"] / med["clk_per_tile"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
