#!/usr/bin/env python
"""One line per bench JSON file: python scripts/show_bench.py gpurun_out/r3_*.json"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        e = d.get("e2e") or {}
        c = d.get("clocks") or {}
        cb = d.get("cpu_baseline") or {}
        print(f"{f.split('/')[-1]:28s} ms/step {d['ms_per_step']:8.3f} value {d['value']:.3e} "
              f"tc_ms {r.get('ms_per_launch', 0):7.3f} ach {r.get('achieved', 0):7.1f} "
              f"iss {r.get('issued_tflops', 0):7.1f} TF frac {r.get('frac', 0):.3f} "
              f"share {r.get('kernel_share_of_step', 0):.2f} e2e {e.get('value', 0):.3e} "
              f"cpu {cb.get('value', 0):.2e} clk {c.get('sm_mhz')} {c.get('reasons')} L {d.get('gpu_launches')}")
    except Exception as ex:  # noqa: BLE001
        print(f, "ERR", ex)
