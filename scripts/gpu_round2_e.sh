#!/bin/bash
# Round 2 (1 GPU): per-launch warm times of the secondary paths (launch trace), and the headline
# kernel's cluster flavours (multicast / CTA pair / single) under the sustained bench loop.
set -u
TAG=${1:-r02e}
mkdir -p gpurun_out
S=gpurun_out/summary_$TAG.txt
: > $S
nvidia-smi --query-gpu=power.limit,enforced.power.limit,power.max_limit,clocks.max.sm --format=csv >> $S
timeout 600 python scripts/trace_once.py cam c3 topk nce > gpurun_out/${TAG}_trace.jsonl 2> gpurun_out/${TAG}_trace.err
echo "trace exit=$?" >> $S; tail -n 3 gpurun_out/${TAG}_trace.err >> $S
python - gpurun_out/${TAG}_trace.jsonl >> $S <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    d = json.loads(ln)
    print(d["what"], d.get("precision"), d.get("n", ""), "total_us", d["total_us"], "n", d["n_launches"])
    for w, us in d["launches"]:
        print("   %-28s %8.2f" % (w, us))
PY
for v in "VTC_CLUSTER=2 VTC_PAIR=0" "VTC_CLUSTER=2 VTC_PAIR=1" "VTC_CLUSTER=1" "VTC_CLUSTER=2 VTC_PAIR=0" "VTC_CLUSTER=2 VTC_PAIR=1" "VTC_CLUSTER=1"; do
  env $v timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e --no-extra \
      > gpurun_out/${TAG}_bench_v.json 2>> gpurun_out/${TAG}_bench_v.err
  python - "$v" gpurun_out/${TAG}_bench_v.json >> $S <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
r = d["roofline"]
print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "blocks", d["ms_per_step_blocks"], "kernel ms %.3f" % r["ms_per_launch"],
      "TF %.0f" % r["achieved"], "clk", d["clocks"]["sm_mhz"], "W", d["clocks"].get("power_w_median"), d["clocks"].get("power_limit_w"), d["clocks"]["reasons"])
PY
done
cat $S
