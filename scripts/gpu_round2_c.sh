#!/bin/bash
# Round 2, third run (1 GPU): parity after the balanced last round, the one-pass InfoNCE column sums
# and the transposing StoreEpi; cluster 1 vs 2 under the sustained bench loop; secondary configs.
set -u
TAG=${1:-r02c}
mkdir -p gpurun_out
S=gpurun_out/summary_$TAG.txt
: > $S
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest -m gpu exit=$?" >> $S; tail -n 25 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-200 >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit=$?" >> $S; tail -n 2 gpurun_out/${TAG}_smoke.log >> $S
for cl in 2 1 2 1; do
  VTC_CLUSTER=$cl timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e --no-extra \
      > gpurun_out/${TAG}_bench_cl$cl.json 2>> gpurun_out/${TAG}_bench_cl.err
  python - "$cl" "gpurun_out/${TAG}_bench_cl$cl.json" >> $S <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
r = d["roofline"]
print("cluster", sys.argv[1], "ms/step %.3f" % d["ms_per_step"], "blocks", d["ms_per_step_blocks"][:3], "..", "kernel ms %.3f" % r["ms_per_launch"],
      "TF %.0f" % r["achieved"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
timeout 600 python scripts/bench_extra.py c2 c35 > gpurun_out/${TAG}_bench_extra.jsonl 2> gpurun_out/${TAG}_bench_extra.err
echo "bench_extra exit=$?" >> $S; grep -v "torch\|train_step" gpurun_out/${TAG}_bench_extra.jsonl | cut -c1-220 >> $S
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches_cam.csv python scripts/cam_once.py \
    > gpurun_out/${TAG}_ncu_launches_cam.log 2>&1
echo "ncu cam launches exit=$?" >> $S
cat $S
