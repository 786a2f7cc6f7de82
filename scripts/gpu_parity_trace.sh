#!/bin/bash
# Parity + launch traces + secondary benches with PDL on / off + one headline bench (1 GPU).  usage: gpu_parity_trace.sh <tag>
set -u
TAG=${1:-r02f}
mkdir -p gpurun_out
S=gpurun_out/summary_$TAG.txt
: > $S
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest -m gpu exit=$?" >> $S; tail -n 25 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-200 >> $S
for pdl in 1 0; do
VTC_PDL=$pdl timeout 600 python scripts/trace_once.py cam c3 rank > gpurun_out/${TAG}_trace_pdl$pdl.jsonl 2> gpurun_out/${TAG}_trace_pdl$pdl.err
echo "trace pdl=$pdl exit=$?" >> $S; tail -n 3 gpurun_out/${TAG}_trace_pdl$pdl.err >> $S
python - gpurun_out/${TAG}_trace_pdl$pdl.jsonl >> $S <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    d = json.loads(ln)
    print(d["what"], d.get("precision"), d.get("n", ""), "total_us", d["total_us"], "n", d["n_launches"])
    for w, us in d["launches"]:
        print("   %-28s %8.2f" % (w, us))
PY
VTC_PDL=$pdl timeout 600 python scripts/bench_extra.py c2 c35 > gpurun_out/${TAG}_bench_extra_pdl$pdl.jsonl 2> gpurun_out/${TAG}_bench_extra_pdl$pdl.err
echo "bench_extra pdl=$pdl exit=$?" >> $S; grep -v "torch\|train_step" gpurun_out/${TAG}_bench_extra_pdl$pdl.jsonl | cut -c1-220 >> $S
done
timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit=$?" >> $S
python scripts/show_bench.py gpurun_out/${TAG}_bench.json >> $S
cat $S
