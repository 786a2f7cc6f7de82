#!/bin/bash
# NCCL parity and the bench line on all GPUs of the box, optionally (TRACE=1) the launch trace of the
# first and of a middle rank (usage: gpurun --gpus N -- bash scripts/gpu_multi_quick.sh [tag])
set -u
TAG=${1:-r02q}
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
S=gpurun_out/summary_${TAG}_n$NG.txt
echo "gpus=$NG" > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 500 python -m pytest tests/test_multigpu.py -m gpu -q --tb=short ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/${TAG}_test_multi_n$NG.log 2>&1
echo "test_multigpu exit=$?" >> $S; grep -E "rank .*(OK|MISMATCH)|passed|failed|Error" gpurun_out/${TAG}_test_multi_n$NG.log | tail -n 30 >> $S
for C in 0; do
  timeout 300 $TR --master-port 2953$C bench.py --gpus $NG --steps 20 --warmup 3 \
      > gpurun_out/${TAG}_scale_n${NG}.json 2> gpurun_out/${TAG}_scale_n${NG}.err
  echo "bench n=$NG exit=$?" >> $S
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/${TAG}_scale_n${NG}.err | tail -n 5 >> $S
  python - gpurun_out/${TAG}_scale_n${NG}.json >> $S <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if not ln.startswith("{"): continue
    d = json.loads(ln)
    print("N", d["n_gpus"], "ms/step", d["ms_per_step"], d["ms_per_step_blocks"][:4], "value %.3e" % d["value"], "e2e", d["e2e"] and "%.3e" % d["e2e"]["value"],
          "graph", d["impl_config"]["cuda_graph"], "hits", d["hits"], "parity", d.get("parity_check"))
    r = d["roofline"]; print("  roofline", r["achieved"], r["frac"], r["ms_per_launch"], r["launches_per_step"], r["kernel_share_of_step"])
    for k in ("exact", "c4_d768", "c5_topk"):
        if k in d: print(" ", k, d[k]["ms_per_step"], "%.3e" % d[k]["value"])
PY
done
if [ -n "${TRACE:-}" ]; then
  timeout 200 $TR --master-port 29539 scripts/dist_trace.py > gpurun_out/${TAG}_trace_n$NG.jsonl 2> gpurun_out/${TAG}_trace_n$NG.err
  echo "trace exit=$?" >> $S
  python - gpurun_out/${TAG}_trace_n$NG.jsonl >> $S <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if not ln.startswith("{"): continue
    d = json.loads(ln)
    print("rank", d.get("rank"), "step_us", d["step_us"], "traced_us", d["traced_us"])
    for w, us in d["launches"]:
        print("   %-28s %8.2f" % (w, us))
PY
fi
cat $S
