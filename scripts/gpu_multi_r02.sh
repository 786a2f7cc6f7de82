#!/bin/bash
# Round 2, all GPUs of the box: NCCL parity, the driver's bench line, and the per-launch trace of the
# sharded step (usage: gpurun --gpus N -- bash scripts/gpu_multi_r02.sh [tag])
set -u
TAG=${1:-r02m}
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
S=gpurun_out/summary_${TAG}_n$NG.txt
echo "gpus=$NG" > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q --tb=short > gpurun_out/${TAG}_test_multi_n$NG.log 2>&1
echo "test_multigpu exit=$?" >> $S; tail -n 6 gpurun_out/${TAG}_test_multi_n$NG.log >> $S
timeout 400 $TR --master-port 29531 bench.py --gpus $NG --steps 20 --warmup 3 \
    > gpurun_out/${TAG}_scale_n$NG.json 2> gpurun_out/${TAG}_scale_n$NG.err
echo "bench n=$NG exit=$?" >> $S
tail -n 5 gpurun_out/${TAG}_scale_n$NG.err >> $S
timeout 200 $TR --master-port 29532 scripts/dist_trace.py > gpurun_out/${TAG}_trace_n$NG.jsonl 2> gpurun_out/${TAG}_trace_n$NG.err
echo "trace exit=$?" >> $S
python - gpurun_out/${TAG}_trace_n$NG.jsonl gpurun_out/${TAG}_scale_n$NG.json >> $S <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if not ln.startswith("{"): continue
    d = json.loads(ln)
    print("rank", d.get("rank"), "step_us", d["step_us"], "traced_us", d["traced_us"])
    for w, us in d["launches"]:
        print("   %-28s %8.2f" % (w, us))
for ln in open(sys.argv[2]):
    if not ln.startswith("{"): continue
    d = json.loads(ln)
    print("N", d["n_gpus"], "ms/step", d["ms_per_step"], d["ms_per_step_blocks"][:4], "value %.3e" % d["value"], "e2e", d["e2e"] and "%.3e" % d["e2e"]["value"],
          "graph", d["impl_config"]["cuda_graph"], "hits", d["hits"], "parity", d.get("parity_check"), "clocks", d["clocks"])
    r = d["roofline"]; print("  roofline", r["achieved"], r["frac"], r["ms_per_launch"], r["launches_per_step"], r["kernel_share_of_step"])
    for k in ("exact", "c4_d768", "c5_topk"):
        if k in d: print(" ", k, d[k]["ms_per_step"], "%.3e" % d[k]["value"])
PY
cat $S
