#!/bin/bash
# Quick perf iteration: rank/top-k tests, headline bench per cluster size, one ncu capture.
# usage: bash scripts/gpu_quick.sh <tag>
set -u
TAG=${1:-q}
mkdir -p gpurun_out
S=gpurun_out/summary_$TAG.txt
: > $S
VTC_CLUSTER=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=20 --tb=short -k "rank or topk or recall or clip_loss" \
    > gpurun_out/test_$TAG.log 2>&1
echo "tests exit=$?" >> $S; tail -n 3 gpurun_out/test_$TAG.log >> $S
for c in 1 2; do
  VTC_CLUSTER=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/${TAG}_bf16_c$c.json 2> gpurun_out/${TAG}_bf16_c$c.err
done
VTC_CLUSTER=2 timeout 300 python bench.py --steps 5 --warmup 3 --precision exact --no-cpu-baseline --no-e2e \
    > gpurun_out/${TAG}_exact_c2.json 2> gpurun_out/${TAG}_exact_c2.err
for d in 256 768; do
  VTC_CLUSTER=2 timeout 300 python bench.py --steps 10 --warmup 3 --d $d --no-cpu-baseline --no-e2e \
      > gpurun_out/${TAG}_bf16_c2_d$d.json 2> gpurun_out/${TAG}_bf16_c2_d$d.err
done
VTC_CLUSTER=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:sim_tc_kernel -s 3 -c 1 \
    -o gpurun_out/prof_rank_$TAG -f python bench.py --steps 1 --warmup 3 --n 20000 --m 100000 \
    --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu exit=$?" >> $S
cat $S
