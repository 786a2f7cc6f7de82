#!/bin/bash
# Run 2: full GPU test suite (per group), cluster-multicast variants, feed-vs-epilogue diagnostics,
# and one full ncu capture of the rank kernel.
set -u
mkdir -p gpurun_out
: > gpurun_out/summary2.txt
run() { # name, cluster, -k expr
  VTC_CLUSTER=$2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=40 --tb=short -s -k "$3" \
      > gpurun_out/test2_$1.log 2>&1
  echo "$1 (cluster=$2) exit=$?" >> gpurun_out/summary2.txt
  tail -n 3 gpurun_out/test2_$1.log >> gpurun_out/summary2.txt
}
run rank_c1 1 "rank and not brute"
run rank_c2 2 "rank and not brute"
run rank_c4 4 "rank and not brute"
run topk_c2 2 "topk"
run topk_c4 4 "topk"
run recall 2 "recall"
run rest 2 "not rank and not topk and not recall"
VTC_CLUSTER=2 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1
echo "smoke exit=$?" >> gpurun_out/summary2.txt
for c in 1 2 4; do
  VTC_CLUSTER=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/diag_bf16_c$c.json 2> gpurun_out/diag_bf16_c$c.err
  VTC_CLUSTER=$c timeout 300 python bench.py --steps 5 --warmup 3 --precision exact --no-cpu-baseline --no-e2e \
      > gpurun_out/diag_exact_c$c.json 2> gpurun_out/diag_exact_c$c.err
done
for d in 128 256 1024; do
  VTC_CLUSTER=1 timeout 300 python bench.py --steps 10 --warmup 3 --d $d --no-cpu-baseline --no-e2e \
      > gpurun_out/diag_bf16_c1_d$d.json 2> gpurun_out/diag_bf16_c1_d$d.err
  VTC_CLUSTER=2 timeout 300 python bench.py --steps 10 --warmup 3 --d $d --no-cpu-baseline --no-e2e \
      > gpurun_out/diag_bf16_c2_d$d.json 2> gpurun_out/diag_bf16_c2_d$d.err
done
VTC_CLUSTER=2 timeout 300 python bench.py --steps 10 --warmup 3 --n 10000 --m 10000 --no-cpu-baseline --no-e2e \
    > gpurun_out/diag_bf16_c2_10k.json 2> gpurun_out/diag_bf16_c2_10k.err
for c in 1 2; do
VTC_CLUSTER=$c timeout 900 ncu --set full --clock-control none --import-source on -k regex:sim_tc_kernel -s 3 -c 1 \
    -o gpurun_out/prof_rank_c${c}_r01 -f python bench.py --steps 1 --warmup 3 --n 20000 --m 100000 \
    --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c$c.log 2>&1
echo "ncu_full c=$c exit=$?" >> gpurun_out/summary2.txt
done
cat gpurun_out/summary2.txt
