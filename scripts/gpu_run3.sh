#!/bin/bash
# Run 3: whole GPU suite per cluster size, bench (bf16 / exact / D sweep), ncu launch list + full.
set -u
mkdir -p gpurun_out
S=gpurun_out/summary3.txt
: > $S
run() { # name, cluster, -k expr
  VTC_CLUSTER=$2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=40 --tb=short -s -k "$3" \
      > gpurun_out/test3_$1.log 2>&1
  echo "$1 (cluster=$2) exit=$?" >> $S
  tail -n 3 gpurun_out/test3_$1.log >> $S
}
run all_c2 2 "not zzz"
run rank_c1 1 "rank or topk"
run rank_c4 4 "rank or topk"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1
echo "smoke exit=$?" >> $S
for c in 1 2 4; do
  VTC_CLUSTER=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/r3_bf16_c$c.json 2> gpurun_out/r3_bf16_c$c.err
done
VTC_CLUSTER=2 timeout 300 python bench.py --steps 5 --warmup 3 --precision exact --no-cpu-baseline --no-e2e \
    > gpurun_out/r3_exact_c2.json 2> gpurun_out/r3_exact_c2.err
for d in 128 256 768 1024; do
  VTC_CLUSTER=2 timeout 300 python bench.py --steps 10 --warmup 3 --d $d --no-cpu-baseline --no-e2e \
      > gpurun_out/r3_bf16_c2_d$d.json 2> gpurun_out/r3_bf16_c2_d$d.err
done
VTC_CLUSTER=2 timeout 300 python bench.py --steps 10 --warmup 3 --n 10000 --m 10000 --no-cpu-baseline --no-e2e \
    > gpurun_out/r3_bf16_c2_10k.json 2> gpurun_out/r3_bf16_c2_10k.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r3_bench_full.json 2> gpurun_out/r3_bench_full.err
echo "bench_full exit=$?" >> $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/launches_r01b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/ncu_launches3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sim_tc_kernel -s 3 -c 1 \
    -o gpurun_out/prof_rank_r01b -f python bench.py --steps 1 --warmup 3 --n 20000 --m 100000 \
    --no-cpu-baseline --no-e2e > gpurun_out/ncu_full3.log 2>&1
echo "ncu_full exit=$?" >> $S
cat $S
