#!/bin/bash
# First GPU call of a round that inherits opt-in kernel variants (DESIGN.md §8 "pending
# validation"): parity of every variant against the oracle, then the headline bench with and
# without it on the SAME box (boxes differ in how hard sw_power_cap bites), plus the shapes where the
# variant should matter most.  usage: gpurun --timeout 1800 -- 'bash scripts/gpu_experimental.sh [tag]'; ~15 GPU-minutes.
set -u
TAG=${1:-exp}
mkdir -p gpurun_out
S=gpurun_out/summary_$TAG.txt
: > $S
# one pytest process per variant: a trap in one kernel poisons its CUDA context, not the other groups
for grp in "fold and cols64" "fold and cols16" "fold and not cols64 and not cols16" \
           "fast_thresholds" "prepared" "more_than_eight" "infinite_rows"; do
  name=$(echo "$grp" | tr -d ' ' | cut -c1-24)
  VTC_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q \
      --tb=short --maxfail=6 -k "$grp" > gpurun_out/${TAG}_pytest_$name.log 2>&1
  echo "parity [$grp] exit=$?" >> $S; tail -n 4 gpurun_out/${TAG}_pytest_$name.log >> $S
done
# fold = 0: default epilogue; 1: fold operands 64 columns wide; 2: 16 columns wide (VTC_FOLD_COLS=16)
for fold in 0 1 2; do
  for args in "" "--d 256 --no-e2e" "--d 768 --no-e2e" "--precision exact --no-e2e"; do
    name=$(echo "fold${fold}${args}" | sed 's/--no-e2e//' | tr -d ' -')
    cols=64; [ $fold -eq 2 ] && cols=16
    VTC_RANK_FOLD=$(( fold > 0 )) VTC_FOLD_COLS=$cols timeout 200 python bench.py --steps 10 \
        --no-cpu-baseline $args > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
    echo "bench fold=$fold $args exit=$?" >> $S
  done
done
# guard-band thresholds from a coalesced fp32 norm (fewer fp64 row walks per library call): the e2e
# number has 11 calls per evaluation, the device-resident step one
for ft in 0 1; do
  VTC_FAST_THR=$ft timeout 200 python bench.py --steps 10 --no-cpu-baseline \
      > gpurun_out/${TAG}_fold0_fastthr$ft.json 2> gpurun_out/${TAG}_fastthr$ft.err
  echo "bench fast_thr=$ft exit=$?" >> $S
done
# host staging of RecallAtK.compute (the e2e number): equal chunks vs the balanced schedule, per-call
# row walks vs per-chunk prepared quantities (vtc_sim_rank_prepared)
for prep in 0 1; do
  for sch in equal balanced; do
    for c in 6 8; do
      VTC_RANK_PREPARED=$prep VTC_PIPELINE_SCHEDULE=$sch VTC_PIPELINE_CHUNKS_2D=$c timeout 200 \
          python bench.py --steps 10 --no-cpu-baseline \
          > gpurun_out/${TAG}_fold0_e2e_prep${prep}_${sch}_c$c.json 2> gpurun_out/${TAG}_e2e_prep${prep}_${sch}_c$c.err
      echo "bench e2e prepared=$prep schedule=$sch c=$c exit=$?" >> $S
    done
  done
done
# top-k sample pass size (default: 1/32 of the gallery, 8..32 tiles): config-5 share of one GPU
for st in 6 8 12 24; do
  VTC_TOPK_SAMPLE_TILES=$st timeout 200 python scripts/bench_extra.py c35 2>/dev/null | grep c5_topk \
      | sed "s/^/sample_tiles=$st /" >> gpurun_out/${TAG}_topk_sample_tiles.txt
done
timeout 200 python scripts/bench_extra.py c35 2>/dev/null | grep c5_topk | sed "s/^/sample_tiles=default /" \
    >> gpurun_out/${TAG}_topk_sample_tiles.txt
cat gpurun_out/${TAG}_topk_sample_tiles.txt | cut -c1-200 >> $S
python scripts/show_bench.py gpurun_out/${TAG}_fold*.json 2>&1 | cut -c1-200 >> $S
# SASS-level proof of what the fold epilogue issues per logit goes with the ncu capture of the round
VTC_RANK_FOLD=1 timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:sim_tc_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_rank_fold -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_fold.log 2>&1
echo "ncu fold exit=$?" >> $S
cat $S
