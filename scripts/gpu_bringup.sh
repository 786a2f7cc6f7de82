#!/bin/bash
# First-contact run on the B200 box: every test group in its own process (a trapped kernel kills
# only that group's CUDA context), then a short bench and an ncu launch list.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
run() { # name, -k expression
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=40 --tb=short -s -k "$2" \
      > gpurun_out/test_$1.log 2>&1
  echo "$1 exit=$?" >> gpurun_out/summary.txt
  tail -n 3 gpurun_out/test_$1.log >> gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run h1 "normalize"
run brute "brute and not topk"
run gemm "sim_matrix"
run rank_tc "rank and not brute"
run recall "recall"
run topk "topk"
run loss "clip_loss"
run cam "cam or averaging"
run misc "sharded or full_size"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
echo "bench exit=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --precision exact --no-cpu-baseline > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err
echo "bench_exact exit=$?" >> gpurun_out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/ncu_launches.log 2>&1
echo "ncu_launches exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
