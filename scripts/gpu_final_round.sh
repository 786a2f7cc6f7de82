#!/bin/bash
# The trimmed end-of-round evidence run (1 GPU): full -m gpu suite, smoke, both bench arms, secondary
# benches, launch traces, ncu launch list + one full capture of the rank kernel.  (The shape / epilogue-off
# variants of gpu_profile_round.sh are not repeated: the tensor-core kernel did not change.)
# usage: gpurun -- bash scripts/gpu_final_round.sh <tag>
set -u
TAG=${1:-r02f}
mkdir -p gpurun_out
S=gpurun_out/summary_$TAG.txt
: > $S
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest -m gpu exit=$?" >> $S; tail -n 3 gpurun_out/${TAG}_pytest_gpu.log >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit=$?" >> $S
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
echo "bench reference exit=$?" >> $S
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit=$?" >> $S
timeout 600 python scripts/bench_extra.py c2 c35 > gpurun_out/${TAG}_bench_extra.jsonl 2> gpurun_out/${TAG}_bench_extra.err
echo "bench_extra exit=$?" >> $S
timeout 300 python scripts/trace_once.py cam c3 rank > gpurun_out/${TAG}_trace_1gpu.jsonl 2> gpurun_out/${TAG}_trace_1gpu.err
echo "trace exit=$?" >> $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches exit=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sim_tc_kernel -s 3 -c 1 \
    -o gpurun_out/${TAG}_prof_rank_full -f python bench.py --steps 1 --warmup 3 \
    --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit=$?" >> $S
python scripts/show_bench.py gpurun_out/${TAG}_bench.json gpurun_out/${TAG}_bench_reference.json >> $S 2>&1
cat $S
