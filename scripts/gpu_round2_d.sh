#!/bin/bash
# Round 2 (1 GPU): the two bench arms exactly as the driver runs them (default flags), timed by wall clock.
set -u
TAG=${1:-r02d}
mkdir -p gpurun_out
S=gpurun_out/summary_$TAG.txt
: > $S
t0=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
echo "reference exit=$? wall=$(( $(date +%s) - t0 ))s" >> $S
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit=$? wall=$(( $(date +%s) - t0 ))s" >> $S
tail -n 5 gpurun_out/${TAG}_bench.err >> $S
cut -c1-1500 gpurun_out/${TAG}_bench_reference.json >> $S
cat $S
