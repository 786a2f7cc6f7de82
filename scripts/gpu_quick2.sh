#!/bin/bash
# usage: gpu_quick2.sh <tag>: whole -m gpu suite + secondary benches + headline bench
set -u
TAG=${1:-q}
mkdir -p gpurun_out
S=gpurun_out/summary_$TAG.txt
: > $S
timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=30 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit=$?" >> $S; tail -n 12 gpurun_out/${TAG}_pytest.log >> $S
timeout 900 python scripts/bench_extra.py c2 c35 > gpurun_out/${TAG}_extra.jsonl 2> gpurun_out/${TAG}_extra.err
echo "extra exit=$?" >> $S
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit=$?" >> $S
cat $S
