#!/bin/bash
# CTA-pair (cta_group::2) bring-up: parity subset + headline bench with VTC_PAIR=1 vs 0.
set -u
mkdir -p gpurun_out
S=gpurun_out/summary_pair.txt
: > $S
VTC_PAIR=1 timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 --tb=short \
    -k "rank or topk or recall" > gpurun_out/test_pair.log 2>&1
echo "pair tests exit=$?" >> $S; tail -n 12 gpurun_out/test_pair.log >> $S
for pm in 1 0; do
  VTC_PAIR=$pm timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/pair${pm}_bf16.json 2> gpurun_out/pair${pm}_bf16.err
  echo "bench bf16 pair=$pm exit=$?" >> $S
  VTC_PAIR=$pm timeout 120 python bench.py --steps 5 --warmup 3 --precision exact --no-cpu-baseline --no-e2e \
      > gpurun_out/pair${pm}_exact.json 2> gpurun_out/pair${pm}_exact.err
  echo "bench exact pair=$pm exit=$?" >> $S
done
cat $S
python scripts/show_bench.py gpurun_out/pair*_*.json 2>&1 | cut -c1-260
