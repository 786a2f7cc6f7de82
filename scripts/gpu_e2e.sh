#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 --tb=short -k "recall or eval" > gpurun_out/test_e2e.log 2>&1
echo "tests exit=$?"; tail -n 5 gpurun_out/test_e2e.log
for c in 4 6 8; do
  VTC_PIPELINE_CHUNKS_2D=$c timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline \
      > gpurun_out/e2e_c$c.json 2> gpurun_out/e2e_c$c.err || echo "c=$c failed"
done
python scripts/show_bench.py gpurun_out/e2e_*.json 2>&1 | cut -c1-220
