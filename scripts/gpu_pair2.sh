#!/bin/bash
set -u
mkdir -p gpurun_out
run() { # name, env..., args
  local name=$1; shift
  env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $ARGS \
      > gpurun_out/$name.json 2> gpurun_out/$name.err || echo "$name failed"
}
ARGS=""
run px_skip_p0 VTC_PAIR=0 VTC_DBG_SKIP_EPILOGUE=1
run px_skip_p1 VTC_PAIR=1 VTC_DBG_SKIP_EPILOGUE=1
ARGS="--d 768"
run px_d768_p0 VTC_PAIR=0
run px_d768_p1 VTC_PAIR=1
ARGS="--d 256"
run px_d256_p0 VTC_PAIR=0
run px_d256_p1 VTC_PAIR=1
ARGS="--precision exact"
run px_exact_skip_p1 VTC_PAIR=1 VTC_DBG_SKIP_EPILOGUE=1
python scripts/show_bench.py gpurun_out/px_*.json 2>&1 | cut -c1-200
