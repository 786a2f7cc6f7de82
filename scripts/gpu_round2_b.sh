#!/bin/bash
# Round 2, second run: the rank kernel with the query tile in tensor memory (VTC_TS) against the
# shared-memory operand kernel on the same box; CAM row kernels after the rewrite.
set -u
TAG=${1:-r02b}
mkdir -p gpurun_out
S=gpurun_out/summary_$TAG.txt
: > $S
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "rank or cached or infinite or recall or eval_one" > gpurun_out/${TAG}_pytest_rank_ts.log 2>&1
echo "pytest rank (TS on) exit=$?" >> $S; tail -n 12 gpurun_out/${TAG}_pytest_rank_ts.log >> $S
VTC_TS=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -x -k "rank or cached or infinite or recall or eval_one" > gpurun_out/${TAG}_pytest_rank_ss.log 2>&1
echo "pytest rank (TS off) exit=$?" >> $S; tail -n 4 gpurun_out/${TAG}_pytest_rank_ss.log >> $S
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "cam or readout or attn or averaging or fusion or transformer or training" > gpurun_out/${TAG}_pytest_cam.log 2>&1
echo "pytest cam exit=$?" >> $S; tail -n 12 gpurun_out/${TAG}_pytest_cam.log >> $S
P=gpurun_out/${TAG}_tc_prof.jsonl
: > $P
for rep in 1 2; do
  VTC_TS=1 timeout 120 python scripts/tc_prof.py >> $P 2>> gpurun_out/${TAG}_tc_prof.err
  VTC_TS=0 timeout 120 python scripts/tc_prof.py >> $P 2>> gpurun_out/${TAG}_tc_prof.err
  VTC_TS=0 VTC_CLUSTER=1 timeout 120 python scripts/tc_prof.py >> $P 2>> gpurun_out/${TAG}_tc_prof.err
done
VTC_TS=1 VTC_DBG_SKIP_EPILOGUE=1 timeout 120 python scripts/tc_prof.py >> $P 2>> gpurun_out/${TAG}_tc_prof.err
VTC_TS=0 VTC_CLUSTER=1 VTC_DBG_SKIP_EPILOGUE=1 timeout 120 python scripts/tc_prof.py >> $P 2>> gpurun_out/${TAG}_tc_prof.err
VTC_TS=1 timeout 120 python scripts/tc_prof.py --d 256 >> $P 2>> gpurun_out/${TAG}_tc_prof.err
VTC_TS=0 timeout 120 python scripts/tc_prof.py --d 256 >> $P 2>> gpurun_out/${TAG}_tc_prof.err
VTC_TS=1 timeout 120 python scripts/tc_prof.py --n 10000 >> $P 2>> gpurun_out/${TAG}_tc_prof.err
python - <<'PY' >> $S
import json
for ln in open(__import__("glob").glob("gpurun_out/*r02b*_tc_prof.jsonl")[0] if False else "gpurun_out/r02b_tc_prof.jsonl"):
    d = json.loads(ln)
    c = d["cfg"]
    print("n=%d d=%d ts=%s cl=%s skip=%s | step %.3f ms kernel %.3f ms %.0f MHz clk/tile %.0f (floor %d) acc %.3f ld %.3f epiwait %.3f TF %.0f" % (
        c["n"], c["d"], c.get("ts"), c["cluster"], c["skip_epilogue"], d["step_ms"], d.get("kernel_ms", 0),
        d.get("sm_mhz", 0), d.get("clk_per_tile", 0), d["floor_clk_per_tile"], d.get("wait_acc_frac", 0),
        d.get("wait_ld_frac", 0), d.get("epi_wait_frac", 0), d.get("tflops_alg", 0)))
PY
timeout 600 python scripts/bench_extra.py c2 > gpurun_out/${TAG}_bench_extra.jsonl 2> gpurun_out/${TAG}_bench_extra.err
echo "bench_extra exit=$?" >> $S; grep "cam_adapt\|train_step_cam" gpurun_out/${TAG}_bench_extra.jsonl | cut -c1-200 >> $S
cat $S
