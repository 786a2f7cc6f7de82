#!/bin/bash
# Round 2, first evidence run after the rank path was rebuilt around rank_stage.cu (1 GPU):
# full parity suite, smoke, where the tensor-core kernel waits (tc_prof.py), a short bench, the
# secondary configs and the CAM launch list.
set -u
TAG=${1:-r02a}
mkdir -p gpurun_out
S=gpurun_out/summary_$TAG.txt
: > $S
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest -m gpu exit=$?" >> $S; tail -n 15 gpurun_out/${TAG}_pytest_gpu.log >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke exit=$?" >> $S; tail -n 2 gpurun_out/${TAG}_smoke.log >> $S
P=gpurun_out/${TAG}_tc_prof.jsonl
: > $P
timeout 120 python scripts/tc_prof.py >> $P 2>> gpurun_out/${TAG}_tc_prof.err
VTC_DBG_SKIP_EPILOGUE=1 timeout 120 python scripts/tc_prof.py >> $P 2>> gpurun_out/${TAG}_tc_prof.err
VTC_PAIR=1 timeout 120 python scripts/tc_prof.py >> $P 2>> gpurun_out/${TAG}_tc_prof.err
VTC_PAIR=1 VTC_DBG_SKIP_EPILOGUE=1 timeout 120 python scripts/tc_prof.py >> $P 2>> gpurun_out/${TAG}_tc_prof.err
VTC_CLUSTER=1 timeout 120 python scripts/tc_prof.py >> $P 2>> gpurun_out/${TAG}_tc_prof.err
timeout 120 python scripts/tc_prof.py --d 256 >> $P 2>> gpurun_out/${TAG}_tc_prof.err
VTC_DBG_SKIP_EPILOGUE=1 timeout 120 python scripts/tc_prof.py --d 256 >> $P 2>> gpurun_out/${TAG}_tc_prof.err
timeout 120 python scripts/tc_prof.py --d 768 >> $P 2>> gpurun_out/${TAG}_tc_prof.err
timeout 120 python scripts/tc_prof.py --precision exact >> $P 2>> gpurun_out/${TAG}_tc_prof.err
timeout 120 python scripts/tc_prof.py --n 10000 >> $P 2>> gpurun_out/${TAG}_tc_prof.err
echo "tc_prof lines: $(wc -l < $P)" >> $S
cut -c1-400 $P >> $S
timeout 600 python bench.py --steps 10 --cpu-seconds 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit=$?" >> $S; tail -n 3 gpurun_out/${TAG}_bench.err >> $S
timeout 600 python scripts/bench_extra.py c2 c35 > gpurun_out/${TAG}_bench_extra.jsonl 2> gpurun_out/${TAG}_bench_extra.err
echo "bench_extra exit=$?" >> $S; cut -c1-300 gpurun_out/${TAG}_bench_extra.jsonl >> $S
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches_cam.csv python scripts/cam_once.py \
    > gpurun_out/${TAG}_ncu_launches_cam.log 2>&1
echo "ncu cam launches exit=$?" >> $S
cat $S
