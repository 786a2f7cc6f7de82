#!/bin/bash
# NCCL parity + one bench line on all GPUs of the box (usage: gpurun --gpus N -- bash scripts/gpu_multi_quick.sh [tag])
set -u
TAG=${1:-multi}
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
S=gpurun_out/summary_$TAG.txt
echo "gpus=$NG" > $S
timeout 500 python -m pytest tests/test_multigpu.py -m gpu -q --tb=short > gpurun_out/${TAG}_test_multi.log 2>&1
echo "test_multigpu exit=$?" >> $S; tail -n 12 gpurun_out/${TAG}_test_multi.log >> $S
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
    --master-port 29531 bench.py --gpus $NG --steps 20 --warmup 3 \
    > gpurun_out/${TAG}_scale_n$NG.json 2> gpurun_out/${TAG}_scale_n$NG.err
echo "bench n=$NG exit=$?" >> $S
tail -n 5 gpurun_out/${TAG}_scale_n$NG.err >> $S
VTC_PHASE_TIMING=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
    --master-port 29532 bench.py --gpus $NG --steps 5 --warmup 3 --no-e2e --no-extra \
    > gpurun_out/${TAG}_phase_n$NG.json 2> gpurun_out/${TAG}_phase_n$NG.err
grep phases_ms gpurun_out/${TAG}_phase_n$NG.err | tail -n 3 >> $S
cut -c1-1500 gpurun_out/${TAG}_scale_n$NG.json >> $S
cat $S
