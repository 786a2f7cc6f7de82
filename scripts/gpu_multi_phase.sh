#!/bin/bash
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
VTC_PHASE_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
    --master-port 29541 bench.py --gpus $NG --steps 6 --warmup 3 --no-cpu-baseline --no-e2e \
    > gpurun_out/phase_n$NG.json 2> gpurun_out/phase_n$NG.err
grep phases_ms gpurun_out/phase_n$NG.err | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
    --master-port 29543 bench.py --gpus $NG --steps 20 --warmup 3 --no-cpu-baseline \
    > gpurun_out/scale2_n$NG.json 2> gpurun_out/scale2_n$NG.err
tail -c 600 gpurun_out/scale2_n$NG.json
