#!/bin/bash
# Multi-GPU run: NCCL parity + scaling bench at N = 1 .. #GPUs
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
S=gpurun_out/summary_multi.txt
echo "gpus=$NG" > $S
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q --tb=short > gpurun_out/test_multi.log 2>&1
echo "test_multigpu exit=$?" >> $S; tail -n 3 gpurun_out/test_multi.log >> $S
for n in 1 2 4 8; do
  [ $n -le $NG ] || continue
  if [ $n -eq 1 ]; then
    timeout 200 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port 29531 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline \
        > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
    echo "bench n=$n exit=$?" >> $S
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port 29532 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-graph \
        > gpurun_out/scale_n${n}_nograph.json 2> gpurun_out/scale_n${n}_nograph.err
  fi
  echo "bench n=$n exit=$?" >> $S
done
cat $S
