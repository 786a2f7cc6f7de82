#!/bin/bash
# Multi-GPU run: NCCL parity + scaling bench at N = 1 .. #GPUs
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
S=gpurun_out/summary_multi.txt
echo "gpus=$NG" > $S
# QUICK=1: only the largest N, no parity test, no un-graphed comparison (8-GPU time is charged 8x)
QUICK=${QUICK:-0}
if [ $QUICK -eq 0 ]; then
  timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q --tb=short > gpurun_out/test_multi.log 2>&1
  echo "test_multigpu exit=$?" >> $S; tail -n 3 gpurun_out/test_multi.log >> $S
fi
for n in 1 2 4 8; do
  [ $n -le $NG ] || continue
  [ $QUICK -eq 0 ] || [ $n -eq $NG ] || continue
  if [ $n -eq 1 ]; then
    timeout 200 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port 29531 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline \
        > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
    echo "bench n=$n exit=$?" >> $S
    [ $QUICK -eq 0 ] && timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port 29532 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-graph \
        > gpurun_out/scale_n${n}_nograph.json 2> gpurun_out/scale_n${n}_nograph.err
  fi
  echo "bench n=$n exit=$?" >> $S
done
# BASELINE config 5 (10k x 1M x 512 top-k, gallery-sharded) and config 4 (D = 768) on all GPUs
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
    --master-port 29533 scripts/dist_topk_bench.py > gpurun_out/c5_topk_n$NG.json 2> gpurun_out/c5_topk_n$NG.err
echo "c5 topk n=$NG exit=$?" >> $S
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
    --master-port 29534 bench.py --gpus $NG --steps 20 --warmup 3 --dim 768 --no-cpu-baseline --no-e2e \
    > gpurun_out/c4_d768_n$NG.json 2> gpurun_out/c4_d768_n$NG.err
echo "c4 d768 n=$NG exit=$?" >> $S
cat $S
