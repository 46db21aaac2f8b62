#!/bin/bash
# multi-GPU bench (run with gpurun --gpus N): N = number of visible GPUs
mkdir -p gpurun_out; O=gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "== visible GPUs: $N"
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 900 python bench.py --gpus 1 --steps 30 --warmup 5 --no-prefill 2>&1 | tail -1 | tee $O/scale_n$n.log | cut -c1-260
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 30 --warmup 5 2>&1 | tail -3 | tee $O/scale_n$n.log | cut -c1-260
    fi
  fi
done
