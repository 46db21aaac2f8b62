#!/bin/bash
O=gpurun_out/r2_10; mkdir -p $O
for n in 2; do
  echo "== bench N=$n"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 50 --warmup 5 2>&1 | tail -2 | cut -c1-1500 | tee $O/bench_n$n.txt
  echo "== bench act13b N=$n"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --config act13b --gpus $n --steps 20 --warmup 5 2>&1 | tail -2 | cut -c1-1200 | tee $O/bench_act13b_n$n.txt
  echo "== bench mixtral N=$n"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --config mixtral --gpus $n --steps 20 --warmup 5 2>&1 | tail -2 | cut -c1-1200 | tee $O/bench_mixtral_n$n.txt
  echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $n --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-900 | tee $O/bench_ref_n$n.txt
done
