#!/bin/bash
O=gpurun_out/r2_11; mkdir -p $O
echo "== bench N=2"; timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 2>&1 | tail -2 | cut -c1-1500 | tee $O/bench_n2.txt
