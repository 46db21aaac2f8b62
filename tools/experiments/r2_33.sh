#!/bin/bash
# scaling point at N GPUs (argument): tagged hand-off with per-call step words vs the round-1 form
N=$1; O=gpurun_out/r2_33; mkdir -p $O
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 100 --warmup 5 2>&1 | tail -1; }
echo "== N=$N node epoch"; run 29611 | tee $O/bench_n$N.json | cut -c1-150
echo "== N=$N round-1 form"; B200Q_BENCH_NO_NODE_EPOCH=1 run 29612 | tee $O/bench_n${N}_r1form.json | cut -c1-150
