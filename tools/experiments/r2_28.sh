#!/bin/bash
# N = 2: tagged hand-off with the per-call step words (no kernel-boundary wait ahead of tagged x) vs round 1's form
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 100 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('replicas_equal'), d.get('peer_wait_timeouts'), d['config'].get('parallelism', '')[:60])"; }
mkdir -p gpurun_out/r2_28
for i in 1 2; do
  echo "== node epoch"; run 2951$i | tee -a gpurun_out/r2_28/n2_node_epoch.txt
  echo "== round-1 form"; B200Q_BENCH_NO_NODE_EPOCH=1 run 2952$i | tee -a gpurun_out/r2_28/n2_epoch_word.txt
done
