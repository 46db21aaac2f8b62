#!/bin/bash
O=gpurun_out/e38; mkdir -p $O
for v in "B200Q_SYNC_FLAGS=0" "B200Q_SYNC_FLAGS=16"; do
echo "== timeline N=2 [$v]"; env $v timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_timeline.py --blocks 2 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -9 | tee -a $O/timeline.txt
done
