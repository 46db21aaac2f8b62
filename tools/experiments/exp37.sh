#!/bin/bash
# N=2 hand-off cost probes (run with gpurun --gpus 2)
O=gpurun_out/e37; mkdir -p $O
fmt='
import sys, json
for l in sys.stdin:
    l = l.strip()
    if not l.startswith("{"): continue
    try:
        d = json.loads(l); c = d["config"]; print(round(d["value"],1), "tok/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"],1), "launches", c.get("launches_per_step"), "| eq", c.get("replicas_equal"), "timeouts", c.get("peer_wait_timeouts"), "err", c.get("fused_sharded_error"))
    except Exception as e: print("ERR", l[:300])
'
for v in "B200Q_SYNC_FLAGS=0" "B200Q_SYNC_FLAGS=7" "B200Q_SYNC_FLAGS=6" "B200Q_SYNC_FLAGS=1" "B200Q_SYNC_FLAGS=2" "B200Q_SYNC_FLAGS=4" "B200Q_SYNC_FLAGS=8" "B200Q_SYNC_FLAGS=16"; do
echo "== N=2 [$v]"; env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/n2.log 2>&1; python -c "$fmt" < $O/n2.log | tee -a $O/probes.txt
done
