#!/bin/bash
# N=4 check of the fused sharded bench (run with gpurun --gpus 4)
O=gpurun_out/e42; mkdir -p $O
fmt='
import sys, json
for l in sys.stdin:
    l = l.strip()
    if not l.startswith("{"): continue
    try:
        d = json.loads(l); c = d["config"]; print(d["n_gpus"], "GPUs", round(d["value"],1), "tok/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"],1), "|", c.get("parallelism"), "| eq", c.get("replicas_equal"), "timeouts", c.get("peer_wait_timeouts"), "err", c.get("fused_sharded_error"))
    except Exception as e: print("ERR", l[:300])
'
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 > $O/n4.log 2>&1; grep -i "error\|Traceback" $O/n4.log | head -5; python -c "$fmt" < $O/n4.log | tee -a $O/scale.txt
