#!/bin/bash
# N=2: fused hand-off vs NCCL all-gather; virtual-rank tests; N=1 regression check (run with gpurun --gpus 2)
O=gpurun_out/e36; mkdir -p $O
echo "== pytest sharding"; timeout 600 python -m pytest tests/test_sharding.py -m gpu -q -x 2>&1 | tail -5 | tee $O/pytest_sharding.txt
fmt='
import sys, json
for l in sys.stdin:
    l = l.strip()
    if not l.startswith("{"): continue
    try:
        d = json.loads(l); c = d["config"]; print(round(d["value"],1), "tok/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"],1), "launches", c.get("launches_per_step"), "|", c.get("parallelism"), "| eq", c.get("replicas_equal"), "timeouts", c.get("peer_wait_timeouts"), "err", c.get("fused_sharded_error"), "finite", c.get("outputs_finite"))
    except Exception as e: print("ERR", l[:300])
'
for v in "" "B200Q_BENCH_NCCL=1"; do
echo "== N=2 [$v]"; env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > $O/n2.log 2>&1; tail -4 $O/n2.log | cut -c1-400 | grep -v "^{" ; python -c "$fmt" < $O/n2.log | tee -a $O/scale.txt
done
echo "== N=1"; timeout 600 python bench.py --no-cpu --no-prefill --steps 30 2>&1 | tail -1 | python -c "$fmt" | tee -a $O/scale.txt
