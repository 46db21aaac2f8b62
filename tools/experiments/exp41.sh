#!/bin/bash
O=gpurun_out/e41; mkdir -p $O
fmt='
import sys, json
for l in sys.stdin:
    l = l.strip()
    if not l.startswith("{"): continue
    try:
        d = json.loads(l); c = d["config"]; print(round(d["value"],1), "tok/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"],1), "|", c.get("parallelism"), "| eq", c.get("replicas_equal"), "timeouts", c.get("peer_wait_timeouts"), "err", c.get("fused_sharded_error"), "clk", d["clocks"])
    except Exception as e: print("ERR", l[:300])
'
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $O/pytest.txt
for v in "" "B200Q_LIB=/root/repo/qllm_b200/libb200q_base.so"; do
echo "== N=1 [$v]"; env $v timeout 300 python bench.py --no-cpu --no-prefill 2>&1 | tail -1 | python -c "$fmt" | tee -a $O/ab.txt
done
echo "== N=2"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/n2.log 2>&1; grep -i "error\|Traceback" $O/n2.log | head -5; python -c "$fmt" < $O/n2.log | tee -a $O/ab.txt
echo "== gemm timeline M=512"; timeout 120 python tools/gemm_timeline.py 512 2>&1 | tail -9 | tee $O/gemm_timeline_m512.txt
