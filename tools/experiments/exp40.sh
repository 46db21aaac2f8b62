#!/bin/bash
O=gpurun_out/e40; mkdir -p $O
fmt='
import sys, json
for l in sys.stdin:
    l = l.strip()
    if not l.startswith("{"): continue
    try:
        d = json.loads(l); c = d["config"]; print(round(d["value"],1), "tok/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"],1), "clk", d["clocks"])
    except Exception as e: print("ERR", l[:300])
'
for rep in 1 2; do for v in "" "B200Q_LIB=/root/repo/qllm_b200/libb200q_base.so"; do
echo "== N=1 [$v]"; env $v timeout 300 python bench.py --no-cpu --no-prefill 2>&1 | tail -1 | python -c "$fmt" | tee -a $O/ab.txt
done; done
echo "== gemm timeline M=512"; timeout 120 python tools/gemm_timeline.py 512 2>&1 | tail -40 | tee $O/gemm_timeline_m512.txt
