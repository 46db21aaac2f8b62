#!/bin/bash
O=gpurun_out/e45; mkdir -p $O
fmt='
import sys, json
for l in sys.stdin:
    l = l.strip()
    if not l.startswith("{"): continue
    try:
        d = json.loads(l); print(round(d["value"],1), "tok/s", round(d["ms_per_step"],3), "ms")
    except Exception as e: print("ERR", l[:300])
'
for v in "" "B200Q_IM_TARGET=288" "B200Q_IM_TARGET=288 B200Q_IM_TPC=2" "B200Q_IM_TPC=2" "B200Q_IM_TARGET=256" "B200Q_IM_TARGET=320" "B200Q_IM_DEPTH=4" "B200Q_IM_TARGET=288 B200Q_IM_DEPTH=4"; do
echo "== [$v]"; env $v timeout 120 python bench.py --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | python -c "$fmt" | tee -a $O/sweep.txt
done
