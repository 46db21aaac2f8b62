#!/bin/bash
# A/B on one box: prefetch v2 (after pdl_wait) on/off x 3-CTA/4-CTA builds
O=gpurun_out/e35; mkdir -p $O
fmt='
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(round(d["value"],1), "tok/s", round(d["ms_per_step"],3), "ms  frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1), "launches", d["config"].get("launches_per_step"))
    except Exception as e: print("ERR", l[:300])
'
for rep in 1 2; do
for v in "" "B200Q_BENCH_NO_CHAIN=1" "B200Q_LIB=/root/repo/qllm_b200/libb200q_4cta.so" "B200Q_LIB=/root/repo/qllm_b200/libb200q_4cta.so B200Q_BENCH_NO_CHAIN=1"; do
  echo "== bench [$v]"; env $v timeout 600 python bench.py --no-cpu --no-prefill --steps 30 2>&1 | tail -1 | python -c "$fmt" | tee -a $O/bench_variants.txt
done; done
echo "== bench timeline"; timeout 300 python tools/bench_timeline.py --blocks 3 2>&1 | tail -13 | tee $O/bench_timeline.txt
