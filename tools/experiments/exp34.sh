#!/bin/bash
# next-layer L2 prefetch: parity, then bench with / without
O=gpurun_out/e34; mkdir -p $O
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $O/pytest.txt
for v in "" "B200Q_BENCH_NO_CHAIN=1" "B200Q_IM_TARGET=222" "B200Q_IM_TARGET=444"; do
  echo "== bench [$v]"; env $v timeout 600 python bench.py --no-cpu --no-prefill --steps 30 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(round(d['value'],1), 'tok/s', round(d['ms_per_step'],3), 'ms  frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['config'].get('launches_per_step'))
    except Exception as e: print('ERR', l[:300])
" | tee -a $O/bench_variants.txt
done
echo "== bench timeline"; timeout 300 python tools/bench_timeline.py --blocks 3 2>&1 | tail -13 | tee $O/bench_timeline.txt
