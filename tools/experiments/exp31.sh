#!/bin/bash
O=gpurun_out/e31; mkdir -p $O
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $O/pytest.txt
echo "== microbench imma"; timeout 300 python tools/microbench.py --m 1 --graph --iters 400 --layouts GPTQ 2>&1 | tee $O/mb_imma.log | cut -c1-150
echo "== timeline (graph) imma"; for sh in 4096x4096 4096x11008; do timeout 200 python tools/timeline.py --layout GPTQ --shape $sh --launches 5 2>&1 | tail -6 | tee -a $O/timeline_imma.txt; done
for v in "" "B200Q_IM_TARGET=260" "B200Q_IM_TARGET=330"; do
  echo "== bench [$v]"; env $v timeout 600 python bench.py --no-cpu --no-prefill --steps 30 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(round(d['value'],1), 'tok/s', round(d['ms_per_step'],3), 'ms  frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['config'].get('launches_per_step'))
    except Exception as e: print('ERR', l[:300])
" | tee -a $O/bench_variants.txt
done
