#!/bin/bash
O=gpurun_out/e25; mkdir -p $O
echo "== pytest decode"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode or stream or sibling or group or full_size or workspace or strided" 2>&1 | tail -4 | tee $O/pytest.txt
echo "== timeline (graph) lean"; for sh in 4096x4096 4096x11008; do timeout 200 python tools/timeline.py --layout GEMM --shape $sh --launches 5 2>&1 | tail -6 | tee -a $O/timeline_lean.txt; done
echo "== microbench lean"; timeout 300 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM 2>&1 | tee $O/mb_lean.log | cut -c1-150
for v in "B200Q_ST_TARGET=120" "B200Q_ST_TARGET=148" "B200Q_ST_TARGET=200" "B200Q_ST_TARGET=250" "B200Q_ST_TARGET=296" "B200Q_ST_TARGET=148 B200Q_ST_DEPTH=4" "B200Q_ST_TARGET=148 B200Q_CARVEOUT=0"; do
  echo "== bench [$v]"; env $v timeout 600 python bench.py --no-cpu --no-prefill --steps 30 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(round(d['value'],1), 'tok/s', round(d['ms_per_step'],3), 'ms  frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['config'].get('launches_per_step'))
    except Exception as e: print('ERR', l[:300])
" | tee -a $O/bench_variants.txt
done
