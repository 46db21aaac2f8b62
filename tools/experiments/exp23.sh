#!/bin/bash
O=gpurun_out/e23; mkdir -p $O
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/pytest.txt
echo "== gemm timeline"; timeout 200 python tools/gemm_timeline.py 512 2>&1 | tail -5 | tee $O/gemm_timeline_512.txt
echo "== microbench gemm"; timeout 600 python tools/microbench.py --m 128,512,1024,2048 --layouts GPTQ --iters 50 2>&1 | tee $O/mb_gemm.log | cut -c1-170
echo "== timeline (graph) lean"; for sh in 4096x4096 4096x11008; do timeout 200 python tools/timeline.py --layout GEMM --shape $sh --launches 6 2>&1 | tail -7 | tee -a $O/timeline_lean.txt; done
echo "== microbench lean"; timeout 300 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM 2>&1 | tee $O/mb_lean.log | cut -c1-170
for v in "" "B200Q_ST_TARGET=148" "B200Q_ST_TARGET=200" "B200Q_ST_TARGET=250" "B200Q_ST_TARGET=200 B200Q_ST_DEPTH=4" "B200Q_ST_TARGET=250 B200Q_ST_DEPTH=4" "B200Q_BENCH_NO_GROUP=1 B200Q_ST_TARGET=200"; do
  echo "== bench [$v]"; env $v timeout 600 python bench.py --no-cpu --no-prefill --steps 30 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(round(d['value'],1), 'tok/s', round(d['ms_per_step'],3), 'ms  frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['config'].get('launches_per_step'))
    except Exception as e: print('ERR', l[:300])
" | tee -a $O/bench_variants.txt
done
