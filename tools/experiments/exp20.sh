#!/bin/bash
# streaming decode kernel: parity, per-shape timing vs the whole-slice kernels, bench variants, timeline
O=gpurun_out/e20; mkdir -p $O
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $O/pytest.txt
echo "== microbench stream"; timeout 300 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM,GPTQ,MARLIN 2>&1 | tee $O/mb_stream.log | cut -c1-200
echo "== microbench legacy"; B200Q_GEMV=rp timeout 300 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM,GPTQ 2>&1 | tee $O/mb_legacy.log | cut -c1-200
echo "== microbench stream M=4"; timeout 300 python tools/microbench.py --m 4 --graph --iters 400 --layouts GEMM,GPTQ 2>&1 | tee $O/mb_stream_m4.log | cut -c1-200
for v in "" "B200Q_BENCH_NO_GROUP=1" "B200Q_GEMV=rp B200Q_BENCH_NO_GROUP=1" "B200Q_ST_TARGET=148" "B200Q_ST_TARGET=240" "B200Q_ST_RING_KB=32" "B200Q_ST_RING_KB=96" "B200Q_ST_DEPTH=4" ; do
  echo "== bench [$v]"; env $v timeout 600 python bench.py --no-cpu --no-prefill --steps 30 2>&1 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(round(d['value'],1), 'tok/s', round(d['ms_per_step'],3), 'ms  frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['config'].get('launches_per_step'))
    except Exception as e: print('ERR', l[:300])
" | tee -a $O/bench_variants.txt
done
echo "== timeline"; for sh in 4096x4096 4096x11008 11008x4096; do timeout 200 python tools/timeline.py --layout GEMM --shape $sh --launches 8 2>&1 | tail -9 | tee -a $O/timeline.txt; done
