#!/bin/bash
# (1) A/B of the settled decode kernel (vote-based NaN check, parameter block back at 1560 B) against round 1;
# (2) token-tile x split-K sweep of the tcgen05 GEMM
O=gpurun_out/r2_19; mkdir -p $O
for v in new r1; do
  lib=$PWD/qllm_b200/libb200q_$v.so; [ $v = new ] && lib=$PWD/qllm_b200/libb200q.so
  echo "== $v"; B200Q_LIB=$lib timeout 300 python bench.py --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | cut -c1-130 | tee -a $O/ab_$v.txt
done
timeout 600 python -m pytest tests/test_round2.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python tools/sweep_gemm_tiles.py 2>&1 | tee $O/sweep_gemm_tiles.jsonl
