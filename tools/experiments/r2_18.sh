#!/bin/bash
# NaN/Inf detection variants of the decode kernel's digit stage, continued: nan4 = CTA flag in shared memory + poisoned partial sums
O=gpurun_out/r2_18; mkdir -p $O
for i in 1 2; do
  for v in nan0 nan2 nan4; do
    lib=$PWD/qllm_b200/libb200q_$v.so
    echo "== $v"; B200Q_LIB=$lib timeout 300 python bench.py --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | cut -c1-130 | tee -a $O/ab_$v.txt
  done
done
B200Q_LIB=$PWD/qllm_b200/libb200q_nan4.so timeout 600 python -m pytest tests/test_round2.py -m gpu -q -x 2>&1 | tail -3
