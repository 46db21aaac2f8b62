#!/bin/bash
# (1) parameter-block size threshold; (2) the cost of the NaN/Inf detection in the decode kernel's digit stage:
#     nan0 none (round 1), new = max.NaN chain, nan2 = one vote beside the fmaxf chain, nan3 = integer maximum of |x| bit patterns
O=gpurun_out/r2_17; mkdir -p $O
timeout 120 tools/ubench/param_size_probe | tee $O/param_size_probe.txt
for i in 1 2; do
  for v in new nan0 nan2 nan3 mixC; do
    lib=$PWD/qllm_b200/libb200q_$v.so; [ $v = new ] && lib=$PWD/qllm_b200/libb200q.so
    echo "== $v"; B200Q_LIB=$lib timeout 300 python bench.py --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | cut -c1-130 | tee -a $O/ab_$v.txt
  done
done
for v in nan2 nan3; do B200Q_LIB=$PWD/qllm_b200/libb200q_$v.so timeout 600 python -m pytest tests/test_round2.py -m gpu -q -x -k "nan or inf or NaN" 2>&1 | tail -3; done
