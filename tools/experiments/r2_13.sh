#!/bin/bash
O=gpurun_out/r2_13; mkdir -p $O
for i in 1 2 3; do
  echo "== new lib"; timeout 300 python bench.py --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | cut -c1-160 | tee -a $O/ab_new.txt
  echo "== r1 lib"; B200Q_LIB=$PWD/qllm_b200/libb200q_r1.so timeout 300 python bench.py --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | cut -c1-160 | tee -a $O/ab_r1.txt
done
echo "== marlin probe"; timeout 900 python tools/marlin_ref_probe.py 2>&1 | tee $O/marlin_ref_probe.txt
