#!/bin/bash
O=gpurun_out/r2_12; mkdir -p $O
echo "== baseline"; timeout 300 python bench.py --no-cpu --no-prefill --steps 50 2>&1 | tail -1 | cut -c1-220 | tee $O/bench_base.txt
for spec in 8,1 16,1 32,1 64,1 16,2 32,2 148,1; do
  echo "== L2PF $spec"; B200Q_BENCH_L2PF=$spec timeout 300 python bench.py --no-cpu --no-prefill --steps 50 2>&1 | tail -1 | cut -c1-220 | tee $O/bench_pf_$spec.txt
done
