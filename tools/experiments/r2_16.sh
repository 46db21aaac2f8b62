#!/bin/bash
# bisect of the 2 % decode gap between the round-1 and round-2 libraries: A = round-2 tree with round-1's decode kernel
# files, C = round-1 tree with the parameter block padded to the round-2 size
O=gpurun_out/r2_16; mkdir -p $O
for i in 1 2; do
  for v in new r1 mixA mixC; do
    lib=$PWD/qllm_b200/libb200q_$v.so; [ $v = new ] && lib=$PWD/qllm_b200/libb200q.so
    echo "== $v"; B200Q_LIB=$lib timeout 300 python bench.py --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | cut -c1-130 | tee -a $O/ab_$v.txt
  done
done
