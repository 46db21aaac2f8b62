#!/bin/bash
O=gpurun_out/e51; mkdir -p $O
for v in "B200Q_GEMM_SPLITK=1" "B200Q_GEMM_SPLITK=0"; do
echo "== small-M GEMM [$v]"
env $v timeout 100 python tools/microbench.py --layouts GPTQ --bits 4 --group 128 --m 9,16,32,64 --iters 40 --shapes 4096x4096,11008x4096,8192x8192,5120x5120 2>&1 | cut -c1-120 | tee -a $O/mb_splitk_ab.jsonl
done
