#!/bin/bash
O=gpurun_out/e49; mkdir -p $O
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -20 | tee $O/pytest.txt
for v in "B200Q_GEMM_SPLITK=1" "B200Q_GEMM_SPLITK=0"; do
echo "== small-M GEMM [$v]"
env $v timeout 100 python tools/microbench.py --layouts GPTQ --bits 4 --group 128 --m 16,64 --iters 30 --shapes 4096x4096,4096x11008,11008x4096 2>&1 | cut -c1-200 | tee -a $O/mb_splitk.jsonl
env $v timeout 100 python tools/microbench.py --layouts HQQ --bits 4 --group 64 --m 16,64 --iters 30 --shapes 4096x14336,14336x4096 2>&1 | cut -c1-200 | tee -a $O/mb_splitk.jsonl
done
