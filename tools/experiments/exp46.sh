#!/bin/bash
O=gpurun_out/e46; mkdir -p $O
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -20 | tee $O/pytest.txt
for v in "B200Q_GEMM_PDL=1" "B200Q_GEMM_PDL=0"; do
echo "== gemm microbench [$v]"; env $v timeout 300 python tools/microbench.py --m 512,2048 --layouts GPTQ --iters 50 2>&1 | cut -c1-190 | tee -a $O/mb_gemm_pdl.txt
echo "== bench prefill [$v]"; env $v timeout 300 python bench.py --no-cpu --steps 20 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['value'],1), 'tok/s | prefill', d.get('prefill'))" | tee -a $O/mb_gemm_pdl.txt
done
