#!/bin/bash
O=gpurun_out/e48; mkdir -p $O
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -20 | tee $O/pytest.txt
echo "== Mixtral w2 shape, HQQ g64"
timeout 100 python tools/microbench.py --layouts HQQ --bits 4 --group 64 --m 16,64,512 --iters 30 --shapes 14336x4096 2>&1 | cut -c1-200 | tee -a $O/mb.jsonl
timeout 100 python tools/microbench.py --layouts GPTQ --bits 4 --group 128 --m 512 --iters 30 --shapes 4096x4096,11008x4096 2>&1 | cut -c1-200 | tee -a $O/mb.jsonl
