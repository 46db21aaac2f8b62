#!/bin/bash
O=gpurun_out/e52; mkdir -p $O
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -30 | tee $O/pytest.txt
echo "== 3-bit HQQ g64 Mixtral shapes"
timeout 100 python tools/microbench.py --layouts HQQ --bits 3 --group 64 --m 1,16,64,512 --iters 30 --shapes 4096x14336,14336x4096 2>&1 | cut -c1-200 | tee -a $O/mb_3bit.jsonl
