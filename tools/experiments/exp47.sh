#!/bin/bash
O=gpurun_out/e47; mkdir -p $O
echo "== pytest sharding"; timeout 600 python -m pytest tests/test_sharding.py -m gpu -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -20 | tee $O/pytest.txt
echo "== Mixtral expert shapes, HQQ (configs[4])"
for bg in "4 64" "4 128" "2 64" "8 128" "3 64"; do set -- $bg
timeout 200 python tools/microbench.py --layouts HQQ --bits $1 --group $2 --m 1 --graph --iters 200 --shapes 4096x14336,14336x4096 2>&1 | cut -c1-200 | tee -a $O/mb_mixtral_hqq.jsonl
timeout 200 python tools/microbench.py --layouts HQQ --bits $1 --group $2 --m 16,64,512 --iters 30 --shapes 4096x14336,14336x4096 2>&1 | cut -c1-200 | tee -a $O/mb_mixtral_hqq.jsonl
done
