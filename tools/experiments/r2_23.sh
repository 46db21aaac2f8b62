#!/bin/bash
# vectorised split-K fix-up (M <= 64) + default bench with the graph-replayed prefill leg
O=gpurun_out/r2_23; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size_parity.py tests/test_round2.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/microbench.py --layouts GPTQ --m 16,32,64,128 --force gemm --graph 2>&1 | tee $O/mb_gemm_small_m.jsonl | cut -c1-200
timeout 600 python tools/microbench.py --layouts HQQ --bits 4 --group 64 --shapes 4096x14336,14336x4096 --m 16,64 --graph 2>&1 | tee -a $O/mb_gemm_small_m.jsonl | cut -c1-200
echo "== default"; timeout 600 python bench.py --no-cpu --steps 50 2>&1 | tail -1 | tee $O/bench_n1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d.get('prefill'))"
