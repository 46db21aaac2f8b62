#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
echo "== gemm tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm or tc or prefill" 2>&1 | tail -5
echo "== gemm timeline M=512"; timeout 120 python tools/gemm_timeline.py 512 2>&1 | tee $O/gemm_timeline9.log | tail -30
echo "== gemm microbench"; timeout 600 python tools/microbench.py --m 64,512,2048,8192 --layouts GPTQ --iters 50 2>&1 | tee $O/mb9_gemm.log
echo "== bench"; timeout 600 python bench.py 2>&1 | tee $O/bench9.json | cut -c1-1500
