#!/bin/bash
O=gpurun_out/e22; mkdir -p $O
echo "== pytest gemm"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tcgen05 or prefill or smoke" 2>&1 | tail -5 | tee $O/pytest_gemm.txt
echo "== gemm timeline"; timeout 200 python tools/gemm_timeline.py 512 2>&1 | tail -12 | tee $O/gemm_timeline_512.txt
echo "== microbench gemm"; timeout 600 python tools/microbench.py --m 64,128,256,512,1024,2048,8192 --layouts GPTQ --iters 50 2>&1 | tee $O/mb_gemm.log | cut -c1-200
