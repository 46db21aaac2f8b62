#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
echo "== gemm tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm or tc or prefill" 2>&1 | tail -5
echo "== gemm timeline M=512"; timeout 120 python tools/gemm_timeline.py 512 2>&1 | tee $O/gemm_timeline11.log | tail -12
echo "== gemm microbench"; timeout 600 python tools/microbench.py --m 64,512,2048,8192 --layouts GPTQ --iters 50 2>&1 | tee $O/mb11_gemm.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o $O/ncu11_gemm512 -f python tools/microbench.py --m 512 --iters 2 --layouts GPTQ --shapes 4096x4096 > $O/ncu11_gemm512.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o $O/ncu11_gemm8192 -f python tools/microbench.py --m 8192 --iters 2 --layouts GPTQ --shapes 4096x4096 > $O/ncu11_gemm8192.log 2>&1
