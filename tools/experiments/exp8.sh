#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
echo "== gemm timeline M=512"; timeout 120 python tools/gemm_timeline.py 512 2>&1 | tee $O/gemm_timeline.log
echo "== gemm timeline M=64"; timeout 120 python tools/gemm_timeline.py 64 2>&1 | tee -a $O/gemm_timeline.log
echo "== ref bench"; timeout 600 python tools/ref_bench.py 2>&1 | tee $O/ref_bench.log | cut -c1-400
echo "== decode"; timeout 300 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM,GPTQ 2>&1 | tee $O/mb8_dec.log
timeout 600 python -m pytest tests/test_vs_reference_cuda.py -m gpu -q 2>&1 | tail -3
# big fake layers for solid per-instruction stall sampling
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_rp -s 2 -c 1 -o $O/ncu8_rp_big -f python tools/microbench.py --m 1 --iters 2 --layouts GEMM --shapes 4096x65536 > $O/ncu8_rp_big.log 2>&1
B200Q_FORCE_FMA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_fma -s 2 -c 1 -o $O/ncu8_fma_big -f python tools/microbench.py --m 1 --iters 2 --layouts GEMM --shapes 4096x65536 > $O/ncu8_fma_big.log 2>&1
