#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest4.log; echo "== pytest all: $(tail -1 $O/pytest4.log)"; grep -E "FAILED|Error" $O/pytest4.log | head
echo "== timelines"; for sh in 4096x4096 4096x11008; do timeout 120 python tools/timeline.py --layout GEMM --shape $sh --launches 8 2>&1 | tee -a $O/timeline.log; done
timeout 120 python tools/timeline.py --layout GPTQ --shape 4096x4096 --launches 8 2>&1 | tee -a $O/timeline.log
echo "== gemm microbench"; timeout 300 python tools/microbench.py --m 64,512,2048,8192 --iters 30 --layouts GPTQ --shapes 4096x4096,4096x11008,11008x4096 2>&1 | tee $O/mb4_gemm.log
echo "== gemm TT128 only at large M"; B200Q_TT256_MIN_M=1000000 timeout 300 python tools/microbench.py --m 2048,8192 --iters 30 --layouts GPTQ --shapes 4096x4096 2>&1 | tee $O/mb4_gemm_tt128.log
echo "== gemm TT256 at 512"; B200Q_TT256_MIN_M=512 timeout 300 python tools/microbench.py --m 512 --iters 30 --layouts GPTQ --shapes 4096x4096,4096x11008 2>&1 | tee $O/mb4_gemm_tt256.log
echo "== decode"; timeout 300 python tools/microbench.py --m 1 --graph --iters 400 2>&1 | tee $O/mb4_dec.log
echo "== decode min_steps=4"; B200Q_MIN_STEPS=4 timeout 300 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM 2>&1 | tee $O/mb4_dec_ms4.log
echo "== bench.py"; timeout 900 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | tee $O/bench4.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 2 -o $O/ncu4_gemm -f python tools/microbench.py --m 512 --iters 4 --layouts GPTQ --shapes 4096x4096 > $O/ncu4_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_rp -s 40 -c 2 -o $O/ncu4_decode -f python tools/microbench.py --m 1 --iters 8 --layouts GEMM --shapes 4096x4096 > $O/ncu4_decode.log 2>&1
