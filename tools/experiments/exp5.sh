#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
echo "== mma probe"; timeout 120 tools/ubench/mma_probe 2>&1 | tee $O/mma_probe.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest5.log; echo "== pytest all (fma default): $(tail -1 $O/pytest5.log)"; grep -E "FAILED|Error" $O/pytest5.log | head
echo "== decode default (fma for M<=2)"; timeout 300 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM,GPTQ 2>&1 | tee $O/mb5_fma.log
echo "== decode M=2 fma"; timeout 300 python tools/microbench.py --m 2 --graph --iters 400 --layouts GEMM,GPTQ --shapes 4096x4096 2>&1 | tee $O/mb5_fma_m2.log
echo "== decode rp smem (v3)"; B200Q_GEMV=v3 timeout 300 python tools/microbench.py --m 1 --graph --iters 400 2>&1 | tee $O/mb5_v3.log
echo "== timelines (rp smem variant)"; B200Q_GEMV=v3 timeout 120 python tools/timeline.py --layout GEMM --shape 4096x4096 --launches 5 2>&1 | tee -a $O/timeline2.log
B200Q_GEMV=v3 timeout 120 python tools/timeline.py --layout GPTQ --shape 4096x4096 --launches 5 2>&1 | tee -a $O/timeline2.log
echo "== bench.py"; timeout 900 python bench.py --steps 30 --warmup 5 --no-prefill 2>&1 | tail -1 | tee $O/bench5.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_fma -s 10 -c 2 -o $O/ncu5_fma -f python tools/microbench.py --m 1 --iters 8 --layouts GEMM --shapes 4096x4096 > $O/ncu5_fma.log 2>&1
echo "== gemm microbench (deferred publish)"; timeout 300 python tools/microbench.py --m 64,512,2048,8192 --iters 30 --layouts GPTQ --shapes 4096x4096,4096x11008 2>&1 | tee $O/mb5_gemm.log
timeout 300 python -m pytest tests -m gpu -q -k tcgen05 2>&1 | tail -3
