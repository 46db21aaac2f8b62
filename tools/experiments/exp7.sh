#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest7.log; echo "== pytest all: $(tail -1 $O/pytest7.log)"; grep -E "FAILED|Error" $O/pytest7.log | head
echo "== gemm microbench (8 stages)"; timeout 300 python tools/microbench.py --m 64,512,2048,8192 --iters 30 --layouts GPTQ --shapes 4096x4096,4096x11008,11008x4096 2>&1 | tee $O/mb7_gemm.log
echo "== decode (dispatch)"; timeout 300 python tools/microbench.py --m 1 --graph --iters 400 2>&1 | tee $O/mb7_dec.log
echo "== bench.py"; timeout 900 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | tee $O/bench7.log | cut -c1-200
echo "== ref bench"; timeout 600 python tools/ref_bench.py 2>&1 | tee $O/ref_bench.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemv|gemm_tc" -c 700 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-prefill > $O/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_rp -s 10 -c 2 -o $O/ncu7_decode -f python tools/microbench.py --m 1 --iters 8 --layouts GEMM --shapes 4096x4096 > $O/ncu7_decode.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 2 -o $O/ncu7_gemm -f python tools/microbench.py --m 512 --iters 4 --layouts GPTQ --shapes 4096x4096 > $O/ncu7_gemm.log 2>&1
