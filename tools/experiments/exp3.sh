#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest3.log; echo "== pytest all: $(tail -1 $O/pytest3.log)"; grep -E "FAILED|Error" $O/pytest3.log | head
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== gemm microbench"; timeout 300 python tools/microbench.py --m 16,64,128,512,2048,8192 --iters 30 --layouts GPTQ --shapes 4096x4096,4096x11008,11008x4096 2>&1 | tee $O/mb3_gemm.log
for kb in 40 80; do
  echo "== decode v3 slice_kb=$kb"; B200Q_SLICE_KB=$kb timeout 300 python tools/microbench.py --m 1 --graph --iters 400 2>&1 | tee $O/mb3_v3_s$kb.log
done
echo "== decode v3 M=2,4,8"; timeout 300 python tools/microbench.py --m 2,4,8 --graph --iters 200 --layouts GEMM,GPTQ --shapes 4096x4096 2>&1 | tee $O/mb3_m.log
echo "== bench.py"; timeout 900 python bench.py --steps 30 --warmup 5 2>&1 | tail -3 | tee $O/bench3.log
echo "== bench.py reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $O/bench3_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-prefill > $O/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_rp -s 40 -c 3 -o $O/ncu3_decode -f python tools/microbench.py --m 1 --iters 8 --layouts GEMM --shapes 4096x4096,4096x11008 > $O/ncu3_decode.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 2 -o $O/ncu3_gemm -f python tools/microbench.py --m 512 --iters 4 --layouts GPTQ --shapes 4096x4096 > $O/ncu3_gemm.log 2>&1
ls $O | wc -l
