#!/bin/bash
# Round evidence: tests, smoke, bench (both arms), ncu launch list of the bench command, ncu --set full of the top kernels.
R=${1:-r1}; O=gpurun_out/$R; mkdir -p $O
( nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/clocks.csv ) & SMI=$!
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $O/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee $O/smoke.txt
echo "== bench"; timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 2500 $O/bench.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 1200 $O/bench_reference.json
kill $SMI
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemv|gemm_tc" -s 448 -c 448 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.log 2>&1
echo "== ncu full: decode kernels (AWQ, bench shapes)"
for sh in 4096x4096 4096x11008 11008x4096; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemv_rp|gemv_fma" -s 3 -c 1 -o $O/ncu_decode_awq_$sh -f python tools/microbench.py --m 1 --iters 1 --layouts GEMM --shapes $sh > $O/ncu_decode_awq_$sh.log 2>&1
done
echo "== ncu full: tcgen05 GEMM M=512 / M=8192"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o $O/ncu_gemm512 -f python tools/microbench.py --m 512 --iters 2 --layouts GPTQ --shapes 4096x4096 > $O/ncu_gemm512.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o $O/ncu_gemm8192 -f python tools/microbench.py --m 8192 --iters 2 --layouts GPTQ --shapes 4096x4096 > $O/ncu_gemm8192.log 2>&1
echo "== microbench tables"
timeout 600 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM,GPTQ,MARLIN 2>&1 | tee $O/mb_decode.log | cut -c1-220
timeout 600 python tools/microbench.py --m 64,512,2048,8192 --layouts GPTQ --iters 50 > $O/mb_gemm.log 2>&1
ls -la $O
