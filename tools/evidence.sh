#!/bin/bash
# Round evidence: tests, smoke, bench (both arms), ncu launch list of the bench command, ncu --set full of the top kernels.
R=${1:-r1}; O=gpurun_out/$R; mkdir -p $O
( nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/clocks.csv ) & SMI=$!
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $O/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee $O/smoke.txt
echo "== bench"; timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 3000 $O/bench.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 1200 $O/bench_reference.json
kill $SMI
echo "== ncu launch list of the bench command (one step = 128 launches after the eager pass and 3 warm-up replays)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemv|gemm_tc" -s 512 -c 128 --csv --log-file $O/launches_bench.csv python bench.py --no-cpu --no-prefill --steps 2 --warmup 3 > $O/bench_under_ncu.log 2>&1
echo "== ncu full: the four launches of one decoder block inside the bench (qkv, o, gate|up, down)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemv_imma" -s 132 -c 4 -o $O/ncu_decode_block -f python bench.py --no-cpu --no-prefill --steps 1 --warmup 3 > $O/ncu_decode_block.log 2>&1
echo "== ncu full: tcgen05 GEMM M=512 / M=8192"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o $O/ncu_gemm512 -f python tools/microbench.py --m 512 --iters 2 --layouts GPTQ --shapes 4096x4096 > $O/ncu_gemm512.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o $O/ncu_gemm8192 -f python tools/microbench.py --m 8192 --iters 2 --layouts GPTQ --shapes 4096x4096 > $O/ncu_gemm8192.log 2>&1
echo "== microbench tables"
timeout 600 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM,GPTQ,MARLIN 2>&1 | tee $O/mb_decode.log | cut -c1-220
timeout 600 python tools/microbench.py --m 64,512,2048,8192 --layouts GPTQ --iters 50 > $O/mb_gemm.log 2>&1; cut -c1-200 $O/mb_gemm.log
echo "== reference CUDA kernels vs the engine"
timeout 600 python tools/ref_bench.py > $O/ref_vs_ours.jsonl 2> $O/ref_vs_ours.err; tail -12 $O/ref_vs_ours.jsonl | cut -c1-250
ls -la $O
