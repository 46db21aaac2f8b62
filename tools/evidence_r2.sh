#!/bin/bash
# Round-2 evidence: tests, smoke, bench (all configs, both arms), ncu launch list + full captures, sanitizer, reference-Marlin probe.
R=${1:-r2}; O=gpurun_out/$R; mkdir -p $O
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
echo "== bench"; timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench.err; tail -c 2500 $O/bench_n1.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_n1.json 2> $O/bench_reference.err; tail -c 900 $O/bench_reference_n1.json
for c in prefill7b act13b mixtral; do echo "== bench $c"; timeout 900 python bench.py --config $c --steps 20 > $O/bench_$c.json 2> $O/bench_$c.err; tail -c 600 $O/bench_$c.json; done
echo "== bench prefill7b M=2048"; timeout 900 python bench.py --config prefill7b --m 2048 --steps 5 > $O/bench_prefill7b_m2048.json 2>/dev/null; tail -c 500 $O/bench_prefill7b_m2048.json
echo "== bench prefill7b, per-layer calls"; B200Q_BENCH_NO_GROUP=1 timeout 900 python bench.py --config prefill7b --steps 20 --no-cpu 2>/dev/null | tail -1 > $O/bench_prefill7b_nogroup.json; cut -c1-200 $O/bench_prefill7b_nogroup.json
echo "== bench decode, tagged hand-off on one GPU"; timeout 600 python bench.py --handoff tagged --no-cpu --no-prefill --steps 100 2>/dev/null | tail -1 > $O/bench_tagged_n1.json; cut -c1-200 $O/bench_tagged_n1.json
echo "== bench decode chain"; for c in 32; do timeout 600 python bench.py --chain $c --no-cpu --no-prefill --steps 100 2>/dev/null | tail -1 > $O/bench_chain$c.json; cut -c1-200 $O/bench_chain$c.json; done
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemv|gemm_tc" -s 1024 -c 128 --csv --log-file $O/launches_bench.csv python bench.py --no-cpu --no-prefill --steps 2 --warmup 3 > $O/bench_under_ncu.log 2>&1
echo "== ncu full: one decoder block inside the bench"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemv_imma" -s 644 -c 4 -o $O/ncu_decode_block -f python bench.py --no-cpu --no-prefill --steps 1 --warmup 3 > $O/ncu_decode_block.log 2>&1
ncu -i $O/ncu_decode_block.ncu-rep --page raw --csv > $O/ncu_decode_block_raw.csv 2>/dev/null
echo "== ncu full: tcgen05 GEMM M=512"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o $O/ncu_gemm512 -f python tools/microbench.py --m 512 --iters 2 --layouts GPTQ --shapes 4096x4096 > $O/ncu_gemm512.log 2>&1
ncu -i $O/ncu_gemm512.ncu-rep --page raw --csv > $O/ncu_gemm512_raw.csv 2>/dev/null
echo "== microbench"
timeout 600 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM,GPTQ,MARLIN > $O/mb_decode.jsonl 2>&1; cut -c1-200 $O/mb_decode.jsonl
timeout 600 python tools/microbench.py --m 64,512,2048,8192 --layouts GPTQ --iters 50 > $O/mb_gemm.jsonl 2>&1; cut -c1-200 $O/mb_gemm.jsonl
timeout 600 python tools/microbench.py --layouts GPTQ --m 16,64,128,256 --force gemm --graph > $O/mb_gemm_small_m.jsonl 2>&1; cut -c1-200 $O/mb_gemm_small_m.jsonl
echo "== parameter-block size probe"; timeout 120 tools/ubench/param_size_probe | tee $O/param_size_probe.txt | grep "pdl 1"
echo "== compute-sanitizer memcheck (decode chain, tcgen05 GEMM, fused PEER hand-off)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest -q -x "tests/test_chain.py::test_chain_llama_like_block" "tests/test_gpu_parity.py::test_tcgen05_gemm_forced_small_m_and_identity" "tests/test_round2.py::test_gemm_cluster_split_k_matches_unsplit" "tests/test_round2.py::test_sibling_group_prefill_matches_single_calls" tests/test_sharding.py -m gpu > $O/sanitizer_memcheck.txt 2>&1; tail -4 $O/sanitizer_memcheck.txt
echo "== reference Marlin kernel on sm_100"; timeout 600 python tools/marlin_ref_probe.py 2>&1 | tee $O/marlin_ref_probe.txt
echo "== reference CUDA kernels vs the engine"; timeout 600 python tools/ref_bench.py > $O/ref_vs_ours.jsonl 2> $O/ref_vs_ours.err; cut -c1-250 $O/ref_vs_ours.jsonl
rm -f $O/*.ncu-rep.tmp; ls $O
