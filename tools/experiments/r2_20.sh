#!/bin/bash
# cluster split-K (CTA pair, DSMEM exchange of half-tiles) + modelled tile choice in the tcgen05 GEMM
O=gpurun_out/r2_20; mkdir -p $O
timeout 900 python tools/sweep_gemm_tiles.py 128,200,256,512,1024,2048 2>&1 | tee $O/sweep_gemm_tiles.jsonl
timeout 900 python -m pytest tests/test_full_size_parity.py tests/test_gpu_parity.py tests/test_vs_reference_cuda.py tests/test_round2.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --config prefill7b --no-cpu --steps 5 2>&1 | tail -1 | tee $O/bench_prefill7b.json | cut -c1-200
timeout 300 python bench.py --config prefill7b --m 2048 --no-cpu --steps 3 2>&1 | tail -1 | tee $O/bench_prefill7b_m2048.json | cut -c1-200
