#!/bin/bash
# Prepared at the end of round 1 (GPU minutes exhausted): the probes DESIGN.md section 9 asks for before touching the GEMM.
O=gpurun_out/r2_first; mkdir -p $O
cd tools/ubench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../qllm_b200/csrc -I../../include umma_rate_probe.cu -o umma_rate_probe && cd ../..
echo "== tcgen05.mma rate probe: issue-slot starvation"; UMMA_PROBE_ALU=1 timeout 60 tools/ubench/umma_rate_probe 2>&1 | tee $O/umma_rate_probe_alu.txt
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/pytest.txt
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -1 | cut -c1-600 | tee $O/bench.txt
