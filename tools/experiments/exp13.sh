#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
echo "== quick decode tests"; timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode" 2>&1 | tail -3
echo "== all parity tests"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/microbench.py --m 1 --graph --iters 400 --layouts GEMM,GPTQ,MARLIN 2>&1 | tee $O/mb13_dec.log | cut -c1-150
echo "== bench"; timeout 600 python bench.py --no-prefill 2>&1 | tee $O/bench13.json | cut -c1-400
