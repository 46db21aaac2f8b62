#!/bin/bash
O=gpurun_out/r2_08; mkdir -p $O
echo "== pytest new"; timeout 1200 python -m pytest tests/test_round2.py tests/test_loader.py tests/test_vs_reference_cuda.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15 | tee $O/pytest_gpu.txt
