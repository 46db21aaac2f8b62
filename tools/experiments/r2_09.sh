#!/bin/bash
O=gpurun_out/r2_09; mkdir -p $O
echo "== pytest"; timeout 1500 python -m pytest tests/test_round2.py tests/test_gpu_parity.py tests/test_sharding.py -m gpu -q -x 2>&1 | tail -15 | tee $O/pytest_gpu.txt
