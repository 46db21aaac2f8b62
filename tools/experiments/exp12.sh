#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
echo "== decode tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python tools/sweep_cluster.py 2>&1 | tee $O/sweep12.log
