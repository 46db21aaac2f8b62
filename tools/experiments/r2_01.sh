#!/bin/bash
# round 2, call 1: design probes (grid barrier, TMA weight stream), the tcgen05 issue-starvation probe, new full-size parity tests
O=gpurun_out/r2_01; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $O/gpu.txt
echo "== chain probe"; timeout 300 tools/ubench/chain_probe 2>&1 | tee $O/chain_probe.txt
echo "== umma ALU probe"; UMMA_PROBE_ALU=1 timeout 60 tools/ubench/umma_rate_probe 2>&1 | tee $O/umma_rate_probe_alu.txt
echo "== full-size parity"; timeout 1500 python -m pytest tests/test_full_size_parity.py -m gpu -q -x 2>&1 | tail -15 | tee $O/pytest_full_size.txt
