#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
echo "== mma probe"; timeout 120 tools/ubench/mma_probe 2>&1 | tee $O/mma_probe.log
echo "== timelines (smem variant default)"; for sh in 4096x4096 4096x11008; do timeout 120 python tools/timeline.py --layout GEMM --shape $sh --launches 6 2>&1 | tee -a $O/timeline2.log; done
timeout 120 python tools/timeline.py --layout GPTQ --shape 4096x4096 --launches 6 2>&1 | tee -a $O/timeline2.log
echo "== decode smem default"; timeout 300 python tools/microbench.py --m 1 --graph --iters 400 2>&1 | tee $O/mb5_dec.log
echo "== sharding tests"; timeout 300 python -m pytest tests/test_sharding.py -m gpu -q 2>&1 | tail -3
