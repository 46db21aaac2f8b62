#!/bin/bash
O=gpurun_out/e32; mkdir -p $O
echo "== bench timeline"; timeout 300 python tools/bench_timeline.py --blocks 3 2>&1 | tail -16 | tee $O/bench_timeline.txt
echo "== bench"; timeout 600 python bench.py --no-cpu --no-prefill --steps 30 2>&1 | tail -1 | cut -c1-400
