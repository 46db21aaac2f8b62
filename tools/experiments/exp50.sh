#!/bin/bash
O=gpurun_out/e50; mkdir -p $O
timeout 200 compute-sanitizer --tool memcheck --print-limit 6 python tools/microbench.py --layouts GPTQ --bits 4 --group 128 --m 16 --iters 1 --shapes 4096x4096 2>&1 | grep -v "^$" | head -60 | cut -c1-240 | tee $O/memcheck.txt
