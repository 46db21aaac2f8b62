#!/bin/bash
O=gpurun_out/e26; mkdir -p $O
echo "== f8 probe"; timeout 120 tools/ubench/f8_probe 2>&1 | tee $O/f8_probe.txt
echo "== dbg"; timeout 300 python tools/dbg_stream.py GEMM 2048 640 2>&1 | tail -14 | tee $O/dbg.txt
