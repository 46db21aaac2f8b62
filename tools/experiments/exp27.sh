#!/bin/bash
O=gpurun_out/e27; mkdir -p $O
echo "== imma probe"; timeout 120 tools/ubench/imma_probe 2>&1 | tee $O/imma_probe.txt
echo "== pytest decode"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode or stream or sibling or group or full_size or workspace or strided" 2>&1 | tail -4 | tee $O/pytest.txt
