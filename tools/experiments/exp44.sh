#!/bin/bash
O=gpurun_out/e44; mkdir -p $O
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^(FAILED|E  )|passed|failed" | head -30 | tee $O/pytest.txt
