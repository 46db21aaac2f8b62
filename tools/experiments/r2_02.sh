#!/bin/bash
# round 2, call 2: decode chain kernel -- parity tests, then the bench with one launch per token / per block / per group
O=gpurun_out/r2_02; mkdir -p $O
echo "== chain tests"; timeout 600 python -m pytest tests/test_chain.py -m gpu -q -x 2>&1 | tail -25 | tee $O/pytest_chain.txt
for c in 32 1 0; do
  echo "== bench --chain $c"; timeout 600 python bench.py --chain $c --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | cut -c1-1500 | tee $O/bench_chain$c.txt
done
