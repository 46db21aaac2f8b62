#!/bin/bash
O=gpurun_out/r2_03; mkdir -p $O
echo "== chain tests"; timeout 600 python -m pytest tests/test_chain.py -m gpu -q 2>&1 | tail -25 | tee $O/pytest_chain.txt
echo "== timeline"; timeout 300 python tools/chain_timeline.py 4 2>&1 | tail -30 | tee $O/chain_timeline.txt
echo "== bench --chain 32"; timeout 600 python bench.py --chain 32 --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | cut -c1-300 | tee $O/bench_chain32.txt
