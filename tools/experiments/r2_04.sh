#!/bin/bash
O=gpurun_out/r2_04; mkdir -p $O
echo "== chain tests"; timeout 600 python -m pytest tests/test_chain.py -m gpu -q 2>&1 | tail -5 | tee $O/pytest_chain.txt
echo "== timeline (window 8)"; timeout 300 python tools/chain_timeline.py 3 2>&1 | tail -14 | tee $O/chain_timeline_w8.txt
echo "== timeline (window 4)"; B200Q_OPTS=chain_window=4 timeout 300 python tools/chain_timeline.py 3 2>&1 | tail -14 | tee $O/chain_timeline_w4.txt
for w in 3 4 6 8 12 16; do
  echo "== bench --chain 32, window $w"; B200Q_OPTS=chain_window=$w timeout 600 python bench.py --chain 32 --no-cpu --no-prefill --steps 50 2>&1 | tail -1 | cut -c1-200 | tee $O/bench_w$w.txt
done
