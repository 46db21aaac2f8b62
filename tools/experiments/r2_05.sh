#!/bin/bash
O=gpurun_out/r2_05; mkdir -p $O
echo "== chain tests"; timeout 600 python -m pytest tests/test_chain.py -m gpu -q 2>&1 | tail -8 | tee $O/pytest_chain.txt
echo "== timeline"; timeout 300 python tools/chain_timeline.py 3 2>&1 | tail -14 | tee $O/chain_timeline.txt
for w in 8 16 32; do
  echo "== bench --chain 32, window $w"; B200Q_OPTS=chain_window=$w timeout 600 python bench.py --chain 32 --no-cpu --no-prefill --steps 50 2>&1 | tail -1 | cut -c1-200 | tee $O/bench_w$w.txt
done
echo "== bench --chain 1"; timeout 600 python bench.py --chain 1 --no-cpu --no-prefill --steps 50 2>&1 | tail -1 | cut -c1-200 | tee $O/bench_c1.txt
