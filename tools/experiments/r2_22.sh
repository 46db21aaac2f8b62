#!/bin/bash
# sibling-aware tile choice (SM time instead of waves inside a b200q_linear_group call)
O=gpurun_out/r2_22; mkdir -p $O
timeout 900 python -m pytest tests/test_round2.py -m gpu -q -x 2>&1 | tail -3
echo "== prefill7b grouped"; timeout 300 python bench.py --config prefill7b --no-cpu --steps 5 2>&1 | tail -1 | tee $O/bench_prefill7b.json | cut -c1-200
echo "== prefill7b m2048 grouped"; timeout 300 python bench.py --config prefill7b --m 2048 --no-cpu --steps 3 2>&1 | tail -1 | tee $O/bench_prefill7b_m2048.json | cut -c1-200
echo "== prefill7b m256 grouped"; timeout 300 python bench.py --config prefill7b --m 256 --no-cpu --steps 5 2>&1 | tail -1 | tee $O/bench_prefill7b_m256.json | cut -c1-200
echo "== prefill7b m1024 grouped"; timeout 300 python bench.py --config prefill7b --m 1024 --no-cpu --steps 5 2>&1 | tail -1 | tee $O/bench_prefill7b_m1024.json | cut -c1-200
