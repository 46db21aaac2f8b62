#!/bin/bash
O=gpurun_out/r2_07; mkdir -p $O
echo "== pytest -m gpu (all)"; timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $O/pytest_gpu.txt
echo "== bench default"; timeout 600 python bench.py --steps 50 2>&1 | tail -1 | cut -c1-2500 | tee $O/bench.txt
echo "== bench prefill7b"; timeout 600 python bench.py --config prefill7b --steps 10 2>&1 | tail -1 | cut -c1-2000 | tee $O/bench_prefill7b.txt
echo "== bench act13b"; timeout 600 python bench.py --config act13b --steps 20 --no-cpu 2>&1 | tail -1 | cut -c1-2000 | tee $O/bench_act13b.txt
echo "== bench mixtral"; timeout 900 python bench.py --config mixtral --steps 20 2>&1 | tail -1 | cut -c1-3500 | tee $O/bench_mixtral.txt
