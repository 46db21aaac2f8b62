#!/bin/bash
# sibling release (b200q_linear_group at M > 64) + refitted tile model
O=gpurun_out/r2_21; mkdir -p $O
timeout 900 python -m pytest tests/test_round2.py tests/test_full_size_parity.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
echo "== prefill7b grouped"; timeout 300 python bench.py --config prefill7b --no-cpu --steps 5 2>&1 | tail -1 | tee $O/bench_prefill7b.json | cut -c1-200
echo "== prefill7b per-layer"; B200Q_BENCH_NO_GROUP=1 timeout 300 python bench.py --config prefill7b --no-cpu --steps 5 2>&1 | tail -1 | tee $O/bench_prefill7b_nogroup.json | cut -c1-200
echo "== prefill7b m2048 grouped"; timeout 300 python bench.py --config prefill7b --m 2048 --no-cpu --steps 3 2>&1 | tail -1 | tee $O/bench_prefill7b_m2048.json | cut -c1-200
echo "== prefill7b m2048 per-layer"; B200Q_BENCH_NO_GROUP=1 timeout 300 python bench.py --config prefill7b --m 2048 --no-cpu --steps 3 2>&1 | tail -1 | tee $O/bench_prefill7b_m2048_nogroup.json | cut -c1-200
echo "== default"; timeout 600 python bench.py --no-cpu --steps 50 2>&1 | tail -1 | tee $O/bench_n1.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d.get('prefill'))"
