#!/bin/bash
# act-order consumers of tagged activations: protocol test on one GPU, act13b with --handoff tagged at N = 1
O=gpurun_out/r2_32; mkdir -p $O
timeout 900 python -m pytest tests/test_sharding.py -m gpu -q 2>&1 | tail -4
show() { python -c "
import sys, json
d = json.loads(sys.stdin.read()); c = d['config']; print(d['value'], d['roofline']['frac'], c.get('matches_kernel_boundary_path'), c.get('peer_wait_timeouts'), c.get('launches_per_step'))"; }
echo "== act13b tagged"; timeout 400 python bench.py --config act13b --handoff tagged --no-cpu --steps 50 2>&1 | tail -1 | tee $O/bench_act13b_tagged.json | show
echo "== act13b kernel boundary"; timeout 400 python bench.py --config act13b --no-cpu --steps 50 2>&1 | tail -1 | tee $O/bench_act13b.json | show
echo "== decode7b tagged"; timeout 400 python bench.py --handoff tagged --no-cpu --no-prefill --steps 50 2>&1 | tail -1 | show
