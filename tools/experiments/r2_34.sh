#!/bin/bash
# step word read at kernel start (off the x / store critical path): protocol tests + N = 1 tagged bench
timeout 900 python -m pytest tests/test_sharding.py -m gpu -q 2>&1 | tail -3
show() { python -c "
import sys, json
d = json.loads(sys.stdin.read()); c = d['config']; print(d['value'], d['roofline']['frac'], c.get('matches_kernel_boundary_path'), c.get('peer_wait_timeouts'))"; }
for i in 1 2; do
echo "== tagged"; timeout 300 python bench.py --handoff tagged --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | show
echo "== plain"; timeout 300 python bench.py --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | show
done
