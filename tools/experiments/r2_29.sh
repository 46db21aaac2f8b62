#!/bin/bash
# N = 1: tagged activations between the QuantLinears (no kernel-boundary wait ahead of x) vs the plain chain
O=gpurun_out/r2_29; mkdir -p $O
show() { python -c "
import sys, json
d = json.loads(sys.stdin.read()); c = d['config']; print(d['value'], d['roofline']['frac'], d['e2e']['value'], c.get('matches_kernel_boundary_path'), c.get('peer_wait_timeouts'), c.get('launches_per_step'), c.get('outputs_finite'))"; }
for i in 1 2; do
  echo "== tagged"; timeout 300 python bench.py --handoff tagged --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | tee $O/bench_tagged_$i.json | show
  echo "== kernel boundary"; timeout 300 python bench.py --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | tee $O/bench_plain_$i.json | show
done
