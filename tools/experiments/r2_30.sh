#!/bin/bash
# tagged hand-off at N = 1: polling back-off 96 ns (default) / 32 ns / none
O=gpurun_out/r2_30; mkdir -p $O
show() { python -c "
import sys, json
d = json.loads(sys.stdin.read()); c = d['config']; print(d['value'], d['roofline']['frac'], c.get('matches_kernel_boundary_path'), c.get('peer_wait_timeouts'))"; }
for i in 1 2; do
  for f in 0 64 32; do
    echo "== tagged sync_flags=$f"; B200Q_OPTS=sync_flags=$f timeout 300 python bench.py --handoff tagged --no-cpu --no-prefill --steps 100 2>&1 | tail -1 | show
  done
done
