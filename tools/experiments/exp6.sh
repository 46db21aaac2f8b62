#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
echo "== launch probe"; timeout 120 tools/ubench/launch_probe 2>&1 | tee $O/launch_probe.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest6.log; echo "== pytest all: $(tail -1 $O/pytest6.log)"; grep -E "FAILED|Error" $O/pytest6.log | head
echo "== ref bench"; timeout 600 python tools/ref_bench.py 2>&1 | tee $O/ref_bench.log | cut -c1-400
