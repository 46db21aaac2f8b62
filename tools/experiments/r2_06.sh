#!/bin/bash
O=gpurun_out/r2_06; mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_chain -s 3 -c 1 -o $O/ncu_chain -f python tools/chain_timeline.py 2 > $O/ncu.log 2>&1
tail -3 $O/ncu.log
ncu -i $O/ncu_chain.ncu-rep --page raw --csv > $O/ncu_chain_raw.csv 2>/dev/null
ncu -i $O/ncu_chain.ncu-rep --page source --csv --print-source sass > $O/ncu_chain_source.csv 2>/dev/null
ls -la $O
