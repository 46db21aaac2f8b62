"""Summarise `ncu --page source --csv` output: opcode mix + top stall sites for the first kernel."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if 'Source' in r and 'Address' in r)
si, ei, st = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
data, seen, started = [], set(), False
for r in rows:
    if r == hdr:
        if started:
            break
        started = True
        continue
    if not started or len(r) <= ei or not r[ei].strip().isdigit():
        continue
    data.append(r)
tot = sum(int(r[ei]) for r in data)
print("static instrs", len(data), "executed warp-instrs", tot)
ops, stalls = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si])
    op = m.group(2).split('.')[0] if m else '?'
    ops[op] += int(r[ei])
    stalls[op] += int(r[st] or 0)
for op, c in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print(f"{op:10s} {c:9d} {100*c/tot:5.1f}%  stall_samples={stalls[op]}")
print("--- top stall sites")
for r in sorted(data, key=lambda r: -int(r[st] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 20]:
    print(r[st].rjust(6), r[ei].rjust(8), r[si][:100])
