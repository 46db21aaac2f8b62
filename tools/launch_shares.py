"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel shares.
usage: python tools/launch_shares.py gpurun_out/r1/launches_bench.csv > profiles/r1_launches_bench.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    key = (r[4], r[7], r[8])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14]) / 1e3
tot = sum(a[1] for a in agg.values())
print(f"# {len(rows)} launches, {tot:.1f} us total (cold-cache, serialised under ncu: compare SHARES, not absolutes)")
print(f"{'kernel':70s} {'block':>12s} {'grid':>14s} {'n':>5s} {'us/launch':>10s} {'share':>7s}")
for (k, b, g), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {b:>12s} {g:>14s} {n:5d} {t / n:10.2f} {100 * t / tot:6.1f}%")
