#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r'\(.*', '', r[4])
    name = re.sub(r'^void ', '', name)[:80]
    agg[name][0] += 1
    agg[name][1] += float(r[-1]) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot:.1f} us total (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}% n={v[0]:4d} avg={v[1] / v[0]:8.1f} us  {k}")
