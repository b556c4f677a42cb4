#!/usr/bin/env python
"""Per-kernel SpeedOfLight summary of an `ncu --section SpeedOfLight --csv` log.
usage: python profiles/summarize_sol.py gpurun_out/sol.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 12]
hdr = rows[0]
iN, iM, iV, iID = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
k = collections.OrderedDict()
for r in rows[1:]:
    d = k.setdefault(r[iID], {'name': re.sub(r'\(.*', '', r[iN]).replace('void ', '')[:56]})
    d[r[iM]] = r[iV]
M = ('Duration', 'Compute (SM) Throughput', 'Memory Throughput', 'DRAM Throughput', 'L2 Cache Throughput', 'L1/TEX Cache Throughput')
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for d in k.values():
    for m in M:
        if m in d:
            try:
                agg[d['name']][m].append(float(d[m].replace(',', '')))
            except ValueError:
                pass
tot = sum(sum(v['Duration']) for v in agg.values())
print(f"{len(k)} launches, {tot / 1e3:.1f} us total (ncu SpeedOfLight section, cold cache, serialised; throughputs = % of peak)")
for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]['Duration'])):
    a = lambda m: sum(v[m]) / len(v[m]) if v[m] else float('nan')
    print(f"{sum(v['Duration']) / 1e3:9.1f} us n={len(v['Duration']):3d} avg={a('Duration') / 1e3:7.1f} us  SM%={a(M[1]):5.1f} Mem%={a(M[2]):5.1f} "
          f"DRAM%={a(M[3]):5.1f} L2%={a(M[4]):5.1f} L1%={a(M[5]):5.1f}  {n}")
