#!/usr/bin/env python
"""Per-kernel DRAM traffic of one step from an
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv
log (long format: one row per launch and metric).  usage: python profiles/summarize_traffic.py gpurun_out/traffic.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = next(r for r in rows if 'Metric Name' in r)
iN, iM, iU, iV, iID = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value', 'ID'))
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3}
per = collections.OrderedDict()
for r in rows:
    if not r[iID].isdigit():
        continue
    d = per.setdefault(r[iID], {'name': re.sub(r'^void ', '', re.sub(r'\(.*', '', r[iN]))[:64]})
    d[r[iM]] = float(r[iV].replace(',', '')) * SCALE.get(r[iU], 1.0)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in per.values():
    a = agg[d['name']]
    a[0] += 1
    a[1] += d.get('gpu__time_duration.sum', 0.0)
    a[2] += d.get('dram__bytes_read.sum', 0.0)
    a[3] += d.get('dram__bytes_write.sum', 0.0)
T = sum(a[1] for a in agg.values())
R = sum(a[2] for a in agg.values())
W = sum(a[3] for a in agg.values())
print(f"{len(per)} launches, {T:.1f} us (serialised), DRAM read {R / 1e6:.1f} MB + write {W / 1e6:.1f} MB = {(R + W) / 1e6:.1f} MB per step")
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1][2] + kv[1][3])):
    gbs = (a[2] + a[3]) / (a[1] * 1e-6) / 1e9 if a[1] else 0.0
    print(f"{(a[2] + a[3]) / 1e6:9.1f} MB (r {a[2] / 1e6:8.1f} w {a[3] / 1e6:8.1f}) n={a[0]:3d} {a[1]:8.1f} us {gbs:7.0f} GB/s  {k}")
