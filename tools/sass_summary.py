#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md):
    UTCHMMA (tcgen05.mma kind::f16), UTCBAR (tcgen05.commit), LDTM (tcgen05.ld), UTMALDG (cp.async.bulk.tensor),
    UBLKCP (cp.async.bulk), SYNCS (mbarrier), plus cluster / peer-memory markers.
usage: python tools/sass_summary.py [lib.so] > profiles/sass_summary.txt   (cuobjdump only; no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "xggm_b200", "libxggm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "UCGABAR", "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS", "RED.E.ADD.F32", "HMMA", "FFMA"]
counts = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[name]["_total"] += 1
        for mn in MNEMONICS:
            if op.startswith(mn):
                counts[name][mn] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except OSError:
        return n


arch = set(re.findall(r"arch = (sm_\w+)", out))
print(f"# {os.path.basename(lib)}: {len(counts)} kernels, architectures {sorted(arch)}")
print(f"# {'instr':>7} " + " ".join(f"{m[:9]:>9}" for m in MNEMONICS) + "  kernel")
tot = collections.Counter()
for n, c in sorted(counts.items(), key=lambda kv: -(kv[1]["UTCHMMA"] * 1000 + kv[1]["UTMALDG"] * 10 + kv[1]["UBLKCP"])):
    if not any(c[m] for m in MNEMONICS[:8]):
        continue
    d = re.sub(r"\(.*", "", demangle(n))[:90]
    print(f"  {c['_total']:>7} " + " ".join(f"{c[m]:>9}" for m in MNEMONICS) + f"  {d}")
    tot.update(c)
print(f"# totals over the kernels listed: " + ", ".join(f"{m}={tot[m]}" for m in MNEMONICS[:8]))
