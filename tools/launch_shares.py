#!/usr/bin/env python
"""Per-kernel launch counts, total time and share of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_shares.py profiles/r2_launches_bench_c3_final.csv [...]"""
import collections
import csv
import re
import sys

for path in sys.argv[1:]:
    rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
    hdr = rows[0]
    ci = {n: i for i, n in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ci['Metric Name']] != 'gpu__time_duration.sum':
            continue
        m = re.search(r'(\w+_kernel)', r[ci['Kernel Name']])
        name = m.group(1) if m else r[ci['Kernel Name']][:40]
        v = float(r[ci['Metric Value']].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[ci['Metric Unit']], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    work = {k: v for k, v in agg.items() if k not in ('vbx_fma_peak_kernel', 'synth_speech_kernel')}   # peak probe / corpus generator: not part of a step
    tot = sum(v for _, v in work.values()) or 1.0
    print(path.split('/')[-1])
    for k, (n, v) in agg.items():
        tag = f"{v / tot:6.3f}" if k in work else "  (setup)"
        print(f"    {k:28s} launches {n:4d}  total {v:9.3f} ms  share {tag}")
