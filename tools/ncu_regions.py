#!/usr/bin/env python
"""Per-region (windows of W SASS instructions) share of stall samples and executed instructions of one kernel of an ncu report.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_regions.py src.csv [W]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
W = int(sys.argv[2]) if len(sys.argv) > 2 else 80
hdr = None; data = []
for r in rows:
    if r and r[0] == 'Address': hdr = r
    elif hdr and len(r) > 5: data.append(r)
ci = {n: i for i, n in enumerate(hdr)}
S, I = ci['# Samples'], ci['Instructions Executed']
stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
tot_s = sum(int(r[S]) for r in data); tot_i = sum(int(r[I]) for r in data)
print("total samples", tot_s, "warp-instr", tot_i, "SASS instructions", len(data))
for a in range(0, len(data), W):
    blk = data[a:a + W]
    s = sum(int(r[S]) for r in blk); i = sum(int(r[I]) for r in blk)
    if i == 0: continue
    ex = sorted(int(r[I]) for r in blk)[len(blk) // 2]
    ops = {}
    for r in blk:
        t = r[1].strip().split(); op = t[1] if t[0].startswith('@') else t[0]; op = op.split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:5]
    agg = {n: sum(int(r[ci[n]] or 0) for r in blk) for n in stalls}
    st = sorted(agg.items(), key=lambda x: -x[1])[:2]
    print(f"{a:5d}: samp {100*s/tot_s:5.1f}% instr {100*i/tot_i:5.1f}% median-exec {ex:8d} {top} {[(n[6:], round(v/max(s,1),2)) for n,v in st]}")
