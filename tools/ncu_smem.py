#!/usr/bin/env python
"""Shared-memory wavefronts (actual / ideal) and instruction share per region of W SASS instructions.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_smem.py src.csv n_units [W]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
F = float(sys.argv[2]); W = int(sys.argv[3]) if len(sys.argv) > 3 else 80
hdr = None; data = []
for r in rows:
    if r and r[0] == 'Address': hdr = r
    elif hdr and len(r) > 5: data.append(r)
ci = {n: i for i, n in enumerate(hdr)}
I = ci['Instructions Executed']; Wv = ci['L1 Wavefronts Shared']; Wi = ci['L1 Wavefronts Shared Ideal']; S = ci['# Samples']
def num(x):
    try: return int(x)
    except ValueError: return 0
tot = sum(num(r[Wv]) for r in data); tot_i = sum(num(r[I]) for r in data); tot_s = sum(num(r[S]) for r in data)
print(f"per unit: wavefronts {tot/F:.1f} (ideal {sum(num(r[Wi]) for r in data)/F:.1f}), warp instructions {tot_i/F:.1f}")
for a in range(0, len(data), W):
    blk = data[a:a + W]
    w = sum(num(r[Wv]) for r in blk); wi = sum(num(r[Wi]) for r in blk); i = sum(num(r[I]) for r in blk); s = sum(num(r[S]) for r in blk)
    if i / tot_i > 0.01 or w / max(tot, 1) > 0.01:
        kinds = {}
        for r in blk:
            if num(r[Wv]):
                op = r[1].strip().split(); op = op[1] if op[0].startswith('@') else op[0]
                kinds[op] = kinds.get(op, 0) + num(r[Wv])
        print(f"{a:5d}: wavefronts {w/F:6.1f} (ideal {wi/F:6.1f})  instr {i/F:6.1f} ({100*i/tot_i:4.1f}%)  samples {100*s/tot_s:4.1f}%  {[(k, round(v/F,1)) for k, v in sorted(kinds.items(), key=lambda x: -x[1])[:3]]}")
