#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: where the stall samples and instructions go.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv [kernel-index]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
# split per kernel
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 5:
        cur["rows"].append(r)
k = kern[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
h = k["hdr"]
ci = {n: i for i, n in enumerate(h)}
S, I = ci["# Samples"], ci["Instructions Executed"]
tot_s = sum(int(r[S]) for r in k["rows"])
tot_i = sum(int(r[I]) for r in k["rows"])
print(k["name"], "samples", tot_s, "warp-instr", tot_i)
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(int(r[ci[n]] or 0) for r in k["rows"]) for n in stalls}
print("stall mix:", {n: round(v / tot_s, 3) for n, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
# regions: cumulative samples in windows of 40 instructions
W = 40
print("region (instr idx) : samples%  instr%  top opcode mix")
for a in range(0, len(k["rows"]), W):
    blk = k["rows"][a:a + W]
    s = sum(int(r[S]) for r in blk)
    i = sum(int(r[I]) for r in blk)
    if s / tot_s > 0.02:
        ops = {}
        for r in blk:
            op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
            ops[op] = ops.get(op, 0) + 1
        top = sorted(ops.items(), key=lambda x: -x[1])[:4]
        print(f"{a:5d}-{a+W:5d}: {100*s/tot_s:5.1f}% {100*i/tot_i:5.1f}%  {top}")
print("top instructions by samples:")
for r in sorted(k["rows"], key=lambda r: -int(r[S]))[:25]:
    st = {n: int(r[ci[n]] or 0) for n in stalls}
    top = max(st.items(), key=lambda x: x[1])
    print(f"{100*int(r[S])/tot_s:5.2f}%  exec={r[I]:>9}  {r[1].strip()[:70]:70s} {top}")
