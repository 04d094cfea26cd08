#!/usr/bin/env python
"""Instruction mix of the hottest loop (most DFMA/FFMA in a backward-branch body) of one kernel.
usage: cuobjdump -sass lib.so > all.sass; python tools/sass_loop.py all.sass <substring of mangled name> [DFMA|FFMA]"""
import re
import sys
from collections import Counter

txt = open(sys.argv[1]).read()
key = sys.argv[2]
op = sys.argv[3] if len(sys.argv) > 3 else "DFMA"
funcs = re.split(r"\n\s*Function : ", txt)
f = [x for x in funcs if x.split("\n", 1)[0].find(key) >= 0][0]
ins = []
for l in f.splitlines():
    m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(f.split("\n", 1)[0][:120], "instructions:", len(ins), f"total {op}:", sum(op in t for _, t in ins))
loops = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+.*0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        tgt = int(m.group(1), 16)
        body = [x for x in ins if tgt <= x[0] <= a]
        loops.append((sum(op in b[1] for b in body), tgt, a, len(body)))
loops.sort(reverse=True)
for nd, tgt, a, n in loops[:3]:
    body = [x for x in ins if tgt <= x[0] <= a]
    c = Counter((b[1].split()[1] if b[1].startswith("@") else b[1].split()[0]) for b in body)
    print(f"loop {tgt:#x}-{a:#x}: {n} instr, {nd} {op}:", c.most_common(14))
