#!/usr/bin/env python
"""Pretty-print bench.py JSON lines: python tools/show_bench.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    d = json.load(open(path))
    cb = d.get("cpu_baseline", {})
    r = d.get("roofline") or {}
    print(f"{path.split('/')[-1]}: value={d['value']:.4g} {d['unit']}  ms/step={d['ms_per_step']:.3f}  e2e={d['e2e']['value']:.4g}  "
          f"cpu={cb.get('value', 0):.4g} ({cb.get('cores')} thr) / {cb.get('single_thread_value', 0):.4g} (1 thr)  launches={d.get('gpu_launches')}")
    if r:
        print(f"    roofline: {r['kernel']} bound={r['bound']} achieved={r['achieved']:.3g} {r['unit']} of {r['peak']:.3g} = {r['frac']:.3f}  "
              f"share_of_step={r.get('share_of_step', 0):.3f} traffic={r.get('traffic')}")
    for k, v in (d.get("kernels") or {}).items():
        print("      ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})
    print("    clocks:", d.get("clocks"))
