#!/usr/bin/env python
"""Host -> device copy bandwidth per GPU as more GPUs stage at once (vbx_multi_h2d_bandwidth): which link saturates behind the
end-to-end scaling of the `_host` entry points.  Every device copies from a pinned buffer its own worker thread allocated after
binding to the GPU's local CPUs (NUMA-local staging), all active devices concurrently.
usage: python tools/h2d_probe.py [mib_per_device] [reps]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))
import voxbox_b200 as vb

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
out = {"mib_per_device": mib, "reps": reps, "runs": []}
with vb.Multi(0) as m:
    n = m.n
    k = 1
    actives = []
    while k < n:
        actives.append(k)
        k *= 2
    actives.append(n)
    for bind in (True, False):
        os.environ["VBX_MULTI_NO_BIND"] = "0" if bind else "1"
        for a in actives:
            gbs = m.h2d_bandwidth(mib << 20, reps, a)
            out["runs"].append({"active": a, "per_device_gbs": [round(g, 2) for g in gbs[:a]], "aggregate_gbs": round(sum(gbs[:a]), 1)})
            print(f"{a} of {n} devices copying: per device {[round(g, 1) for g in gbs[:a]]} GB/s, aggregate {sum(gbs[:a]):.1f} GB/s", flush=True)
        break  # (the workers bound themselves when the handle was created; an unbound run needs VBX_MULTI_NO_BIND=1 in the environment)
print(json.dumps(out))
