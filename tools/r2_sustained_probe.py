#!/usr/bin/env python
"""GPU box: per-step time of the C3 device-resident loop as a function of how many steps are queued back to back, with and
without the per-launch profiling events."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))
import bench, voxbox_b200 as vb
ctx = vb.Context(0)
cfg = dict(bench.CONFIGS["c3"])
wl = bench.Workload(ctx, vb, cfg, 0, want_host=False)
for _ in range(5): wl.step()
ctx.sync()
for prof in (False, True, False, True):
    for n in (10, 40):
        if prof: ctx.profile_begin()
        ctx.timer_start()
        for _ in range(n): wl.step()
        ms = ctx.timer_stop_ms()
        names = ctx.profile_end() if prof else {}
        print(f"profiling {prof}: {n:4d} steps: {ms / n:.3f} ms/step", {k: round(v[0] / n, 3) for k, v in names.items()}, flush=True)
