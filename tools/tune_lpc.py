#!/usr/bin/env python
"""Times vbx_lpc on the C2 / C3 shapes for several VBX_LPC_PLAN="k:threads" choices (GPU box only)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))
import voxbox_b200 as vb  # noqa: E402

ctx = vb.Context(0)
rng = np.random.default_rng(0)
for name, (N, hop, U, ns) in {"c2": (400, 160, 360, 160000), "c3": (1102, 441, 120, 441000)}.items():
    audio = (0.1 * rng.standard_normal((U, ns))).astype(np.float32)
    J = ctx.n_frames_of(ns, N, hop)
    F = U * J
    d = ctx.to_device(audio)
    r = ctx.empty((F, 13), np.float32)
    a = ctx.empty((F, 13), np.float32)
    fr = ctx.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
    plans = [None] + [f"{k}:{t}" for t in (64, 128) for k in (1, 2, 4, 8)]
    for plan in plans:
        if plan is None:
            os.environ.pop("VBX_LPC_PLAN", None)
        else:
            os.environ["VBX_LPC_PLAN"] = plan
        try:
            for _ in range(3):
                ctx._check(ctx.lib.vbx_lpc(ctx.h, C.byref(fr), 12, r.ptr, a.ptr, None, vb.F32), "lpc")
            ctx.sync()
            ctx.timer_start()
            K = 20
            for _ in range(K):
                ctx._check(ctx.lib.vbx_lpc(ctx.h, C.byref(fr), 12, r.ptr, a.ptr, None, vb.F32), "lpc")
            ms = ctx.timer_stop_ms() / K
            print(f"{name} plan={plan}: {ms*1e3:8.1f} us  {F/ms/1e6:8.2f} Mframes/s  fp64 {F*(26*N+N+350)/ms/1e9:6.2f} TF", flush=True)
        except vb.VoxBoxError as e:
            print(name, plan, "failed:", str(e)[:100])
    d.free(); r.free(); a.free()
print(ctx.measure_peaks())
