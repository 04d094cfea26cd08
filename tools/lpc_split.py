#!/usr/bin/env python
"""A/B: fused window->autocorr->Levinson kernel vs autocorrelation kernel + stand-alone Levinson kernel (C2 shape)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))
import voxbox_b200 as vb
ctx = vb.Context(0)
rng = np.random.default_rng(0)
N, hop, U, ns, p = 400, 160, 360, 160000, 12
audio = (0.1 * rng.standard_normal((U, ns))).astype(np.float32)
J = ctx.n_frames_of(ns, N, hop); F = U * J
d = ctx.to_device(audio)
r32, a32 = ctx.empty((F, 13), np.float32), ctx.empty((F, 13), np.float32)
r64 = ctx.empty((F, 13), np.float64)
fr = ctx.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
L = ctx.lib
def t(fn, K=50):
    for _ in range(3): fn()
    ctx.sync(); ctx.timer_start()
    for _ in range(K): fn()
    return ctx.timer_stop_ms() / K * 1e3
fused = lambda: ctx._check(L.vbx_lpc(ctx.h, C.byref(fr), p, r32.ptr, a32.ptr, None, vb.F32), "lpc")
ac_only = lambda: ctx._check(L.vbx_autocorrelate(ctx.h, C.byref(fr), 13, r64.ptr, vb.F64), "ac")
lev = lambda: ctx._check(L.vbx_lpc_levinson(ctx.h, r64.ptr, vb.F64, F, 13, p, a32.ptr, None, vb.F32), "lev")
def split():
    ac_only(); lev()
print(f"fused {t(fused):.1f} us | autocorr only {t(ac_only):.1f} us | levinson only {t(lev):.1f} us | split {t(split):.1f} us")
