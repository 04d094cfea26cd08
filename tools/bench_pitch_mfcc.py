#!/usr/bin/env python
"""Timings of the pitch path on the C4 shape (16 kHz, N=640, hop=160, 75-600 Hz) and of MFCC on the C5 shape
(16 kHz, N=400, hop=160, 40 bands, 13 kept) — GPU box only.  Used for tuning, not the contract bench."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))
import voxbox_b200 as vb  # noqa: E402
from voxbox_b200 import synth  # noqa: E402

ctx = vb.Context(0)
U = int(os.environ.get("UTTS", "120"))
reps = int(os.environ.get("REPS", "3"))
which = os.environ.get("WHICH", "pitch,mfcc,mfcc32").split(",")


def timeit(fn, reps=reps, warm=1):
    for _ in range(warm):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop_ms() / reps


fs = 16000
base = synth.corpus(12, fs, 10.0)
audio = np.tile(base, (max(U // 12, 1), 1))[:U]
ns = audio.shape[1]
d = ctx.to_device(audio)
L = ctx.lib
if "pitch" in which:
    N, hop = 640, 160
    J = ctx.n_frames_of(ns, N, hop)
    F = U * J
    fr = ctx.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
    cand = ctx.empty((F, 16, 2), np.float32)
    nc = ctx.empty((F,), np.int32)
    st = ctx.empty((F,), np.uint8)
    ms = timeit(lambda: ctx._check(L.vbx_pitch(ctx.h, C.byref(fr), float(fs), 0.45, 75.0, 600.0, 16, cand.ptr, nc.ptr, st.ptr, vb.F32), "pitch"))
    n = nc.to_host()
    print(f"pitch C4: F={F}  {ms:9.3f} ms  {F/ms/1e3:9.3f} Mframes/s   mean candidates/frame {n.mean()-1:.2f} max {n.max()-1}", flush=True)
if "mfcc" in which or "mfcc32" in which:
    N, hop = 400, 160
    J = ctx.n_frames_of(ns, N, hop)
    F = U * J
    fr = ctx.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
    out = ctx.empty((F, 13), np.float32)
    for name, dt in (("mfcc", vb.F64), ("mfcc32", vb.F32)):
        if name not in which:
            continue
        ctx.mfcc_set_fft_precision(dt)
        ms = timeit(lambda: ctx._check(L.vbx_mfcc(ctx.h, C.byref(fr), 40, 13, 133.0, 6855.0, float(fs), out.ptr, None, vb.F32), "mfcc"), reps=reps * 5)
        print(f"{name} C5 (fft {'f64' if dt == vb.F64 else 'f32'}): F={F}  {ms:9.3f} ms  {F/ms/1e3:9.3f} Mframes/s", flush=True)
print(ctx.measure_peaks())
