#!/usr/bin/env python
"""Stage timings of the formant path on the C3 shape (44.1 kHz, N=1102, hop=441, order 12) and the C2 shape
(GPU box only).  Prints frames/s per stage; used for tuning, not the contract bench."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))
import voxbox_b200 as vb  # noqa: E402
from voxbox_b200 import synth  # noqa: E402

ctx = vb.Context(0)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop_ms() / reps


for name, fs, N, hop, U in (("c3", 44100, 1102, 441, 120), ("c2", 16000, 400, 160, 360)):
    base = synth.corpus(12, fs, 10.0)
    audio = np.tile(base, (U // 12, 1))
    ns = audio.shape[1]
    J = ctx.n_frames_of(ns, N, hop)
    F = U * J
    p = 12
    d = ctx.to_device(audio)
    male = np.tile(np.array([[f, 1.0] for f in (320., 1440., 2760., 3200.)]), (U, 1, 1))
    fr_sym = ctx.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
    fr_per = ctx.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_PERIODIC, frames_per_segment=J, segment_stride=ns)
    ac = ctx.empty((F, p + 1), np.float64)
    co = ctx.empty((F, p), np.float64)
    st = ctx.empty((F,), np.uint8)
    res = ctx.empty((F, p, 2), np.float32)
    nres = ctx.empty((F,), np.int32)
    est = ctx.to_device(male.astype(np.float32))
    trk = ctx.empty((F, 4, 2), np.float32)
    L = ctx.lib
    t = {}
    t["lpc(autocorr+levinson)"] = timeit(lambda: ctx._check(L.vbx_lpc(ctx.h, C.byref(fr_sym), p, None, ac.ptr, None, vb.F64), "lpc"))
    t["burg"] = timeit(lambda: ctx._check(L.vbx_lpc_burg(ctx.h, C.byref(fr_per), p, co.ptr, st.ptr, vb.F64), "burg"))
    for prec in (0, 1):
        t[f"roots+resonances prec={prec}"] = timeit(lambda: ctx._check(L.vbx_lpc_to_resonances(
            ctx.h, ac.ptr, vb.F64, F, p + 1, p, 1, float(fs), 1, None, res.ptr, p, nres.ptr, None, st.ptr, vb.F32, prec), "roots"))
    t["tracker"] = timeit(lambda: ctx._check(L.vbx_estimate_formants(
        ctx.h, res.ptr, vb.F32, p, 32, U, J, None, est.ptr, 4, trk.ptr, vb.F32), "trk"))
    t["find_formants path A (autocorr)"] = timeit(lambda: ctx._check(L.vbx_find_formants(
        ctx.h, C.byref(fr_sym), float(fs), p, vb.LPC_AUTOCORR, est.ptr, 4, trk.ptr, None, None, None, vb.F32), "ffA"))
    t["find_formants path B (burg)"] = timeit(lambda: ctx._check(L.vbx_find_formants(
        ctx.h, C.byref(fr_per), float(fs), p, vb.LPC_BURG, est.ptr, 4, trk.ptr, None, None, None, vb.F32), "ffB"))
    print(f"== {name}: F={F} frames (N={N}, hop={hop})")
    for k, ms in t.items():
        print(f"  {k:36s} {ms:9.3f} ms   {F/ms/1e3:10.2f} Mframes/s", flush=True)
    for a in (d, ac, co, st, res, nres, est, trk):
        a.free()
print(ctx.measure_peaks())
