#!/usr/bin/env python
"""Can the roots kernel run in the shadow of the persistent LPC kernel?  Two contexts (two streams) on one GPU: the C3-shape LPC
launch alone, the pair-deflation roots launch alone (128- and 96-thread CTAs), and both at once (GPU box only).
usage: python tools/corun_probe.py [utterances]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))
import voxbox_b200 as vb  # noqa: E402

U = int(sys.argv[1]) if len(sys.argv) > 1 else 1125
fs, N, hop, p = 44100, 1102, 441, 12
a, b = vb.Context(0), vb.Context(0)
ns = fs * 10
d = a.synth_speech(U, ns, fs, first_utt=100)
J = a.n_frames_of(ns, N, hop)
F = U * J
fr = a.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
ac = a.empty((F, p + 1), np.float64)
La, Lb = a.lib, b.lib


def lpc():
    a._check(La.vbx_lpc(a.h, C.byref(fr), p, None, ac.ptr, None, vb.F64), "lpc")


lpc()
a.sync()
res = b.empty((F, p, 2), np.float32)
nres = b.empty((F,), np.int32)
st = b.empty((F,), np.uint8)


FR = [F]


def roots():
    b._check(Lb.vbx_lpc_to_resonances(b.h, ac.ptr, vb.F64, FR[0], p + 1, p, 1, float(fs), 1, None, res.ptr, p, nres.ptr, None, st.ptr, vb.F32, 0), "roots")


def wall(fns, reps=5):
    for f in fns:
        f()
    a.sync(); b.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        for f in fns:
            f()
    a.sync(); b.sync()
    return (time.perf_counter() - t0) / reps * 1e3


print(f"{F} frames ({U} utterances, N={N}, hop={hop})")
print(f"lpc alone                      {wall([lpc]):7.3f} ms")
for frac in (1.0, 0.25, 0.05):
    FR[0] = int(F * frac)
    for thr in ("128", "96"):
        os.environ["VBX_ROOTS_THREADS"] = thr
        tr = wall([roots])
        tb = wall([lpc, roots])
        tc = wall([roots, lpc])
        print(f"roots on {frac:4.0%} of the frames, {thr:>3}-thread CTAs: alone {tr:7.3f} ms;  lpc then roots {tb:7.3f} ms;  roots then lpc {tc:7.3f} ms")

# do the two contexts' streams overlap at all?  the same under-filling roots launch (1 % of the frames) on one and on both contexts
res2 = a.empty((F, p, 2), np.float32); nres2 = a.empty((F,), np.int32); st2 = a.empty((F,), np.uint8)
FR[0] = int(F * 0.01)


def roots_a():
    a._check(La.vbx_lpc_to_resonances(a.h, ac.ptr, vb.F64, FR[0], p + 1, p, 1, float(fs), 1, None, res2.ptr, p, nres2.ptr, None, st2.ptr, vb.F32, 0), "roots")


os.environ["VBX_ROOTS_THREADS"] = "128"
print(f"roots on 1 %: context b alone {wall([roots], 20):.3f} ms, context a alone {wall([roots_a], 20):.3f} ms, both {wall([roots_a, roots], 20):.3f} ms")
