import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, "vox_box.rs_b200/python")
import voxbox_b200 as vb
from voxbox_b200 import synth
ctx = vb.Context(0)
L = ctx.lib
fs, N, hop, p = 44100, 1102, 441, 12
base = synth.corpus(24, fs, 10.0)
for U in [int(u) for u in os.environ.get("ULIST", "120,240,480,1125").split(",")]:
    audio = np.tile(base, (-(-U // 24), 1))[:U]
    ns = audio.shape[1]
    J = ctx.n_frames_of(ns, N, hop); F = U * J
    d = ctx.to_device(audio)
    fr = ctx.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
    ac = ctx.empty((F, p + 1), np.float64); res = ctx.empty((F, p, 2), np.float32); st = ctx.empty((F,), np.uint8)
    ctx._check(L.vbx_lpc(ctx.h, C.byref(fr), p, None, ac.ptr, None, vb.F64), "lpc")
    for prec in (0, 1):
        f = lambda: ctx._check(L.vbx_lpc_to_resonances(ctx.h, ac.ptr, vb.F64, F, p + 1, p, 1, float(fs), 1, None, res.ptr, p, None, None, st.ptr, vb.F32, prec), "roots")
        f(); ctx.sync(); ctx.timer_start()
        for _ in range(3): f()
        ms = ctx.timer_stop_ms() / 3
        print(f"U={U} F={F} prec={prec}: {ms:.3f} ms  {F/ms/1e3:.1f} Mframes/s", flush=True)
    for a in (d, ac, res, st): a.free()
