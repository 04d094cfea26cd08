#!/usr/bin/env python
"""GPU box: (1) the oracle against its own rounding variants on the DEVICE-synthesised corpus; (2) the device lag function
(vbx_pitch_lag_function) against the oracle's self_lag; (3) the oracle's refinement run on the DEVICE lag function: does it give
the device's candidates (then the lag function's last bits decide) or the oracle's (then the refinement kernels differ)?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import oracle, voxbox_b200 as vb
from pitch_sensitivity import compare
oracle.build()
U = int(sys.argv[1]) if len(sys.argv) > 1 else 6
threads = max(1, len(os.sched_getaffinity(0)))
c = vb.Context(0)
fs, N, hop, K = 16000, 640, 160, 40
ns = fs * 10
d = c.synth_speech(U, ns, fs, seed=0x5EED, first_utt=7000)
audio = d.to_host()
J = c.n_frames_of(ns, N, hop); F = U * J
ref = [oracle.batch_pitch(audio[u], J, N, hop, oracle.WIN_HANN_SYMMETRIC, float(fs), 0.45, 75.0, 600.0, K, n_threads=threads) for u in range(U)]
rc, rn = np.concatenate([r[0] for r in ref]), np.concatenate([r[1] for r in ref])
for v in (1, 2):
    var = [oracle.batch_pitch_variant(audio[u], J, N, hop, oracle.WIN_HANN_SYMMETRIC, float(fs), 0.45, 75.0, 600.0, v, K, n_threads=threads) for u in range(U)]
    compare(np.concatenate([r[0] for r in var]), np.concatenate([r[1] for r in var]), rc, rn, K, f"oracle[acf variant {v}] vs oracle (device corpus)")
fr = c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
res = c.pitch(fr, float(fs), 0.45, 75.0, 600.0, K)
cand, n = res["candidates"].to_host(), res["n_cand"].to_host()
compare(cand, n, rc, rn, K, "gpu vs oracle")
ylag = c.pitch_lag_function(fr).to_host()
w = oracle.hanning_window(N)
ixmax = N // 2; offset = -ixmax - 1; nx = ixmax - offset
bad = np.nonzero((np.abs(cand[..., 0] - rc[..., 0]) > 0.1).any(axis=1) & (n == rn))[0]
worst = 0.0
np.set_printoptions(linewidth=200, precision=12)
for f in range(0, F, 7):
    u, j = divmod(f, J)
    xw = audio[u, j * hop:j * hop + N].astype(np.float64) * w
    st, oc, ex = oracle.pitch(xw, float(fs), 0.45, 75.0, 600.0, K, want_lag=True)
    worst = max(worst, float(np.max(np.abs(ex["lag"][:N] - ylag[f]))))
print(f"lag function: max |device - oracle| over {len(range(0, F, 7))} frames = {worst:.3e}")
shown = 0
for f in bad[:40]:
    u, j = divmod(f, J)
    xw = audio[u, j * hop:j * hop + N].astype(np.float64) * w
    st, oc, ex = oracle.pitch(xw, float(fs), 0.45, 75.0, 600.0, K, want_lag=True)
    lo = ex["lag"]
    lg = np.concatenate([ylag[f], np.zeros(N)])
    dl = np.abs(lo[:N] - ylag[f])
    # refine every in-range maximum with the ORACLE's improve_extremum on both lag functions
    for ci in range(1, ixmax - 1):
        if not (lo[ci - 1] < lo[ci] and lo[ci + 1] < lo[ci]):
            continue
        outs = []
        for lag in (lo, lg):
            dr = 0.5 * (lag[ci + 1] - lag[ci - 1]); d2r = 2. * lag[ci] - (lag[ci - 1] - lag[ci + 1])
            freq = fs / (ci + dr / d2r)
            if not (75.0 < freq < 600.0):
                outs.append(None); continue
            x, y, _ = oracle.improve_extremum(lag, offset, nx, fs / freq - offset)
            outs.append((fs / (x + offset), y))
        if outs[0] and outs[1] and abs(outs[0][0] - outs[1][0]) > 0.1:
            near = np.min(np.abs(cand[f, :n[f], 0] - outs[1][0]))
            print(f"frame {f} lag-index {ci}: oracle refinement on the oracle lag function -> {outs[0][0]:.8f} Hz / {outs[0][1]:.12f}; on the DEVICE lag function -> "
                  f"{outs[1][0]:.8f} Hz / {outs[1][1]:.12f} (nearest device candidate {near:.2e} Hz away); max |dlag| of the frame {dl.max():.2e}, at the peak {dl[ci]:.2e}")
            shown += 1
    if shown >= 12:
        break
print("frames with list differences:", len(bad), "; cases where the ORACLE's own refinement changes with the device lag function:", shown)
