#!/usr/bin/env python
"""GPU box: full-list pitch parity vs the oracle for every refinement-kernel variant (VBX_PITCH_REFINE = v0 | v1 | default) and both
lag sweeps, positional and set-wise; dumps example frames.  usage: python tools/pitch_parity_diag.py [n_utterances]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import oracle, voxbox_b200 as vb
from pitch_sensitivity import compare
oracle.build()
U = int(sys.argv[1]) if len(sys.argv) > 1 else 12
threads = max(1, len(os.sched_getaffinity(0)))
c = vb.Context(0)
fs, N, hop, K = 16000, 640, 160, 40
ns = fs * 10
d = c.synth_speech(U, ns, fs, seed=0x5EED, first_utt=7000)
audio = d.to_host()
J = c.n_frames_of(ns, N, hop); F = U * J
fr = c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
ref = [oracle.batch_pitch(audio[u], J, N, hop, oracle.WIN_HANN_SYMMETRIC, float(fs), 0.45, 75.0, 600.0, K, n_threads=threads) for u in range(U)]
rc, rn = np.concatenate([r[0] for r in ref]), np.concatenate([r[1] for r in ref])
first = None
for lag in ("f64", "f32"):
    for refine in ("default", "v1", "v0"):
        os.environ["VBX_PITCH_LAG"] = lag
        if refine == "default": os.environ.pop("VBX_PITCH_REFINE", None)
        else: os.environ["VBX_PITCH_REFINE"] = refine
        res = c.pitch(fr, float(fs), 0.45, 75.0, 600.0, K)
        cand, n = res["candidates"].to_host(), res["n_cand"].to_host()
        compare(cand, n, rc, rn, K, f"gpu[lag {lag}, refine {refine}] vs oracle")
        if first is None: first = (cand, n)
cand, n = first
bad = np.nonzero((np.abs(cand[..., 0] - rc[..., 0]) > 0.1).any(axis=1) & (n == rn))[0]
np.set_printoptions(linewidth=200, precision=9, suppress=True)
for f in bad[:6]:
    k = min(n[f], K)
    print(f"--- frame {f}: {k} candidates (gpu freq, gpu strength | oracle freq, oracle strength)")
    for i in range(k):
        flag = " <--" if abs(cand[f, i, 0] - rc[f, i, 0]) > 0.1 else ""
        print(f"   {cand[f, i, 0]:14.8f} {cand[f, i, 1]:.12f} | {rc[f, i, 0]:14.8f} {rc[f, i, 1]:.12f}{flag}")
