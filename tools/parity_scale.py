#!/usr/bin/env python
"""Parity at scale (GPU box): every path on >= 1e5 frames of the on-device synthetic corpus (copied device -> host and fed to the
CPU oracle, never regenerated) — reports mismatch counts, never masks them.  Also counts the reference's Laguerre solves
that run into the 20-iteration cap without having converged (polynomial.rs:34-72): the pair-deflation kernel of the fused formant
path is a different solver, so such frames would disagree by construction.
usage: python tools/parity_scale.py [n_utterances_16k] [n_utterances_44k]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import oracle, voxbox_b200 as vb
from pitch_sensitivity import compare
oracle.build()
U = int(sys.argv[1]) if len(sys.argv) > 1 else 101      # 101 x 997 = 100 697 pitch frames
U44 = int(sys.argv[2]) if len(sys.argv) > 2 else 101    # 101 x 998 = 100 798 formant frames at 44.1 kHz
threads = max(1, len(os.sched_getaffinity(0)))
c = vb.Context(0)
fs = 16000
ns = fs * 10
d = c.synth_speech(U, ns, fs, seed=0x5EED, first_utt=500000)
audio = d.to_host()
print(f"corpus: {U} utterances x 10 s at 16 kHz and {U44} at 44.1 kHz, synthesised on the device (seed 0x5EED, utterances 500000+), {threads} CPU threads")
# ---- pitch (C4 shape)
N, hop, K = 640, 160, 40
J = c.n_frames_of(ns, N, hop); F = U * J
fr = c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
res = c.pitch(fr, float(fs), 0.45, 75.0, 600.0, K)
cand, n, st = res["candidates"].to_host(), res["n_cand"].to_host(), res["status"].to_host()
t0 = time.time()
refs = [oracle.batch_pitch(audio[u], J, N, hop, oracle.WIN_HANN_SYMMETRIC, float(fs), 0.45, 75.0, 600.0, K, n_threads=threads) for u in range(U)]
rc = np.concatenate([r[0] for r in refs]); rn = np.concatenate([r[1] for r in refs]); rs = np.concatenate([r[2] for r in refs])
print(f"pitch: {F} frames, oracle {time.time()-t0:.1f} s; status mismatches {int(np.count_nonzero(st != rs))}")
compare(cand, n, rc, rn, K, "pitch gpu vs oracle (full candidate lists)")
# ---- formants: both LPC methods, tracker from the MALE estimates per utterance (C3 semantics at 16 kHz and 44.1 kHz)
MALE = np.array([[f, 1.0] for f in (320., 1440., 2760., 3200.)])
for fs2, N2, hop2, U2 in ((16000, 400, 160, U), (44100, 1102, 441, U44)):
    ns2 = fs2 * 10
    d2 = d if fs2 == fs else c.synth_speech(U2, ns2, fs2, seed=0x5EED, first_utt=600000)
    a2 = audio if fs2 == fs else d2.to_host()
    J2 = c.n_frames_of(ns2, N2, hop2); F2 = U2 * J2
    est = np.tile(MALE, (U2, 1, 1))
    for name, method, win in (("burg (find_formants)", vb.LPC_BURG, vb.WINDOW_HANN_PERIODIC), ("autocorr path A", vb.LPC_AUTOCORR, vb.WINDOW_HANN_SYMMETRIC)):
        fr2 = c.frames(d2.ptr, F2, N2, hop2, win, frames_per_segment=J2, segment_stride=ns2)
        out = c.find_formants(fr2, float(fs2), 12, method, est)
        oracle.laguerre_stats(reset=True)
        tr, nr_, rs_ = [], [], []
        t0 = time.time()
        # utterances in parallel on the oracle: one call over the flattened batch with per-utterance frame ranges
        starts = np.arange(U2, dtype=np.int64) * (ns2 // hop2) if ns2 % hop2 == 0 else None
        if starts is not None:
            offs = np.stack([starts, starts + J2], axis=1).reshape(-1)
            o = oracle.batch_formants(a2.reshape(-1), int(offs[-1]), N2, hop2, win, 0 if method == vb.LPC_BURG else 1, float(fs2), 12, offs, MALE, n_threads=threads)
            idx = (starts[:, None] + np.arange(J2)[None, :]).reshape(-1)
            tr, nr_, rs_ = o["tracks"][idx], o["n_res"][idx], o["resonances"][idx]
        else:
            for u in range(U2):
                o = oracle.batch_formants(a2[u], J2, N2, hop2, win, 0 if method == vb.LPC_BURG else 1, float(fs2), 12, np.array([0, J2]), MALE, n_threads=1)
                tr.append(o["tracks"]); nr_.append(o["n_res"]); rs_.append(o["resonances"])
            tr, nr_, rs_ = np.concatenate(tr), np.concatenate(nr_), np.concatenate(rs_)
        solves, capped, unconv = oracle.laguerre_stats(reset=True)
        dres = np.abs(out["resonances"] - rs_).max(axis=(1, 2))
        dtr = np.abs(out["tracks"] - tr).max(axis=(1, 2))
        print(f"formants fs={fs2} {name}: {F2} frames (oracle {time.time()-t0:.1f} s); n_res mismatches {int(np.count_nonzero(out['n_res'] != nr_))}; "
              f"resonance |d| max {dres.max():.3e} Hz (>0.5: {int(np.count_nonzero(dres > 0.5))}); track |d| max {dtr.max():.3e} Hz (>0.5: {int(np.count_nonzero(dtr > 0.5))}); "
              f"slot assignment differences (any track value off by > 0.5 Hz): {int(np.count_nonzero(dtr > 0.5))}")
        print(f"   reference Laguerre solves {solves}: ran all 20 iterations {capped}, of those NOT converged (last step > 1e-8 max(1,|z|)) {unconv}")
    if d2 is not d: d2.free()
# ---- MFCC (C5 shape)
Nm, hm = 400, 160
Jm = c.n_frames_of(ns, Nm, hm); Fm = U * Jm
frm = c.frames(d.ptr, Fm, Nm, hm, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=Jm, segment_stride=ns)
out = c.mfcc(frm, 40, 133.0, 6855.0, float(fs), n_keep=13).to_host()
ref = np.concatenate([oracle.batch_mfcc(audio[u], Jm, Nm, hm, oracle.WIN_HANN_SYMMETRIC, 40, 133.0, 6855.0, float(fs), n_keep=13, n_threads=threads) for u in range(U)])
err = np.max(np.abs(out - ref), axis=1) / np.max(np.abs(ref), axis=1)
print(f"mfcc: {Fm} frames; norm-wise error max {err.max():.3e} (>1e-5: {int(np.count_nonzero(err > 1e-5))})")
# ---- LPC (C2 shape)
r, ac, _ = c.lpc(frm, 12)
refs = [oracle.batch_lpc(audio[u], Jm, Nm, hm, oracle.WIN_HANN_SYMMETRIC, 12, n_threads=threads) for u in range(U)]
rr, ra = np.concatenate([x[0] for x in refs]), np.concatenate([x[1] for x in refs])
er = np.max(np.abs(r.to_host() - rr), axis=1) / np.max(np.abs(rr), axis=1)
ea = np.max(np.abs(ac.to_host() - ra), axis=1) / np.max(np.abs(ra), axis=1)
print(f"lpc: {Fm} frames; r norm-wise error max {er.max():.3e}; lpc norm-wise error max {ea.max():.3e} (>1e-5: {int(np.count_nonzero(ea > 1e-5))})")
