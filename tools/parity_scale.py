#!/usr/bin/env python
"""Parity at scale (GPU box): pitch / formants / MFCC on U synthetic utterances vs the CPU oracle; reports mismatch
counts (never masks them).  usage: python tools/parity_scale.py [U]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python")); sys.path.insert(0, ROOT)
import oracle, voxbox_b200 as vb
from voxbox_b200 import synth
oracle.build()
U = int(sys.argv[1]) if len(sys.argv) > 1 else 16
threads = max(1, len(os.sched_getaffinity(0)))
c = vb.Context(0)
fs = 16000
audio = synth.corpus(U, fs, 10.0, first=1000)
ns = audio.shape[1]
d = c.to_device(audio)
# ---- pitch (C4 shape)
N, hop, K = 640, 160, 40
J = c.n_frames_of(ns, N, hop); F = U * J
fr = c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
res = c.pitch(fr, float(fs), 0.45, 75.0, 600.0, K)
cand, n, st = res["candidates"].to_host(), res["n_cand"].to_host(), res["status"].to_host()
t0 = time.time()
refs = [oracle.batch_pitch(audio[u], J, N, hop, oracle.WIN_HANN_SYMMETRIC, float(fs), 0.45, 75.0, 600.0, K, n_threads=threads) for u in range(U)]
rc = np.concatenate([r[0] for r in refs]); rn = np.concatenate([r[1] for r in refs]); rs = np.concatenate([r[2] for r in refs])
print(f"pitch: {F} frames, oracle {time.time()-t0:.1f} s on {threads} threads")
print("  status mismatches:", int(np.count_nonzero(st != rs)), " candidate-count mismatches:", int(np.count_nonzero(n != rn)))
print("  voiced/unvoiced flips:", int(np.count_nonzero((cand[:, 0, 0] != 0) != (rc[:, 0, 0] != 0))))
dtop = np.abs(cand[:, 0, 0] - rc[:, 0, 0])
print(f"  top candidate |df|: max {dtop.max():.3e} Hz, >0.1 Hz: {int(np.count_nonzero(dtop > 0.1))}, >1e-3 Hz: {int(np.count_nonzero(dtop > 1e-3))}")
ok = n == rn
k = np.minimum(n, K); mask = (np.arange(K)[None, :] < k[:, None]) & ok[:, None]
dall = np.abs(cand[..., 0] - rc[..., 0])[mask]; dstr = np.abs(cand[..., 1] - rc[..., 1])[mask]
print(f"  all candidates ({mask.sum()}): |df| max {dall.max():.3e}, >0.1 Hz: {int(np.count_nonzero(dall > 0.1))}; |dstrength| max {dstr.max():.3e}, >1e-6: {int(np.count_nonzero(dstr > 1e-6))}")
bad = (np.abs(cand[..., 0] - rc[..., 0]) > 0.1) & mask
strong = bad & (np.maximum(cand[..., 1], rc[..., 1]) > 0.45)
print(f"  of the {int(bad.sum())} list positions that differ by > 0.1 Hz, {int(strong.sum())} involve a candidate stronger than the unvoiced threshold 0.45;"
      f" max strength involved {float(np.max(np.where(bad, np.maximum(cand[..., 1], rc[..., 1]), -9))):.3f}")
# ---- formants: both LPC methods, tracker from the MALE estimates per utterance (C3 semantics at 16 kHz and 44.1 kHz)
for fs2, N2, hop2 in ((16000, 400, 160), (44100, 1102, 441)):
    U2 = max(2, U // 4)
    a2 = synth.corpus(U2, fs2, 10.0, first=2000)
    ns2 = a2.shape[1]
    d2 = c.to_device(a2)
    J2 = c.n_frames_of(ns2, N2, hop2); F2 = U2 * J2
    est = np.tile(np.array([[f, 1.0] for f in (320., 1440., 2760., 3200.)]), (U2, 1, 1))
    for name, method, win in (("burg (find_formants)", vb.LPC_BURG, vb.WINDOW_HANN_PERIODIC), ("autocorr path A", vb.LPC_AUTOCORR, vb.WINDOW_HANN_SYMMETRIC)):
        fr2 = c.frames(d2.ptr, F2, N2, hop2, win, frames_per_segment=J2, segment_stride=ns2)
        out = c.find_formants(fr2, float(fs2), 12, method, est)
        tr, nr_, rs_ = [], [], []
        for u in range(U2):
            o = oracle.batch_formants(a2[u], J2, N2, hop2, win, 0 if method == vb.LPC_BURG else 1, float(fs2), 12, np.array([0, J2]), est[0], n_threads=1)
            tr.append(o["tracks"]); nr_.append(o["n_res"]); rs_.append(o["resonances"])
        tr, nr_, rs_ = np.concatenate(tr), np.concatenate(nr_), np.concatenate(rs_)
        dres = np.abs(out["resonances"] - rs_).max(axis=(1, 2))
        dtr = np.abs(out["tracks"] - tr).max(axis=(1, 2))
        print(f"formants fs={fs2} {name}: {F2} frames; n_res mismatches {int(np.count_nonzero(out['n_res'] != nr_))}; resonance |d| max {dres.max():.3e} Hz (>0.5: {int(np.count_nonzero(dres > 0.5))}); "
              f"track |d| max {dtr.max():.3e} Hz (>0.5: {int(np.count_nonzero(dtr > 0.5))})")
    d2.free()
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
