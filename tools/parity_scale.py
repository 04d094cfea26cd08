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
