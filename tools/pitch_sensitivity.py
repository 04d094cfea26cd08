#!/usr/bin/env python
"""How well-defined is the reference's pitch candidate LIST?  (CPU only: the oracle against itself.)

Pitched::pitch (periodic.rs:396-456) refines every local maximum of the lag function with a Brent search that, as
written, MINIMISES the sinc interpolant over (ixmid - 1, ixmid + 1) (brent_maximize, periodic.rs:103-188, is a
minimiser of +f for is_max = true) with termination tolerance sqrt(eps)·|x| + 1e-10/3 — about 5e-6 samples at the lags in
play.  Around a maximum that search runs towards a bracket edge or a shallow local minimum; its path (golden vs parabolic
steps, `ft <= fx` comparisons) depends on the last bits of the interpolant.  This script measures how far the RETURNED LIST
moves when only the rounding of the autocorrelation fold changes — the same sums added in descending order (variant 1) or
with fused multiply-adds (variant 2): perturbations of ~1e-16·r[0], which no implementation other than a bit-for-bit copy of
the reference's scalar loop can avoid.  It reports, per variant:
  * positional mismatches (list position k of the variant vs position k of the reference), the metric of tools/parity_scale.py,
  * set-wise mismatches (every candidate matched to the nearest-frequency candidate of the other list), which separates
    re-ORDERING of near-equal strengths from candidates that actually moved.
usage: python tools/pitch_sensitivity.py [n_utterances]
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python")); sys.path.insert(0, ROOT)
import oracle
from voxbox_b200 import synth


def compare(cand, n, rc, rn, K, label):
    ok = n == rn
    k = np.minimum(n, K)
    mask = (np.arange(K)[None, :] < k[:, None]) & ok[:, None]
    df = np.abs(cand[..., 0] - rc[..., 0])
    ds = np.abs(cand[..., 1] - rc[..., 1])
    bad = (df > 0.1) & mask
    print(f"{label}: {int(mask.sum())} list positions in {len(n)} frames; count mismatches {int(np.count_nonzero(~ok))}; "
          f"voiced/unvoiced flips {int(np.count_nonzero((cand[:, 0, 0] != 0) != (rc[:, 0, 0] != 0)))}; "
          f"top candidate max |df| {np.max(np.abs(cand[:, 0, 0] - rc[:, 0, 0])):.3e} Hz")
    print(f"   positional: {int(bad.sum())} positions differ by > 0.1 Hz (max {df[mask].max():.3e} Hz), |dstrength| max {ds[mask].max():.3e}, "
          f"> 1e-6: {int(np.count_nonzero(ds[mask] > 1e-6))}")
    # set-wise: match every candidate of the variant to the nearest frequency of the reference list of the same frame
    moved = moved_strong = 0
    worst = 0.0
    swaps = 0
    for f in np.nonzero(bad.any(axis=1))[0]:
        a, b = cand[f, :k[f]], rc[f, :k[f]]
        d = np.abs(a[:, None, 0] - b[None, :, 0])
        j = d.argmin(axis=1)
        dmin = d[np.arange(len(a)), j]
        m = dmin > 0.1
        moved += int(m.sum())
        moved_strong += int(np.count_nonzero(m & (a[:, 1] > 0.45)))
        worst = max(worst, float(dmin.max()))
        swaps += int(np.count_nonzero(bad[f, :k[f]] & ~m))
    print(f"   set-wise  : of those, {swaps} are re-orderings of candidates whose strengths differ by less than the search tolerance resolves; "
          f"{moved} candidates have no partner within 0.1 Hz (max distance {worst:.3e} Hz), {moved_strong} of them above the unvoiced threshold")
    return int(bad.sum()), moved


if __name__ == "__main__":
    oracle.build()
    U = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    threads = max(1, len(os.sched_getaffinity(0)))
    fs, N, hop, K = 16000, 640, 160, 40
    audio = synth.corpus(U, fs, 10.0, first=1000)
    J = oracle.n_frames_of(audio.shape[1], N, hop)
    t0 = time.time()
    ref = [oracle.batch_pitch(audio[u], J, N, hop, oracle.WIN_HANN_SYMMETRIC, float(fs), 0.45, 75.0, 600.0, K, n_threads=threads) for u in range(U)]
    rc, rn = np.concatenate([r[0] for r in ref]), np.concatenate([r[1] for r in ref])
    print(f"reference: {U * J} frames in {time.time() - t0:.1f} s on {threads} threads")
    for v, name in ((1, "autocorrelation terms added in descending order"), (2, "autocorrelation with fused multiply-adds")):
        var = [oracle.batch_pitch_variant(audio[u], J, N, hop, oracle.WIN_HANN_SYMMETRIC, float(fs), 0.45, 75.0, 600.0, v, K, n_threads=threads)
               for u in range(U)]
        vc, vn = np.concatenate([r[0] for r in var]), np.concatenate([r[1] for r in var])
        compare(vc, vn, rc, rn, K, f"oracle[{name}] vs oracle")
