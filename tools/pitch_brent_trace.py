#!/usr/bin/env python
"""GPU box: where do the device refinement and the oracle part ways?  For the candidates of frames whose lists differ, run the
reference's Brent loop in Python twice — once with the oracle's interpolate_sinc, once with the device's (vbx_interpolate_sinc) —
on the ORACLE's lag function, and print the first evaluation at which the two differ by more than 1e-12 or take a different branch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python")); sys.path.insert(0, ROOT)
import oracle, voxbox_b200 as vb
oracle.build()
c = vb.Context(0)
fs, N, hop, K = 16000, 640, 160, 40
ns = fs * 10
U = 3
d = c.synth_speech(U, ns, fs, seed=0x5EED, first_utt=7000)
audio = d.to_host()
J = c.n_frames_of(ns, N, hop)
w = oracle.hanning_window(N)
ixmax = N // 2
offset = -ixmax - 1
nx = ixmax - offset


def brent(f, a, b, tol=1e-10):
    golden = 1. - 0.6180339887498948482045868343656381177203091798057628621
    EPS = np.finfo(float).eps
    sq = np.sqrt(EPS)
    trace = []
    v = a + golden * (b - a)
    fv = f(v); trace.append((v, fv))
    x = w_ = v
    fx = fw = fv
    for it in range(1, 61):
        rng = b - a
        mid = (a + b) * 0.5
        tol_act = sq * abs(x) + tol / 3.
        if abs(x - mid) + rng * 0.5 <= 2. * tol_act:
            return x, fx, trace
        new_step = golden * (b - x) if x < mid else golden * (a - x)
        if abs(x - w_) >= tol_act:
            t = (x - w_) * (fx - fv)
            q = (x - v) * (fx - fw)
            p = (x - v) * q - (x - w_) * t
            q = 2. * q - t
            if q > 0.: p = -p
            else: q = -q
            if abs(p) < abs(new_step * q) and p > q * (a - x + 2. * tol_act) and p < q * (b - x - 2. * tol_act):
                new_step = p / q
        if abs(new_step) < tol_act:
            new_step = tol_act if new_step > 0. else -tol_act
        t = x + new_step
        ft = f(t); trace.append((t, ft))
        if ft <= fx:
            if t < x: b = x
            else: a = x
            v, w_, x = w_, x, t
            fv, fw, fx = fw, fx, ft
        else:
            if t < x: a = t
            else: b = t
            if ft <= fw or abs(w_ - x) < EPS:
                v, w_ = w_, t
                fv, fw = fw, ft
            elif ft <= fv or abs(v - x) < EPS or abs(v - w_) < EPS:
                v = t
                fv = ft
    return x, fx, trace


shown = 0
for u in range(U):
    for j in range(J):
        xw = audio[u, j * hop:j * hop + N].astype(np.float64) * w
        st, cand, ex = oracle.pitch(xw, float(fs), 0.45, 75.0, 600.0, K, want_lag=True)
        lag = ex["lag"]
        # local maxima + parabolic start, as the reference
        for ci in range(1, ixmax - 1):
            if not (lag[ci - 1] < lag[ci] and lag[ci + 1] < lag[ci]):
                continue
            dr = 0.5 * (lag[ci + 1] - lag[ci - 1])
            d2r = 2. * lag[ci] - (lag[ci - 1] - lag[ci + 1])
            freq = fs / (ci + dr / d2r)
            if not (freq > 75.0 and freq < 600.0):
                continue
            ixmid = fs / freq - offset
            fo = lambda t: oracle.interpolate_sinc(lag, offset, nx, t, 1200)
            fg = lambda t: float(c.interpolate_sinc(lag, offset, nx, np.array([t]), 1200)[0, 0])
            xo, yo, tro = brent(fo, ixmid - 1., ixmid + 1.)
            xg, yg, trg = brent(fg, ixmid - 1., ixmid + 1.)
            xr, yr = oracle.improve_extremum(lag, offset, nx, ixmid)[:2]
            if abs(xo - xg) > 1e-3:
                print(f"utt {u} frame {j} lag-index {ci}: ixmid {ixmid + offset:.9f}: oracle-f Brent -> x {xo + offset:.9f} f {yo:.12f} ({len(tro)} evals); "
                      f"device-f Brent -> x {xg + offset:.9f} f {yg:.12f} ({len(trg)} evals); oracle improve_extremum -> x {xr + offset:.9f} f {yr:.12f}")
                for i, ((ta, fa), (tb, fb)) in enumerate(zip(tro, trg)):
                    mark = ""
                    if ta != tb: mark = "  <-- different abscissa"
                    elif abs(fa - fb) > 1e-12: mark = "  <-- values differ"
                    print(f"     eval {i:2d}: t {ta + offset:.12f} f_oracle {fa:+.15f} | t {tb + offset:.12f} f_device {fb:+.15f}  diff {fa - fb:+.3e}{mark}")
                    if ta != tb:
                        break
                shown += 1
                if shown >= 5:
                    sys.exit(0)
