#!/usr/bin/env python
"""numpy emulation of mfcc_lane5_kernel's transform (vbx_mfcc_lane5.cuh): a 400-sample real frame packed as 200 complex
points, in-place decimation-in-frequency passes of radix 8, 5, 5 done by 5 lanes, the untangle fused into the last pass by
pairing butterfly (kA, kB) with (8-kA, 4-kB).  Checks the index maps against numpy's rfft and counts shared-memory
wavefronts (quarter-warp model for 16-byte accesses) for the frame stride.  CPU only: python tools/mfcc_lane5_emulation.py"""
import numpy as np

MC, N = 200, 400
W = lambda n, e: np.exp(-2j * np.pi * e / n)


def transform(x):
    z = x[0::2] + 1j * x[1::2]
    buf = np.zeros(MC + 1, complex)
    # pass A: radix 8 over t, butterfly n1 = s + 5 i, output u at 25 u + n1, twiddle W200^(n1 u)
    for s in range(5):
        for i in range(5):
            n1 = s + 5 * i
            v = np.array([z[n1 + 25 * t] for t in range(8)])
            Y = np.array([sum(v[t] * W(8, t * u) for t in range(8)) for u in range(8)])
            for u in range(8):
                buf[25 * u + n1] = Y[u] * W(200, n1 * u)
    # pass B: radix 5 inside every 25-block b, butterfly j = s; output kB stored TRANSPOSED at 25 b + 5 s + kB
    nb = buf.copy()
    for b in range(8):
        for s in range(5):
            v = np.array([buf[25 * b + s + 5 * t] for t in range(5)])
            for kB in range(5):
                nb[25 * b + 5 * s + kB] = W(25, s * kB) * sum(v[t] * W(5, t * kB) for t in range(5))
    buf = nb
    # pass C + untangle.  butterfly (kA, kB): inputs buf[25 kA + 5 j + kB], output kC = Z[kA + 8 kB + 40 kC]
    def bf(kA, kB):
        v = np.array([buf[25 * kA + 5 * j + kB] for j in range(5)])
        return np.array([sum(v[j] * W(5, j * kC) for j in range(5)) for kC in range(5)])
    slot = lambda kA, kB, kC: 25 * kA + 5 * kC + kB   # where X_k lands (the slot butterfly (kA, kB) read input j = kC from)
    out = np.zeros(MC + 1, complex)
    def pair(za, zb, k):
        E = 0.5 * (za + np.conj(zb)); O = 0.5 * (za - np.conj(zb)); T = W(N, k) * O
        return E - 1j * T, np.conj(E) - 1j * np.conj(T)
    units = []
    for s in range(5):
        for kA in (1, 2, 3):
            units.append((s, (kA, s), (8 - kA, 4 - s), "regular"))
        units.append((s, [(4, 0), (4, 1), (0, 1), (0, 2), (0, 0)][s], [(4, 4), (4, 3), (0, 4), (0, 3), (4, 2)][s], "singles" if s == 4 else "regular"))
    seen = np.zeros(MC + 1, int)
    for s, c, cp, kind in units:
        A, B = bf(*c), bf(*cp)
        k_of = lambda cc, kC: cc[0] + 8 * cc[1] + 40 * kC
        if kind == "regular":
            ops = [(A[i], B[4 - i], k_of(c, i), slot(*c, i), slot(*cp, 4 - i)) for i in range(5)]
        else:  # c = (0,0): Z[40 kC]; cp = (4,2): Z[20 + 40 kC]
            ops = [(A[0], A[0], 0, slot(*c, 0), MC), (A[1], A[4], 40, slot(*c, 1), slot(*c, 4)), (A[2], A[3], 80, slot(*c, 2), slot(*c, 3)),
                   (B[0], B[4], 20, slot(*cp, 0), slot(*cp, 4)), (B[1], B[3], 60, slot(*cp, 1), slot(*cp, 3)), (B[2], B[2], 100, slot(*cp, 2), slot(*cp, 2))]
        for za, zb, k, sa, sb in ops:
            xa, xb = pair(za, zb, k)
            out[sa] = xa; out[sb] = xb
            seen[sa] += 1; seen[sb] += 1
    # natural-order view through the slot map: bin k = kA + 8 kB + 40 kC at 25 kA + 5 kC + kB, Nyquist at MC
    X = np.zeros(MC + 1, complex)
    for k in range(MC):
        kA, r = k % 8, k // 8
        X[k] = out[25 * kA + 5 * (r // 5) + (r % 5)]
    X[MC] = out[MC]
    return X, seen


rng = np.random.default_rng(1)
x = rng.standard_normal(N)
X, seen = transform(x)
ref = np.fft.rfft(x)
print("max |X - rfft|:", np.max(np.abs(X - ref)), " every slot written:", bool(np.all(seen >= 1)), " written twice:", np.flatnonzero(seen > 1).tolist())


# ---- shared-memory wavefronts of a warp-wide 16-byte access: lane = 5 q + s (30 lanes), element index q FS + f(s) ------------
def wavefronts(FS, f):
    tot = 0
    lanes = [(q, s) for q in range(6) for s in range(5)]
    for qw in range(0, 32, 8):
        grp = lanes[qw:qw + 8]
        per_bank = {}
        for q, s in grp:
            e = q * FS + f(s)
            per_bank.setdefault(e % 8, set()).add(e)
        tot += max((len(v) for v in per_bank.values()), default=0)
    return tot


pats = {"+s": lambda s: s, "-s": lambda s: 100 - s, "+5s": lambda s: 5 * s, "-5s": lambda s: 100 - 5 * s}
for FS in range(200, 216):
    print(FS, {k: wavefronts(FS, f) for k, f in pats.items()})
