#!/usr/bin/env python
"""numpy emulation of mfcc_lane5_kernel's transform (vbx_mfcc_lane5.cuh): a 400-sample real frame packed as 200 complex
points, in-place decimation-in-frequency passes of radix 8, 5, 5 done by 5 lanes, the untangle fused into the last pass by
pairing butterfly (kA, kB) with (8-kA, 4-kB).  Checks the index maps against numpy's rfft and counts shared-memory
wavefronts (quarter-warp model for 16-byte accesses) for the frame stride.  CPU only: python tools/mfcc_lane5_emulation.py"""
import numpy as np

MC, N = 200, 400
W = lambda n, e: np.exp(-2j * np.pi * e / n)


def col(b, kB):
    """column of butterfly (b, kB)'s inputs inside block b: blocks 5..7 are mirrored, so the partner's loads are base + s too"""
    return 4 - kB if b >= 5 else kB


def slot(k):
    """where the spectrum's bin k lives after pass C ("kA-major": consecutive bins are 25 slots = one bank group apart)"""
    return MC if k == MC else 25 * (k % 8) + k // 8


def transform(x):
    """the kernel's data flow on one frame: returns (X in natural order read through slot(), per-slot write counts)"""
    z = x[0::2] + 1j * x[1::2]
    buf = np.zeros(MC + 5, complex)
    # pass A: radix 8 over t, butterfly n1 = s + 5 i, output u at 25 u + n1, twiddle W200^(n1 u)
    for s in range(5):
        for i in range(5):
            n1 = s + 5 * i
            v = np.array([z[n1 + 25 * t] for t in range(8)])
            Y = np.array([sum(v[t] * W(8, t * u) for t in range(8)) for u in range(8)])
            for u in range(8):
                buf[25 * u + n1] = Y[u] * W(200, n1 * u)
    # pass B: radix 5 inside every 25-block b, butterfly j = s; output kB stored TRANSPOSED at 25 b + 5 s + col(b, kB)
    nb = buf.copy()
    for b in range(8):
        for s in range(5):
            v = np.array([buf[25 * b + s + 5 * t] for t in range(5)])
            for kB in range(5):
                nb[25 * b + 5 * s + col(b, kB)] = W(25, s * kB) * sum(v[t] * W(5, t * kB) for t in range(5))
    buf = nb
    # pass C + untangle.  butterfly (kA, kB): inputs buf[25 kA + 5 j + col], output kC = Z[kA + 8 kB + 40 kC]
    def bf(kA, kB):
        v = np.array([buf[25 * kA + 5 * j + col(kA, kB)] for j in range(5)])
        return np.array([sum(v[j] * W(5, j * kC) for j in range(5)) for kC in range(5)])
    out = np.zeros(MC + 5, complex)
    def pair(za, zb, k):
        E = 0.5 * (za + np.conj(zb)); O = 0.5 * (za - np.conj(zb)); T = W(N, k) * O
        return E - 1j * T, np.conj(E) - 1j * np.conj(T)
    units = []
    for s in range(5):
        for kA in (1, 2, 3):
            units.append((s, (kA, s), (8 - kA, 4 - s), "regular"))
        units.append((s, [(4, 0), (4, 1), (0, 1), (0, 2), (0, 0)][s], [(4, 4), (4, 3), (0, 4), (0, 3), (4, 2)][s], "singles" if s == 4 else "regular"))
    seen = np.zeros(MC + 5, int)
    for s, c, cp, kind in units:
        A, B = bf(*c), bf(*cp)
        k_of = lambda cc, kC: cc[0] + 8 * cc[1] + 40 * kC
        if kind == "regular":
            ops = [(A[i], B[4 - i], k_of(c, i), k_of(cp, 4 - i)) for i in range(5)]
        else:  # c = (0,0): Z[40 kC]; cp = (4,2): Z[20 + 40 kC]
            ops = [(A[0], A[0], 0, MC), (A[1], A[4], 40, 160), (A[2], A[3], 80, 120), (B[0], B[4], 20, 180), (B[1], B[3], 60, 140), (B[2], B[2], 100, 100)]
        for za, zb, k, k2 in ops:
            assert k + k2 == MC
            xa, xb = pair(za, zb, k)
            out[slot(k)] = xa; out[slot(k2)] = xb
            seen[slot(k)] += 1; seen[slot(k2)] += 1
    X = np.array([out[slot(k)] for k in range(MC + 1)])
    return X, seen


def check(seed=1):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(N)
    X, seen = transform(x)
    ref = np.fft.rfft(x)
    twice = np.flatnonzero(seen > 1).tolist()
    return float(np.max(np.abs(X - ref))), bool(np.all(seen[:MC + 1] >= 1)), twice


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



if __name__ == "__main__":
    err, every, twice = check()
    print("max |X - rfft|:", err, " every slot written:", every, " written twice (bin 100 pairs with itself):", twice)
    pats = {"+s": lambda s: s, "-s": lambda s: 100 - s, "+5s": lambda s: 5 * s, "-5s": lambda s: 100 - 5 * s}
    for FS in range(200, 216):
        print(FS, {k: wavefronts(FS, f) for k, f in pats.items()})

