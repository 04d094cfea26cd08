#!/usr/bin/env python
"""fp32 numpy emulation of lpc_roots_pair_kernel (vbx_roots_kernel.cuh): Laguerre with the reference's update formula and
n = P, conjugate-pair / real-root deflation of the real working polynomial, two fp64 Newton steps on the original
coefficients, from_root.  Runs on the CPU (oracle for the LPC coefficients, numpy.roots as ground truth) and prints, per
sample rate, the Horner work and the resonance parity for a sweep of start points and convergence thresholds — the evidence
behind the kernel's start point (0.3 + 0.9i instead of the reference's -2-2i) and threshold (1e-5 instead of 3e-7).
usage: python tools/roots_pair_emulation.py [utterances-per-rate]"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from voxbox_b200 import synth  # noqa: E402

warnings.filterwarnings("ignore")
f32, c64 = np.float32, np.complex64


def laguerre_step(a0, a1, a2, nn, nref):
    """laguerre_step<float, FAST> of the kernel (polynomial.rs:48-69 with one reciprocal per division)."""
    inv0 = f32(1) / f32(abs(a0) ** 2)
    ca = c64(-(a1 * np.conj(a0)) * inv0)
    ca2 = c64(ca * ca)
    cb = c64(ca2 - c64(2 * (a2 * np.conj(a0)) * inv0))
    c1 = np.sqrt(c64(nn * cb - ca2))
    cc1, cc2 = c64(ca + c1), c64(ca - c1)
    n1, n2 = abs(cc1) ** 2, abs(cc2) ** 2
    den, nd = (cc1, n1) if n1 > n2 else (cc2, n2)
    return c64(nref * np.conj(den) / nd)


def pair_roots(a_asc, z0, eps):
    """Roots of the real polynomial a_asc (ascending powers); returns (roots, Horner coefficient steps, capped solves)."""
    P = len(a_asc) - 1
    c = np.array(a_asc, dtype=f32)
    M, work, capped, roots = P, 0, 0, []
    nn, nref = f32((P - 1) * P), f32(P)
    while M >= 3:
        z, conv = c64(z0), False
        for _ in range(20):
            a0, a1, a2 = c64(c[M]), c64(0), c64(0)
            for j in range(M - 1, -1, -1):
                a2 = c64(a2 * z + a1)
                a1 = c64(a1 * z + a0)
                a0 = c64(a0 * z + c[j])
            work += M
            if abs(a0) ** 2 <= 1e-32:
                conv = True
                break
            st = laguerre_step(a0, a1, a2, nn, nref)
            z = c64(z + st)
            if abs(st) ** 2 <= eps ** 2 * abs(z) ** 2:
                conv = True
                break
        capped += 0 if conv else 1
        if abs(z.imag) <= 1e-5 * abs(z.real):           # real root: divide x - r out
            roots.append(complex(z.real, 0))
            carry, c[M] = c[M], 0
            for i in range(M - 1, -1, -1):
                old, c[i] = c[i], carry
                carry = f32(carry * z.real + old)
            M -= 1
        else:                                           # conjugate pair: divide x^2 - 2 Re(z) x + |z|^2 out
            roots += [complex(z.real, abs(z.imag)), complex(z.real, -abs(z.imag))]
            pq, qq, b1, b2 = f32(-2 * z.real), f32(abs(z) ** 2), f32(0), f32(0)
            for i in range(M, 1, -1):
                b0 = f32(-pq * b1 - qq * b2 + c[i])
                c[i], b2, b1 = b2, b1, b0
            c[1], c[0] = b2, b1
            M -= 2
    if M == 2:
        q0, q1, q2 = c[0], c[1], c[2]
        disc, inv = q1 * q1 - 4 * q2 * q0, 1 / (2 * q2)
        sq = np.sqrt(abs(disc))
        roots += ([complex(-q1 * inv, abs(sq * inv)), complex(-q1 * inv, -abs(sq * inv))] if disc < 0
                  else [complex((-q1 + sq) * inv, 0), complex((-q1 - sq) * inv, 0)])
    elif M == 1:
        roots.append(complex(-c[0] / c[1], 0))
    return roots, work, capped


def pair_roots_real(a_asc, z0, eps):
    """The same solver with the polynomial evaluated in REAL arithmetic: P, P', P''/2 at z from three chained synthetic
    divisions by x^2 - 2 Re(z) x + |z|^2 (6 real multiply-adds per coefficient instead of complex Horner's 12), and the
    deflation done by one more sweep at the converged root whose quotient (the first division's b) becomes the working
    polynomial.  Returns (roots, coefficient steps incl. the deflation sweeps, capped solves)."""
    P = len(a_asc) - 1
    c = np.array(a_asc, dtype=f32)
    M, work, capped, roots = P, 0, 0, []
    nn, nref = f32((P - 1) * P), f32(P)

    def sweep(c, M, p, q):
        b = np.zeros(M + 3, dtype=f32)
        d = np.zeros(M + 3, dtype=f32)
        e = np.zeros(M + 3, dtype=f32)
        for k in range(M, -1, -1):
            b[k] = f32(f32(c[k] + f32(p * b[k + 1])) - f32(q * b[k + 2])) if False else f32(np.float32(c[k]) + np.float32(p) * b[k + 1] - np.float32(q) * b[k + 2])
            if k >= 2:
                d[k - 2] = f32(b[k] + p * d[k - 1] - q * d[k])
            if k >= 4:
                e[k - 4] = f32(d[k - 2] + p * e[k - 3] - q * e[k - 2])
        return b, d, e

    while M >= 3:
        z, conv = c64(z0), False
        for _ in range(20):
            x0, y0 = f32(z.real), f32(z.imag)
            p, q = f32(2 * x0), f32(x0 * x0 + y0 * y0)
            b, d, e = sweep(c, M, p, q)
            work += M
            a0 = c64(complex(b[0] - x0 * b[1], y0 * b[1]))
            Qz = c64(complex(d[0] - x0 * d[1], y0 * d[1]))
            Sz = c64(complex(e[0] - x0 * e[1], y0 * e[1]))
            a1 = c64(2j * y0 * Qz + b[1])
            a2 = c64(Qz - 4 * y0 * y0 * Sz + 2j * y0 * d[1])
            if abs(a0) ** 2 <= 1e-32:
                conv = True
                break
            st = laguerre_step(a0, a1, a2, nn, nref)
            z = c64(z + st)
            if abs(st) ** 2 <= eps ** 2 * abs(z) ** 2:
                conv = True
                break
        capped += 0 if conv else 1
        work += M  # the deflation sweep
        if abs(z.imag) <= 1e-5 * abs(z.real):
            roots.append(complex(z.real, 0))
            b, _, _ = sweep(c, M, f32(z.real), f32(0))
            c = np.concatenate([b[1:M + 1], np.zeros(1, dtype=f32)]).astype(f32)
            M -= 1
        else:
            roots += [complex(z.real, abs(z.imag)), complex(z.real, -abs(z.imag))]
            x0, y0 = f32(z.real), f32(z.imag)
            b, _, _ = sweep(c, M, f32(2 * x0), f32(x0 * x0 + y0 * y0))
            c = np.concatenate([b[2:M + 1], np.zeros(2, dtype=f32)]).astype(f32)
            M -= 2
    if M == 2:
        q0, q1, q2 = c[0], c[1], c[2]
        disc, inv = q1 * q1 - 4 * q2 * q0, 1 / (2 * q2)
        sq = np.sqrt(abs(disc))
        roots += ([complex(-q1 * inv, abs(sq * inv)), complex(-q1 * inv, -abs(sq * inv))] if disc < 0
                  else [complex((-q1 + sq) * inv, 0), complex((-q1 - sq) * inv, 0)])
    elif M == 1:
        roots.append(complex(-c[0] / c[1], 0))
    return roots, work, capped


def resonances(roots, a_desc, fs, polish):
    out = []
    for z in roots:
        if not z.imag > 0:
            continue
        zz = complex(z)
        for _ in range(polish):                          # Newton on the ORIGINAL polynomial, fp64
            p0, p1 = a_desc[0] + 0j, 0j
            for cj in a_desc[1:]:
                p1 = p1 * zz + p0
                p0 = p0 * zz + cj
            if p1 == 0:
                break
            zz = zz - p0 / p1
        f = fs / (2 * np.pi) * np.arctan2(zz.imag, zz.real)
        bw = fs / (2 * np.pi) * abs(np.log(abs(zz) ** 2))
        if 50 < f < fs / 2 - 50:
            out.append((f, bw))
    return sorted(out)


def main():
    n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    for fs, N, hop in ((16000, 400, 160), (44100, 1102, 441)):
        lpcs = []
        for u in range(n_utts):
            x = synth.utterance(40 + u, fs, seconds=2.0)
            F = (x.size - N) // hop + 1
            _, ac = oracle.batch_lpc(x, F, N, hop, oracle.WIN_HANN_SYMMETRIC, 12)
            lpcs += [ac[f] for f in range(F)]
        truth = [resonances([complex(z) for z in np.roots(a)], a, fs, 0) for a in lpcs]
        print(f"fs = {fs}: {len(lpcs)} LPC-12 polynomials")
        real = len(sys.argv) > 2 and sys.argv[2] == "real"
        if real:
            print("  (real-arithmetic evaluation: three chained synthetic divisions + a deflation sweep per root)")
        for z0, eps in ((-2 - 2j, 3e-7), (0 + 1j, 3e-7), (0.7 + 0.7j, 3e-7), (0.3 + 0.9j, 3e-7), (0.3 + 0.9j, 1e-5),
                        (0.3 + 0.9j, 1e-4), (0.3 + 0.9j, 1e-3)):
            work = capped = mism = 0
            worst = 0.0
            for a, ref in zip(lpcs, truth):
                roots, w, cp = (pair_roots_real if real else pair_roots)(a[::-1].copy(), z0, eps)
                work += w
                capped += cp
                got = resonances(roots, a, fs, 2)
                if len(got) != len(ref):
                    mism += 1
                elif got:
                    worst = max(worst, max(max(abs(g[0] - r[0]), abs(g[1] - r[1])) for g, r in zip(got, ref)))
            print(f"  start {z0!s:>12} eps {eps:7.0e}: Horner steps / frame {work / len(lpcs):6.0f}, solves at the 20-iteration cap "
                  f"{capped:4d}, resonance-count mismatches {mism:3d}, worst |d| {worst:.1e} Hz")


if __name__ == "__main__":
    main()
