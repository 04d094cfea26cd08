// vbx_complex.cuh — complex arithmetic in registers, Laguerre iteration, deflation, root polish.
//
// Follows polynomial.rs:34-72 (laguerre), :155-195 (div_polynomial_mut) and the num-complex 0.2
// operator definitions they rely on (naive multiply, (a·conj b)/|b|² division, hypot norm).  The
// principal square root is computed algebraically (same branch as the polar form sqrt(r)·e^{iθ/2},
// θ ∈ (−π, π]: Re >= 0, Im carries the sign of the argument's Im).
#pragma once
#include <cuda_runtime.h>

template <typename T> struct vcx {
    T re, im;
};
template <typename T> __device__ __forceinline__ vcx<T> cmk(T re, T im) { vcx<T> r; r.re = re; r.im = im; return r; }
template <typename T> __device__ __forceinline__ vcx<T> cadd(vcx<T> a, vcx<T> b) { return cmk<T>(a.re + b.re, a.im + b.im); }
template <typename T> __device__ __forceinline__ vcx<T> csub(vcx<T> a, vcx<T> b) { return cmk<T>(a.re - b.re, a.im - b.im); }
template <typename T> __device__ __forceinline__ vcx<T> cneg(vcx<T> a) { return cmk<T>(-a.re, -a.im); }
template <typename T> __device__ __forceinline__ vcx<T> cmul(vcx<T> a, vcx<T> b) {
    return cmk<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
// a·z + c
template <typename T> __device__ __forceinline__ vcx<T> cfma(vcx<T> a, vcx<T> z, vcx<T> c) {
    return cmk<T>(fma(a.re, z.re, fma(-a.im, z.im, c.re)), fma(a.re, z.im, fma(a.im, z.re, c.im)));
}
template <typename T> __device__ __forceinline__ T cnorm_sqr(vcx<T> a) { return a.re * a.re + a.im * a.im; }
__device__ __forceinline__ float cnorm(vcx<float> a) { return hypotf(a.re, a.im); }
__device__ __forceinline__ double cnorm(vcx<double> a) { return hypot(a.re, a.im); }
template <typename T> __device__ __forceinline__ vcx<T> cdiv(vcx<T> a, vcx<T> b) {
    const T ns = cnorm_sqr(b);
    return cmk<T>((a.re * b.re + a.im * b.im) / ns, (a.im * b.re - a.re * b.im) / ns);
}
__device__ __forceinline__ float vsqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double vsqrt(double x) { return sqrt(x); }
// principal square root
template <typename T> __device__ __forceinline__ vcx<T> csqrt_principal(vcx<T> a) {
    const T r = cnorm(a);
    if (r == (T)0) return cmk<T>((T)0, (T)0);
    T re, im;
    if (a.re >= (T)0) {
        re = vsqrt((T)0.5 * (r + a.re));
        im = a.im / ((T)2 * re);
    } else {
        const T t = vsqrt((T)0.5 * (r - a.re));
        re = fabs(a.im) / ((T)2 * t);
        // θ = atan2(im, re) ∈ (−π, π]: im == +0 with re < 0 gives θ = π ⇒ +i·sqrt(|re|)
        im = (a.im < (T)0 || (a.im == (T)0 && signbit(a.im))) ? -t : t;
    }
    return cmk<T>(re, im);
}

// One Laguerre solve on c[0..=M] (degree M, ascending powers) with the reference's fixed `n = NREF`
// (the slice length − 1, polynomial.rs:35 — never the deflated degree) and
// c1 = sqrt((n−1)·n·cb − ca2).  Coefficients above M are zero in the reference's buffer, so starting
// Horner at M is arithmetically identical.  Up to 20 iterations, exit only if |P(z)| <= 1e-16.
// FAST adds a convergence exit (|Δz| tiny relative to |z|) for the fp32 path, whose result is
// polished in fp64 afterwards.
template <typename T, int M, int NREF, bool FAST>
__device__ __forceinline__ vcx<T> laguerre_solve(const vcx<T>* c, vcx<T> z, int* iters_out = nullptr) {
    int it = 0;
    for (; it < 20; ++it) {
        vcx<T> a0 = c[M], a1 = cmk<T>((T)0, (T)0), a2 = cmk<T>((T)0, (T)0);
#pragma unroll
        for (int j = M - 1; j >= 0; --j) {
            a2 = cfma(a2, z, a1);
            a1 = cfma(a1, z, a0);
            a0 = cfma(a0, z, c[j]);
        }
        if (cnorm(a0) <= (T)1.0e-16) break;
        const vcx<T> ca = cdiv(cneg(a1), a0);
        const vcx<T> ca2 = cmul(ca, ca);
        const vcx<T> t2 = cdiv(cmk<T>((T)2 * a2.re, (T)2 * a2.im), a0);
        const vcx<T> cb = csub(ca2, t2);
        const T nn = (T)((NREF - 1) * NREF);
        const vcx<T> c1 = csqrt_principal(cmk<T>(nn * cb.re - ca2.re, nn * cb.im - ca2.im));
        const vcx<T> cc1 = cadd(ca, c1), cc2 = csub(ca, c1);
        const vcx<T> den = (cnorm(cc1) > cnorm(cc2)) ? cc1 : cc2;
        const vcx<T> step = cdiv(cmk<T>((T)NREF, (T)0), den);
        z = cadd(z, step);
        if (FAST) {
            // converged for the purpose of the fp64 polish that follows
            const T eps = (sizeof(T) == 4) ? (T)3.0e-7 : (T)1.0e-15;
            if (cnorm_sqr(step) <= eps * eps * cnorm_sqr(z)) { ++it; break; }
        }
    }
    if (iters_out) *iters_out = it;
    return z;
}

// Runtime-degree variant (generic find_roots path): c has `len` entries, n = len − 1.
template <typename T> __device__ inline vcx<T> laguerre_solve_rt(const vcx<T>* c, int len, vcx<T> z) {
    const int n = len - 1;
    for (int it = 0; it < 20; ++it) {
        vcx<T> a0 = c[n], a1 = cmk<T>((T)0, (T)0), a2 = cmk<T>((T)0, (T)0);
        for (int j = n - 1; j >= 0; --j) {
            a2 = cadd(cmul(a2, z), a1);
            a1 = cadd(cmul(a1, z), a0);
            a0 = cadd(cmul(a0, z), c[j]);
        }
        if (cnorm(a0) <= (T)1.0e-16) break;
        const vcx<T> ca = cdiv(cneg(a1), a0);
        const vcx<T> ca2 = cmul(ca, ca);
        const vcx<T> cb = csub(ca2, cdiv(cmk<T>((T)2 * a2.re, (T)2 * a2.im), a0));
        const T nn = (T)(n - 1) * (T)n;
        const vcx<T> c1 = csqrt_principal(cmk<T>(nn * cb.re - ca2.re, nn * cb.im - ca2.im));
        const vcx<T> cc1 = cadd(ca, c1), cc2 = csub(ca, c1);
        const vcx<T> den = (cnorm(cc1) > cnorm(cc2)) ? cc1 : cc2;
        z = cadd(z, cdiv(cmk<T>((T)n, (T)0), den));
    }
    return z;
}

// Deflation by the root z (polynomial.rs:155-195 with other = −z): q[i] = c[i+1] + z·q[i+1],
// q[M−1] = c[M]; the quotient replaces c[0..M) and c[M] becomes 0.
template <typename T, int M> __device__ __forceinline__ void deflate(vcx<T>* c, vcx<T> z) {
    vcx<T> carry = c[M];
    c[M] = cmk<T>((T)0, (T)0);
#pragma unroll
    for (int i = M - 1; i >= 0; --i) {
        const vcx<T> old = c[i];
        c[i] = carry;
        // rem[i] = rem[i] − self[i]·other, other = −z  ⇒  rem[i] = old + carry·z
        carry = cmk<T>(old.re + (carry.re * z.re - carry.im * z.im), old.im + (carry.re * z.im + carry.im * z.re));
    }
}

// Two fp64 Newton steps on the ORIGINAL real-coefficient polynomial a[0..=P] (ascending powers).
template <int P> __device__ __forceinline__ vcx<double> newton_polish(const double* a, vcx<double> z, int steps) {
    for (int s = 0; s < steps; ++s) {
        vcx<double> p0 = cmk<double>(a[P], 0.0), p1 = cmk<double>(0.0, 0.0);
#pragma unroll
        for (int j = P - 1; j >= 0; --j) {
            p1 = cfma(p1, z, p0);
            p0 = cmk<double>(fma(p0.re, z.re, fma(-p0.im, z.im, a[j])), fma(p0.re, z.im, p0.im * z.re));
        }
        const double ns = cnorm_sqr(p1);
        if (ns == 0.0) break;
        z = csub(z, cdiv(p0, p1));
    }
    return z;
}
