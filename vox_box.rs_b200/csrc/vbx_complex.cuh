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
