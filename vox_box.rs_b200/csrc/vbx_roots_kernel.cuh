// vbx_roots_kernel.cuh — LPC coefficients → roots → resonances, one thread per frame (K3 + K4).
// Included by vbx_formants.cu (launcher) and by the vbx_roots_inst_*.cu translation units that
// instantiate the kernel for ranges of the LPC order (split only to compile in parallel).
#pragma once
#include "vbx_complex.cuh"
#include "vbx_internal.cuh"

namespace vbx_roots {

constexpr double kPi = 3.14159265358979323846264338327950288;

// spectrum.rs:166-192 from_root (fp64).  Returns true and fills (f, bw) if the root yields a resonance.
__device__ __forceinline__ bool from_root_f64(double re, double im, double fs, bool strict_im, double* f_out, double* bw_out) {
    const double freq_mul = fs / (kPi * 2.0);
    if (strict_im ? !(im > 0.0) : !(im >= 0.0)) return false;
    double r = hypot(re, im), theta = atan2(im, re);
    if (r > 1.0) {  // reflect around the unit circle: 1/conj(z)
        const double ns = re * re + im * im;
        const double ire = re / ns, iim = im / ns;  // inv(conj(z)) = conj(conj z)/|z|² = z/|z|²
        r = hypot(ire, iim);
        theta = atan2(iim, ire);
    }
    const double f = freq_mul * theta;
    const double bw = -2.0 * freq_mul * log(r);
    if (f > 50.0 && f < fs * 0.5 - 50.0) {
        *f_out = f;
        *bw_out = bw;
        return true;
    }
    return false;
}

struct RootsParams {
    const void* lpc;       // [F][lpc_stride] LPC coefficients (f64 or f32)
    const uint8_t* status_in;  // per-frame status of the LPC stage (or null)
    void* res_out;         // [F][R] resonance pairs, sorted ascending, zero padded
    int32_t* nres_out;     // [F] or null
    void* roots_out;       // [F][P] complex roots in find_roots order (or null)
    uint8_t* status_out;   // [F] or null
    int64_t n_frames;
    double fs;
    int lpc_stride;        // entries per frame in `lpc`
    int lpc_has_one;       // 1: lpc[f] = [1, a1..ap] (Levinson `ac`), 0: [a1..ap] (Burg)
    int lpc_f64, out_f64;
    int R;                 // resonance slots per frame in res_out
    int strict_im;         // 1: keep roots with im > 0 (lib.rs:95), 0: im >= 0 (to_resonance)
    int polish_steps;
};

// Laguerre + deflation for degrees M = P, P−1, …, 3 (statically unrolled so every index is a register).
template <typename TR, int P, int M, bool FAST> struct SolveAll {
    static __device__ __forceinline__ void run(vcx<TR>* c, vcx<TR>* roots) {
        if constexpr (M >= 3) {
            const vcx<TR> z = laguerre_solve<TR, M, P, FAST>(c, cmk<TR>((TR)-2, (TR)-2));
            roots[P - M] = z;
            deflate<TR, M>(c, z);
            SolveAll<TR, P, M - 1, FAST>::run(c, roots);
        }
    }
};

// One thread per frame.  TR = arithmetic of Laguerre/deflation (float: fast path + fp64 polish;
// double: the reference's own precision).
template <int P, typename TR>
__global__ void __launch_bounds__(128) lpc_roots_kernel(const RootsParams Q) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= Q.n_frames) return;
    const int R = Q.R;
    auto write_res = [&](int slot, double fr_, double bw_) {
        if (Q.out_f64) {
            double* o = reinterpret_cast<double*>(Q.res_out) + ((size_t)f * R + slot) * 2;
            o[0] = fr_; o[1] = bw_;
        } else {
            float* o = reinterpret_cast<float*>(Q.res_out) + ((size_t)f * R + slot) * 2;
            o[0] = (float)fr_; o[1] = (float)bw_;
        }
    };
    if (Q.status_in && Q.status_in[f] != VBX_OK) {  // the LPC stage failed: find_formants returns Err before root finding
        if (Q.status_out) Q.status_out[f] = Q.status_in[f];
        if (Q.nres_out) Q.nres_out[f] = 0;
        if (Q.res_out) for (int s = 0; s < R; ++s) write_res(s, 0.0, 0.0);
        return;
    }
    // polynomial in ascending powers: a[k] = coefficient of z^k = lpc_{P-k}, a[P] = 1   (lib.rs:78-91)
    double a[P + 1];
#pragma unroll
    for (int k = 0; k < P; ++k) {
        const int idx = (P - k) - (Q.lpc_has_one ? 0 : 1);
        a[k] = Q.lpc_f64 ? reinterpret_cast<const double*>(Q.lpc)[(size_t)f * Q.lpc_stride + idx]
                         : (double)reinterpret_cast<const float*>(Q.lpc)[(size_t)f * Q.lpc_stride + idx];
    }
    a[P] = Q.lpc_has_one ? (Q.lpc_f64 ? reinterpret_cast<const double*>(Q.lpc)[(size_t)f * Q.lpc_stride]
                                      : (double)reinterpret_cast<const float*>(Q.lpc)[(size_t)f * Q.lpc_stride])
                         : 1.0;
    vcx<TR> c[P + 1];
#pragma unroll
    for (int k = 0; k <= P; ++k) c[k] = cmk<TR>((TR)a[k], (TR)0);
    vcx<TR> roots[P];
    // polynomial.rs:116-128: for m = P down to 3: z = laguerre(coeffs, −2−2i); deflate
    constexpr bool FAST = (sizeof(TR) == 4);
    SolveAll<TR, P, P, FAST>::run(c, roots);
    if (P >= 2) {
        // polynomial.rs:131-139 quadratic tail: (−c1 ± sqrt(c1² − 4 c2 c0)) / 2c2, "+" first
        const vcx<TR> a2 = cadd(c[2], c[2]);
        const vcx<TR> four_ac = cmul(cmul(cmk<TR>((TR)4, (TR)0), c[2]), c[0]);
        const vcx<TR> d = csqrt_principal(csub(cmul(c[1], c[1]), four_ac));
        const vcx<TR> x = cneg(c[1]);
        roots[P - 2] = cdiv(cadd(x, d), a2);
        roots[P - 1] = cdiv(csub(x, d), a2);
    } else {
        roots[0] = cdiv(cneg(c[0]), c[1]);  // polynomial.rs:141-144 linear tail
    }
    // resonances: fp64 polish of the roots that can become resonances, then from_root
    double rf[P], rb[P];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < P; ++k) {
        vcx<double> z = cmk<double>((double)roots[k].re, (double)roots[k].im);
        const bool cand = Q.strict_im ? (z.im > 0.0) : (z.im >= 0.0);
        if (cand && Q.polish_steps > 0 && FAST) z = newton_polish<P>(a, z, Q.polish_steps);
        if (Q.roots_out) {
            if (Q.out_f64) {
                double* o = reinterpret_cast<double*>(Q.roots_out) + ((size_t)f * P + k) * 2;
                o[0] = z.re; o[1] = z.im;
            } else {
                float* o = reinterpret_cast<float*>(Q.roots_out) + ((size_t)f * P + k) * 2;
                o[0] = (float)z.re; o[1] = (float)z.im;
            }
        }
        double fr_, bw_;
        const bool ok = cand && from_root_f64(z.re, z.im, Q.fs, Q.strict_im != 0, &fr_, &bw_);
        rf[k] = ok ? fr_ : -1.0;  // −1 marks "no resonance"
        rb[k] = ok ? bw_ : 0.0;
        cnt += ok ? 1 : 0;
    }
    if (Q.status_out) Q.status_out[f] = VBX_OK;  // NaN/inf roots yield no resonance, silently, as in the reference
    if (Q.nres_out) Q.nres_out[f] = cnt;
    if (Q.res_out) {
        // stable rank sort by frequency (lib.rs:105-110 / spectrum.rs:207), zero padding behind
#pragma unroll
        for (int k = 0; k < P; ++k) {
            if (rf[k] >= 0.0) {
                int rank = 0;
#pragma unroll
                for (int j = 0; j < P; ++j)
                    rank += (rf[j] >= 0.0 && (rf[j] < rf[k] || (rf[j] == rf[k] && j < k))) ? 1 : 0;
                if (rank < R) write_res(rank, rf[k], rb[k]);
            }
        }
        for (int s = cnt; s < R; ++s) write_res(s, 0.0, 0.0);
    }
}


typedef void (*roots_kernel_t)(const RootsParams);
constexpr int kMaxRootsOrder = 24;
template <typename TR, int LO, int HI> struct RootsFill {
    static void fill(roots_kernel_t* t) {
        t[HI] = lpc_roots_kernel<HI, TR>;
        if constexpr (HI > LO) RootsFill<TR, LO, HI - 1>::fill(t);
    }
};
// defined in vbx_roots_inst_*.cu
void fill_f32_lo(roots_kernel_t* t);
void fill_f32_hi(roots_kernel_t* t);
void fill_f64_lo(roots_kernel_t* t);
void fill_f64_hi(roots_kernel_t* t);

}  // namespace vbx_roots
