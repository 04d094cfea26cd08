// vox_box_oracle.hpp — CPU restatement of vox_box's framewise speech-analysis path.
//
// TEST INFRASTRUCTURE ONLY.  This header is the parity oracle for the CUDA
// product in vox_box.rs_b200/: it may be included, linked or executed only by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs.  Nothing under vox_box.rs_b200/ includes it.
//
// What it is: a from-scratch C++17 re-expression (templates on T = double or
// float) of the algorithms in the reference crate andrewcsmith/vox_box.rs,
// following the reference bug-for-bug.  Each function
// cites the reference file:line it follows.  The parity precision is T=double
// fed with fp32 samples widened to double.
//
// Parity pinning: the Rust reference cannot be compiled in the build container
// (no cargo/rustc), so this oracle is pinned against every known-answer test
// the reference's own test-suite asserts (tests/test_oracle_kat.py lists them
// with their reference file:line).  Third-party crate semantics it restates
// (sample 0.10 window/phase/sine, num-complex 0.2 arithmetic, rustfft 1.0 as a
// plain forward DFT) are described in SURVEY.md Appendix C; the rustfft and
// sample::interpolate::Linear boundaries have no asserting reference test and
// are marked "parity unpinned" where they are used.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

namespace vbo {

constexpr double kPi = 3.14159265358979323846264338327950288;

// ----------------------------------------------------------------------------
// error.rs:6-16  VoxBoxError  -> status codes shared with include/voxbox_b200.h
// ----------------------------------------------------------------------------
enum Status : int {
    OK = 0,
    ERR_LPC = 1,         // VoxBoxError::LPC("Denum was <= 0.0")            spectrum.rs:124
    ERR_PITCH = 2,       // VoxBoxError::Pitch (declared, never raised)
    ERR_POLYNOMIAL = 3,  // VoxBoxError::Polynomial(..)                      polynomial.rs:95,123,192
    ERR_WORKSPACE = 4,   // VoxBoxError::Workspace                           lib.rs:46-48
    ERR_BADARG = 6       // where the reference would panic (assert!/index)
};

// ----------------------------------------------------------------------------
// num-complex 0.2 arithmetic: naive mul, (a·conj b)/|b|²
// division, hypot norm, polar sqrt.
// ----------------------------------------------------------------------------
template <class T>
struct Cx {
    T re, im;
    Cx() : re(0), im(0) {}
    Cx(T r, T i) : re(r), im(i) {}
    explicit Cx(T r) : re(r), im(0) {}
};
template <class T> inline Cx<T> operator+(Cx<T> a, Cx<T> b) { return {a.re + b.re, a.im + b.im}; }
template <class T> inline Cx<T> operator-(Cx<T> a, Cx<T> b) { return {a.re - b.re, a.im - b.im}; }
template <class T> inline Cx<T> operator-(Cx<T> a) { return {-a.re, -a.im}; }
template <class T> inline Cx<T> operator*(Cx<T> a, Cx<T> b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <class T> inline T norm_sqr(Cx<T> a) { return a.re * a.re + a.im * a.im; }
template <class T> inline T norm(Cx<T> a) { return std::hypot(a.re, a.im); }
template <class T> inline Cx<T> operator/(Cx<T> a, Cx<T> b) {
    T ns = norm_sqr(b);
    return {(a.re * b.re + a.im * b.im) / ns, (a.im * b.re - a.re * b.im) / ns};
}
template <class T> inline bool operator==(Cx<T> a, Cx<T> b) { return a.re == b.re && a.im == b.im; }
template <class T> inline bool operator!=(Cx<T> a, Cx<T> b) { return !(a == b); }
template <class T> inline Cx<T> conj(Cx<T> a) { return {a.re, -a.im}; }
template <class T> inline Cx<T> inv(Cx<T> a) {  // conj / norm_sqr
    T ns = norm_sqr(a);
    return {a.re / ns, -a.im / ns};
}
// Complex::sqrt of num-complex 0.2.0: sqrt(r)·e^{iθ/2}, θ = atan2(im, re) ∈ (−π, π]
template <class T> inline Cx<T> csqrt(Cx<T> a) {
    T r = norm(a), th = std::atan2(a.im, a.re);
    T sr = std::sqrt(r), h = th / T(2);
    return {sr * std::cos(h), sr * std::sin(h)};
}

// ----------------------------------------------------------------------------
// sample 0.10 crate semantics
// ----------------------------------------------------------------------------
// signal::Phase::next_phase: returns the current phase, then
// phase = (phase + step) % 1.0  — ACCUMULATED phase (not i*step).
struct Phase {
    double step, next;
    explicit Phase(double s) : step(s), next(0.0) {}
    double next_phase() {
        double p = next;
        next = std::fmod(next + step, 1.0);
        return p;
    }
};
// window::Hanning::at_phase
inline double hanning_at_phase(double phase) { return 0.5 * (1.0 - std::cos(2.0 * kPi * phase)); }
// periodic.rs:236-247  HanningLag::at_phase (autocorrelation of the Hann window)
inline double hanning_lag_at_phase(double phase) {
    double pi_2 = kPi * 2.0;
    double v = phase * pi_2;
    return (1.0 - phase) * (2.0 / 3.0 + (1.0 / 3.0) * std::cos(v)) + (1.0 / pi_2) * std::sin(v);
}
// window::Window::<_, Hanning>::new(len).take(len): phase step 1/(len-1), accumulated.
inline std::vector<double> window_phases(size_t len) {
    std::vector<double> ph(len);
    Phase p(1.0 / (double(len) - 1.0));
    for (size_t i = 0; i < len; ++i) ph[i] = p.next_phase();
    return ph;
}
inline std::vector<double> hanning_window(size_t len) {  // Windower::hanning's window
    std::vector<double> w = window_phases(len);
    for (auto& v : w) v = hanning_at_phase(v);
    return w;
}
inline std::vector<double> hanning_lag_window(size_t len) {  // periodic.rs:400
    std::vector<double> w = window_phases(len);
    for (auto& v : w) v = hanning_lag_at_phase(v);
    return w;
}
// lib.rs:66-70: periodic Hann used by find_formants, phase = idx * (1/len)
inline std::vector<double> hanning_periodic(size_t count, size_t len) {
    std::vector<double> w(count);
    double len_inv = 1.0 / double(len);
    for (size_t i = 0; i < count; ++i) w[i] = hanning_at_phase(double(i) * len_inv);
    return w;
}
// signal::rate(fs).const_hz(f).sine().take(n): sin(2π·phase), phase accumulated by f/fs.
inline std::vector<double> sine_signal(double fs, double hz, size_t n) {
    std::vector<double> s(n);
    Phase p(hz / fs);
    for (size_t i = 0; i < n; ++i) s[i] = std::sin(2.0 * kPi * p.next_phase());
    return s;
}
// window::Windower::{hanning,rectangle}(frames, bin, hop): number of yielded
// frames — a window is produced while bin <= remaining, then advance by hop.
inline size_t windower_count(size_t len, size_t bin, size_t hop) {
    if (bin == 0 || hop == 0 || len < bin) return 0;
    return (len - bin) / hop + 1;
}

// ----------------------------------------------------------------------------
// waves.rs
// ----------------------------------------------------------------------------
// waves.rs:10-23  RMS::rms
template <class T> T rms(const T* x, size_t n) {
    T sum = T(0);
    for (size_t i = 0; i < n; ++i) sum = sum + x[i] * x[i];
    return std::sqrt(sum / T(double(n)));
}
// waves.rs:39-59  MaxAmplitude::max_amplitude (partial_cmp == Greater; NaN never wins)
template <class T> T max_amplitude(const T* x, size_t n) {
    auto amp = [](T v) { return v < T(0) ? v * T(-1.0) : v; };  // waves.rs:29-36
    T acc = amp(x[0]);
    for (size_t i = 1; i < n; ++i) {
        T a = amp(x[i]);
        if (a > acc) acc = a;
    }
    return acc;
}
// waves.rs:61-76  Normalize::normalize_with_max: x *= 1/max (no zero guard)
template <class T> void normalize_with_max(T* x, size_t n, bool has_max, T max) {
    T scale = T(1) / (has_max ? max : max_amplitude(x, n));
    for (size_t i = 0; i < n; ++i) x[i] = x[i] * scale;
}
template <class T> void normalize(T* x, size_t n) { normalize_with_max(x, n, false, T(0)); }
// waves.rs:82-96  Filter::preemphasis: anti-causal additive IIR, a = 2π·factor
template <class T> void preemphasis(T* x, size_t n, double factor) {
    T last = x[n - 1];
    T filter = T(2.0 * kPi * factor);
    for (size_t k = n - 1; k-- > 0;) {
        x[k] = x[k] + last * filter;
        last = x[k];
    }
}

// ----------------------------------------------------------------------------
// periodic.rs:265-289  Autocorrelate::autocorrelate_mut
// r[lag] = x[0] + Σ_{i=1}^{n-lag-1} x[i]·x[i+lag]   (fold seeded with self[0], .skip(1))
// ----------------------------------------------------------------------------
template <class T> int autocorrelate(const T* x, size_t n, T* r, size_t n_lags) {
    if (n == 0) return ERR_BADARG;  // self[0] panics
    for (size_t lag = 0; lag < n_lags; ++lag) {
        if (lag > n) return ERR_BADARG;  // self.len() - lag underflows
        T acc = x[0];
        for (size_t i = 1; i + lag < n; ++i) acc = acc + x[i] * x[i + lag];
        r[lag] = acc;
    }
    return OK;
}

// ----------------------------------------------------------------------------
// spectrum.rs:63-84  LPC::lpc_mut (Levinson–Durbin).  ac[0..=p], kc[0..p), tmp[0..p)
// ----------------------------------------------------------------------------
template <class T> void lpc_levinson(const T* r, size_t p, T* ac, T* kc, T* tmp) {
    T err = r[0];
    ac[0] = T(1);
    for (size_t i = 1; i <= p; ++i) {
        T acc = r[i];
        for (size_t j = 1; j < i; ++j) acc = acc + ac[j] * r[i - j];
        kc[i - 1] = (-acc) / err;
        ac[i] = kc[i - 1];
        for (size_t j = 0; j < p; ++j) tmp[j] = ac[j];
        for (size_t j = 1; j < i; ++j) ac[j] = ac[j] + kc[i - 1] * tmp[i - j];
        err = err * (T(1) - kc[i - 1] * kc[i - 1]);
    }
}
template <class T> std::vector<T> lpc(const T* r, size_t p) {  // spectrum.rs:86-92
    std::vector<T> ac(p + 1, T(0)), kc(p, T(0)), tmp(p, T(0));
    lpc_levinson(r, p, ac.data(), kc.data(), tmp.data());
    return ac;
}

// ----------------------------------------------------------------------------
// spectrum.rs:101-146  LPC::lpc_praat_mut (Burg, Praat NUMburg form).
// coeffs[0..p) out (no leading 1, sign flipped at the end); work >= 2n+p.
// ----------------------------------------------------------------------------
template <class T> int lpc_burg(const T* x, size_t n, size_t p, T* coeffs, T* work) {
    if (n < 2) return ERR_BADARG;  // b2[n-2] index panics
    T* b1 = work;
    T* b2 = work + n;
    T* aa = work + 2 * n;
    b1[0] = x[0];
    b2[n - 2] = x[n - 1];
    for (size_t j = 2; j < n; ++j) {
        b1[j - 1] = x[j - 1];
        b2[j - 2] = x[j - 1];
    }
    for (size_t i = 1; i <= p; ++i) {
        T num = T(0), denum = T(0);
        for (size_t j = 1; j + i < n + 1; ++j) {  // j in 1..(n-i+1)
            num = num + b1[j - 1] * b2[j - 1];
            denum = denum + b1[j - 1] * b1[j - 1] + b2[j - 1] * b2[j - 1];
        }
        if (denum <= T(0)) return ERR_LPC;
        coeffs[i - 1] = T(2.0) * num / denum;
        for (size_t j = 1; j < i; ++j) coeffs[j - 1] = aa[j - 1] - coeffs[i - 1] * aa[i - j - 1];
        if (i < p) {
            for (size_t j = 1; j <= i; ++j) aa[j - 1] = coeffs[j - 1];
            for (size_t j = 1; j + i < n; ++j) {  // j in 1..(n-i)
                b1[j - 1] = b1[j - 1] - aa[i - 1] * b2[j - 1];
                b2[j - 1] = b2[j] - aa[i - 1] * b1[j];
            }
        }
    }
    for (size_t j = 0; j < p; ++j) coeffs[j] = coeffs[j] * T(-1.0);
    return OK;
}

// ----------------------------------------------------------------------------
// polynomial.rs
// ----------------------------------------------------------------------------
// polynomial.rs:26-32 degree / off_low (last / first non-zero, 0 if none)
template <class T> size_t poly_degree(const Cx<T>* c, size_t len) {
    for (size_t i = len; i-- > 0;)
        if (c[i] != Cx<T>()) return i;
    return 0;
}
template <class T> size_t poly_off_low(const Cx<T>* c, size_t len) {
    for (size_t i = 0; i < len; ++i)
        if (c[i] != Cx<T>()) return i;
    return 0;
}
// polynomial.rs:34-72 laguerre.  n = len-1 of the SLICE (never the deflated
// degree); c1 = sqrt((n-1)·n·cb − ca2); ≤20 iterations; exit only if |P| ≤ 1e-16.
// *iters (optional) receives the number of completed update steps.
// Statistics over every solve since the last reset (test infrastructure: tools/parity_scale.py counts the frames on which
// the reference's own 20-iteration cap leaves a NON-root — it then deflates by that non-root, and parity must copy the
// result): solves, solves that ran all 20 iterations, and of those the ones whose last update was still larger than
// 1e-8·max(1, |z|), i.e. not converged.
struct LaguerreStats {
    std::atomic<long long> solves{0}, capped{0}, unconverged{0};
};
inline LaguerreStats& laguerre_stats() {
    static LaguerreStats s;
    return s;
}
template <class T> Cx<T> laguerre(const Cx<T>* c, size_t len, Cx<T> start, int* iters = nullptr) {
    size_t n = len - 1;
    Cx<T> z = start;
    int it = 0;
    double last_step = 0.0;
    for (; it < 20; ++it) {
        Cx<T> abg0 = c[n], abg1, abg2;
        for (size_t j = n; j-- > 0;) {
            abg2 = abg2 * z + abg1;
            abg1 = abg1 * z + abg0;
            abg0 = abg0 * z + c[j];
        }
        if (norm(abg0) <= T(1.0e-16)) break;
        Cx<T> ca = (-abg1) / abg0;
        Cx<T> ca2 = ca * ca;
        Cx<T> cb = ca2 - ((Cx<T>(T(1) + T(1)) * abg2) / abg0);
        Cx<T> c1 = csqrt((Cx<T>(T(double(n - 1))) * Cx<T>(T(double(n))) * cb) - ca2);
        Cx<T> cc1 = ca + c1;
        Cx<T> cc2 = ca - c1;
        Cx<T> cc = (norm(cc1) > norm(cc2)) ? Cx<T>(T(double(n))) / cc1 : Cx<T>(T(double(n))) / cc2;
        z = z + cc;
        last_step = std::sqrt(double(cc.re) * double(cc.re) + double(cc.im) * double(cc.im));
    }
    {
        LaguerreStats& st = laguerre_stats();
        st.solves.fetch_add(1, std::memory_order_relaxed);
        if (it == 20) {
            st.capped.fetch_add(1, std::memory_order_relaxed);
            const double za = std::sqrt(double(z.re) * double(z.re) + double(z.im) * double(z.im));
            if (!(last_step <= 1e-8 * (za > 1.0 ? za : 1.0))) st.unconverged.fetch_add(1, std::memory_order_relaxed);
        }
    }
    if (iters) *iters = it;
    return z;
}
// polynomial.rs:155-195 div_polynomial_mut: synthetic division of self by (x + other).
// Quotient left in self, remainder in rem[0]; literal zeroing semantics kept.
template <class T> int div_polynomial(Cx<T>* self, size_t len, Cx<T> other, Cx<T>* rem) {
    for (size_t i = 0; i < len; ++i) rem[i] = self[i];
    if (other != Cx<T>()) {
        size_t ns = poly_degree(self, len);
        const size_t ds = 1;
        // for i in (0..(ns - ds + 1)).rev(); ns == 0 would underflow (panic) in the reference
        if (ns < ds) return ERR_BADARG;
        for (size_t i = ns - ds + 1; i-- > 0;) {
            self[i] = rem[ds + i];
            rem[i] = rem[i] - (self[i] * other);  // j == i branch only (ds == 1)
        }
        for (size_t k = ds; k < ns + 1; ++k) rem[poly_degree(rem, len)] = Cx<T>();
        size_t l = poly_degree(self, len);
        // for _ in 0..((l + 1) - ns - ds + 1): usize arithmetic, evaluated left to right
        long long cnt = (long long)(l + 1) - (long long)ns - (long long)ds + 1;
        if (cnt < 0) return ERR_BADARG;  // would panic on underflow
        for (long long k = 0; k < cnt; ++k) self[poly_degree(self, len)] = Cx<T>();
        return OK;
    }
    return ERR_POLYNOMIAL;  // "Tried to divide by zero"
}
// polynomial.rs:92-152 find_roots_mut.  Roots are written back into self[0..],
// one extra element copied from the (fresh, zero) work buffer, rest zeroed.
// laguerre_iters (optional, len entries) receives per-solve iteration counts.
template <class T> int find_roots_mut(Cx<T>* self, size_t len, int* laguerre_iters = nullptr) {
    size_t hi = poly_degree(self, len);
    if (hi < 1) return ERR_POLYNOMIAL;  // "Zero degree polynomial: no roots to be found."
    size_t lo = poly_off_low(self, len);
    size_t m = hi - lo;
    std::vector<Cx<T>> z_roots(2 * len);  // fresh work ⇒ zeros
    size_t zi = 0;
    for (size_t i = 0; i < lo; ++i) { z_roots[i] = Cx<T>(); ++zi; }
    size_t clen = hi - lo + 1;
    std::vector<Cx<T>> rem(clen), coeffs(clen);
    for (size_t co = lo; co <= hi; ++co) {
        if (co >= clen) return ERR_BADARG;  // polynomial.rs:110-112 indexes un-shifted ⇒ panics if lo>0
        coeffs[co] = self[co];
    }
    int solve = 0;
    for (size_t k = m; k >= 3; --k) {
        int it = 0;
        Cx<T> z = laguerre(coeffs.data(), clen, Cx<T>(T(-2.0), T(-2.0)), &it);
        if (laguerre_iters) laguerre_iters[solve] = it;
        ++solve;
        z_roots[zi++] = z;
        if (div_polynomial(coeffs.data(), clen, -z, rem.data()) != OK) return ERR_POLYNOMIAL;  // "Failed to find roots"
        m = m - 1;
    }
    if (m == 2) {
        Cx<T> a2 = coeffs[2] + coeffs[2];
        Cx<T> d = csqrt((coeffs[1] * coeffs[1]) - (Cx<T>(T(4)) * coeffs[2] * coeffs[0]));
        Cx<T> x = -coeffs[1];
        z_roots[zi] = (x + d) / a2;
        z_roots[zi + 1] = (x - d) / a2;
        zi += 2;
    }
    if (m == 1) {
        z_roots[zi] = (-coeffs[0]) / coeffs[1];
        zi += 1;
    }
    for (size_t i = 0; i < zi + 1; ++i) {
        if (i >= len) return ERR_BADARG;
        self[i] = z_roots[i];
    }
    for (size_t i = zi + 1; i < len; ++i) self[i] = Cx<T>();
    return OK;
}
// polynomial.rs:79-89 find_roots: allocating wrapper, pops trailing exact zeros.
template <class T> int find_roots(const Cx<T>* c, size_t len, std::vector<Cx<T>>& out) {
    out.assign(c, c + len);
    int st = find_roots_mut(out.data(), len);
    if (st != OK) return st;
    while (!out.empty() && out.back() == Cx<T>()) out.pop_back();
    return OK;
}

// ----------------------------------------------------------------------------
// spectrum.rs:149-210  Resonance, from_root, to_resonance
// ----------------------------------------------------------------------------
template <class T> struct Resonance {
    T frequency, bandwidth;
};
template <class T> inline bool operator==(Resonance<T> a, Resonance<T> b) {
    return a.frequency == b.frequency && a.bandwidth == b.bandwidth;
}
template <class T> bool from_root(Cx<T> root, T fs, Resonance<T>* out) {
    T freq_mul = T(double(fs) / (kPi * 2.0));
    if (root.im >= T(0)) {
        T r = norm(root), theta = std::atan2(root.im, root.re);
        if (r > T(1)) {  // reflect around the unit circle: 1/conj(z)
            Cx<T> n = inv(conj(root));
            r = norm(n);
            theta = std::atan2(n.im, n.re);
        }
        Resonance<T> res{freq_mul * theta, T(-2.) * freq_mul * std::log(r)};
        T safety = T(50.), nyquist = fs * T(0.5);
        if (res.frequency > safety && res.frequency < nyquist - safety) {
            *out = res;
            return true;
        }
    }
    return false;
}
template <class T> std::vector<Resonance<T>> to_resonance(const Cx<T>* roots, size_t n, T fs) {
    std::vector<Resonance<T>> res;
    for (size_t i = 0; i < n; ++i) {
        Resonance<T> r;
        if (from_root(roots[i], fs, &r)) res.push_back(r);
    }
    std::stable_sort(res.begin(), res.end(),
                     [](const Resonance<T>& a, const Resonance<T>& b) { return a.frequency < b.frequency; });
    return res;
}

// ----------------------------------------------------------------------------
// spectrum.rs:216-334  EstimateFormants::estimate_formants (one McCandless step)
// estimates[0..n_est) is state (in/out); resonances[0..n_res), n_res >= 1.
// ----------------------------------------------------------------------------
template <class T>
void estimate_formants(Resonance<T>* est, size_t n_est, const Resonance<T>* res, size_t n_res) {
    constexpr size_t NS = 6;
    struct Slot { bool some; Resonance<T> v; };
    Slot slots[NS];
    for (auto& s : slots) s = {false, {T(0), T(0)}};
    auto diff = [](T a, T b) { return std::fabs(a - b); };
    // Step 2: nearest resonance per estimate (first wins ties, strict <)       :235-245
    for (size_t k = 0; k < n_est && k < NS; ++k) {
        Resonance<T> best = res[0];
        T bd = diff(res[0].frequency, est[k].frequency);
        for (size_t j = 1; j < n_res; ++j) {
            T d = diff(res[j].frequency, est[k].frequency);
            if (d < bd) { best = res[j]; bd = d; }
        }
        slots[k] = {true, best};
    }
    // Step 3: remove duplicates                                                :250-272
    size_t w = 0;
    bool has_unassigned = false;
    for (size_t r = 1; r < NS; ++r) {
        if (!slots[r].some) continue;
        Resonance<T> v = slots[r].v;
        // slots[w].unwrap(): w always points at a Some slot when reached from valid input
        if (slots[w].some && v == slots[w].v) {
            if (diff(v.frequency, est[r].frequency) < diff(v.frequency, est[w].frequency)) {
                slots[w].some = false;
                has_unassigned = true;
                w = r;
            } else {
                slots[r].some = false;
                has_unassigned = true;
            }
        } else {
            w = r;
        }
    }
    // Step 4: place unassigned peaks, resonance index j used as slot index    :274-310
    if (has_unassigned) {
        auto contains = [&](Resonance<T> p) {
            for (auto& s : slots)
                if (s.some && s.v == p) return true;
            return false;
        };
        for (size_t j = 0; j < n_res; ++j) {
            Resonance<T> peak = res[j];
            if (contains(peak)) continue;
            if (j < NS && !slots[j].some) { slots[j] = {true, peak}; continue; }
            if (j > 0 && j < NS) {
                if (!slots[j - 1].some) { std::swap(slots[j], slots[j - 1]); slots[j] = {true, peak}; continue; }
            }
            if (j + 1 < NS && !slots[j + 1].some) { std::swap(slots[j], slots[j + 1]); slots[j] = {true, peak}; continue; }
        }
    }
    // Step 5: stable sort, None first then ascending frequency                 :312-324
    std::stable_sort(slots, slots + NS, [](const Slot& a, const Slot& b) {
        if (!a.some) return b.some;          // None < Some ; None vs None equal
        if (!b.some) return false;           // Some > None
        return a.v.frequency < b.v.frequency;  // partial_cmp, unordered ⇒ Equal
    });
    // winners with f > 0 overwrite the leading estimates                       :327-332
    size_t k = 0;
    for (size_t s = 0; s < NS && k < n_est; ++s) {
        if (slots[s].some && slots[s].v.frequency > T(0)) est[k++] = slots[s].v;
    }
}

// ----------------------------------------------------------------------------
// lib.rs:26-116  find_formants
// ----------------------------------------------------------------------------
constexpr size_t MAX_RESONANCES = 32;
constexpr double MALE_FORMANT_ESTIMATES[4] = {320., 1440., 2760., 3200.};
constexpr double FEMALE_FORMANT_ESTIMATES[4] = {480., 1760., 3200., 3520.};
inline size_t find_formants_real_work_size(size_t buf_len, size_t n_coeffs) { return buf_len * 2 + n_coeffs * 23 + 2; }
inline size_t find_formants_complex_work_size(size_t n_coeffs) { return n_coeffs * 7 + 4; }

// sample::interpolate::{Linear, Converter::scale_sample_hz} as used at lib.rs:57-61.
// PARITY UNPINNED: no reference test exercises resample_ratio != 1;
// restated from the crate's documented behaviour: the
// converter advances the source by 1/ratio per output sample, interpolating
// linearly between the two most recent source frames; an exhausted source
// yields equilibrium (0).
template <class T> void linear_resample(const T* buf, size_t n, double ratio, T* out, size_t out_len) {
    // Linear::new(buf[0], buf[1]); the remaining iterator starts at buf[2].
    auto src = [&](size_t i) -> T { return i < n ? buf[i] : T(0); };
    T left = src(0), right = src(1);
    size_t next_src = 2;
    double interp = 0.0;                 // Converter::interpolation_value
    double step = 1.0 / ratio;           // source_to_target_ratio
    for (size_t k = 0; k < out_len; ++k) {
        // Converter::next: advance whole source frames first, then interpolate
        while (interp >= 1.0) {
            left = right;
            right = src(next_src++);
            interp -= 1.0;
        }
        out[k] = T(double(left) + (double(right) - double(left)) * interp);
        interp += step;
    }
}

// Detailed outputs of one find_formants call (for parity tests of the stages).
template <class T> struct FormantDebug {
    std::vector<T> lpc;                      // Burg coefficients (p values)
    std::vector<Cx<T>> roots;                // complex_lpc after find_roots_mut (p+1 slots)
    Resonance<T> resonances[MAX_RESONANCES];  // sorted, zero padded
    int n_resonances = 0;
};

// find_formants.  buf[0..n) is one frame; resampled_buf has resampled_buf_len
// entries and persists between calls exactly like the caller-owned buffer in
// the reference (tests/lib.rs:66); formants[0..n_formants) is in/out state.
template <class T>
int find_formants(const T* buf, size_t n, T fs, double resample_ratio, T* resampled_buf, size_t resampled_buf_len,
                  size_t p, size_t work_len, Resonance<T>* formants, size_t n_formants,
                  FormantDebug<T>* dbg = nullptr) {
    size_t resampled_len = (size_t)std::ceil(resample_ratio * double(n));
    if (work_len < find_formants_real_work_size(resampled_len, p)) return ERR_WORKSPACE;  // lib.rs:46-48
    if (!(resampled_len <= resampled_buf_len)) return ERR_BADARG;                            // assert! lib.rs:54
    if (resample_ratio != 1.0) {
        size_t cnt = std::min(resampled_buf_len, resampled_len);
        linear_resample(buf, n, resample_ratio, resampled_buf, cnt);
    } else {
        size_t cnt = std::min(resampled_buf_len, n);
        for (size_t i = 0; i < cnt; ++i) resampled_buf[i] = buf[i];
    }
    // periodic Hann over the WHOLE resampled_buf, phase idx/resampled_len (lib.rs:66-70)
    double len_inv = 1.0 / double(resampled_len);
    for (size_t i = 0; i < resampled_buf_len; ++i)
        resampled_buf[i] = resampled_buf[i] * T(hanning_at_phase(double(T(double(i) * len_inv))));
    // lib.rs:72: lpc_work = 2*resampled_buf.len() + p must fit in what is left of work
    if (work_len < p + resampled_buf_len * 2 + p) return ERR_BADARG;  // split_at_mut panics
    std::vector<T> lpc_coeffs(p), lpc_work(resampled_buf_len * 2 + p);
    int st = lpc_burg(resampled_buf, resampled_buf_len, p, lpc_coeffs.data(), lpc_work.data());
    if (st != OK) return st;
    // complex_lpc = rev([1, a1..ap])  (lib.rs:78-91): c[0]=a_p … c[p-1]=a_1, c[p]=1
    std::vector<Cx<T>> clpc(p + 1);
    for (size_t k = 0; k < p; ++k) clpc[k] = Cx<T>(lpc_coeffs[p - 1 - k]);
    clpc[p] = Cx<T>(T(1));
    st = find_roots_mut(clpc.data(), p + 1);
    if (st != OK) return st;
    Resonance<T> resonances[MAX_RESONANCES];
    for (auto& r : resonances) r = {T(0), T(0)};
    size_t count = 0;
    for (size_t k = 0; k < p + 1; ++k) {
        if (clpc[k].im > T(0)) {
            Resonance<T> r;
            if (from_root(clpc[k], fs, &r)) {
                if (count >= MAX_RESONANCES) return ERR_BADARG;
                resonances[count++] = r;
            }
        }
    }
    size_t rpos = 0;
    for (size_t k = MAX_RESONANCES; k-- > 0;)
        if (resonances[k].frequency != T(0)) { rpos = k; break; }
    std::stable_sort(resonances, resonances + rpos + 1,
                     [](const Resonance<T>& a, const Resonance<T>& b) { return a.frequency < b.frequency; });
    if (dbg) {
        dbg->lpc = lpc_coeffs;
        dbg->roots = clpc;
        for (size_t k = 0; k < MAX_RESONANCES; ++k) dbg->resonances[k] = resonances[k];
        dbg->n_resonances = int(count);
    }
    estimate_formants(formants, n_formants, resonances, MAX_RESONANCES);  // lib.rs:114 (all 32, zeros included)
    return OK;
}

// ----------------------------------------------------------------------------
// periodic.rs:29-87  interpolate_sinc (with the swapped-neighbour quirk)
// y has y_len entries; index arithmetic uses wrapping usize adds in the
// reference — restated with signed arithmetic.
// ----------------------------------------------------------------------------
inline double interpolate_sinc(const double* y, size_t y_len, long long offset, size_t nx, double x, size_t max_depth_in) {
    long long max_depth = (long long)max_depth_in;
    // `x.floor() as usize`: negative/NaN saturate to 0
    double fl = std::floor(x);
    long long nl = (fl > 0.0) ? (long long)fl : 0;
    long long nr = nl + 1;
    double phil = x - double(nl);
    double phir = 1. - phil;
    double result = 0.;
    auto at = [&](long long idx) -> double {
        if (idx < 0 || (size_t)idx >= y_len) return std::numeric_limits<double>::quiet_NaN();  // would panic
        return y[idx];
    };
    if (nx < 1) return std::numeric_limits<double>::quiet_NaN();
    if (x > double(nx)) return at(offset + (long long)nx - 1);
    if (x < 0.) return y[0];
    if (std::fabs(x - double(nl)) < 1.0e-10) return at(offset + nl);
    if (std::fabs(x - double(nr)) < 1.0e-10) return at(offset + nr);
    if ((offset + nr) < max_depth) {                 // :46-52
        if ((offset + nr) < 0) max_depth = 0;
        else max_depth = offset + nr;
    }
    if ((offset + nl + max_depth) >= (long long)nx)  // :55-57
        max_depth = (long long)nx - offset + nl - 1;
    for (long long n = 0; n < max_depth + 1; ++n) {
        {   // "left": pairs phil with y[offset+nr-n]
            double a = kPi * (phil + double(n));
            long long lag_val = offset + nr - n;
            if (lag_val < 0) lag_val = 0;
            double r_lag = at(lag_val);
            double first = std::sin(a) / a;
            double second = 0.5 + 0.5 * std::cos(a / (phil + double(max_depth)));
            result += r_lag * first * second;
        }
        {   // "right": pairs phir with y[offset+nl+n], clamped both sides
            double a = kPi * (phir + double(n));
            long long lag_val = offset + nl + n;
            if (lag_val < 0) lag_val = 0;
            if (lag_val >= (long long)y_len) lag_val = (long long)y_len - 1;
            double r_lag = y[lag_val];
            double first = std::sin(a) / a;
            double second = 0.5 + 0.5 * std::cos(a / (phir + double(max_depth)));
            result += r_lag * first * second;
        }
    }
    return result;
}

// ----------------------------------------------------------------------------
// periodic.rs:103-188  brent_maximize (a Brent MINIMISER of f as written)
// ----------------------------------------------------------------------------
template <class F> double brent_maximize(F&& f, double a, double b, double tol, double* fx, int* n_evals = nullptr) {
    const double golden = 1. - 0.6180339887498948482045868343656381177203091798057628621;
    const double EPS = std::numeric_limits<double>::epsilon();
    const double sqrt_epsilon = std::sqrt(EPS);
    const int itermax = 60;
    int evals = 0;
    double v = a + golden * (b - a);
    double fv = f(v); ++evals;
    double x = v, w = v;
    *fx = fv;
    double fw = fv;
    for (int iter = 1; iter <= itermax; ++iter) {
        double range = b - a;
        double middle_range = (a + b) * 0.5;
        double tol_act = sqrt_epsilon * std::fabs(x) + tol / 3.;
        if (std::fabs(x - middle_range) + range * 0.5 <= 2. * tol_act) {
            if (n_evals) *n_evals = evals;
            return x;
        }
        double new_step = (x < middle_range) ? golden * (b - x) : golden * (a - x);
        if (std::fabs(x - w) >= tol_act) {
            double t = (x - w) * (*fx - fv);
            double q = (x - v) * (*fx - fw);
            double p = (x - v) * q - (x - w) * t;
            q = 2. * q - t;
            if (q > 0.) p = -p; else q = -q;
            if (std::fabs(p) < std::fabs(new_step * q) && p > q * (a - x + 2. * tol_act) && p < q * (b - x - 2. * tol_act))
                new_step = p / q;
        }
        if (std::fabs(new_step) < tol_act) new_step = (new_step > 0.) ? tol_act : -tol_act;
        {
            double t = x + new_step;
            double ft = f(t); ++evals;
            if (ft <= *fx) {
                if (t < x) b = x; else a = x;
                v = w; w = x; x = t;
                fv = fw; fw = *fx; *fx = ft;
            } else {
                if (t < x) a = t; else b = t;
                if (ft <= fw || std::fabs(w - x) < EPS) {
                    v = w; w = t;
                    fv = fw; fw = ft;
                } else if (ft <= fv || std::fabs(v - x) < EPS || std::fabs(v - w) < EPS) {
                    v = t;
                    fv = ft;
                }
            }
        }
    }
    if (n_evals) *n_evals = evals;
    return x;
}

// periodic.rs:89-93,192-230  improve_extremum
enum Interp { INTERP_NONE = 0, INTERP_PARABOLIC = 1, INTERP_SINC = 2 };
inline void improve_extremum(const double* y, size_t y_len, long long offset, size_t nx, double ixmid, int interp,
                             size_t sinc_depth, bool is_max, double* xmid, double* ymid, int* n_evals = nullptr) {
    if (n_evals) *n_evals = 0;
    if (ixmid == 0.) { *xmid = 0.; *ymid = y[0]; return; }
    if (ixmid >= double(nx)) { *xmid = double(nx); *ymid = y[nx - 1]; return; }
    switch (interp) {
    case INTERP_NONE: *xmid = 0.; *ymid = y[0]; return;
    case INTERP_PARABOLIC: {
        size_t k = (size_t)std::floor(ixmid);
        double d = y[k + 1] - y[k - 1];
        double mid = y[k];
        double dy = 0.5 * d;
        double d2y = 2.0 * mid - d;
        *xmid = ixmid + dy / d2y;
        *ymid = mid + 0.5 * dy * dy / d2y;
        return;
    }
    default: {
        auto f = [&](double x) {
            double out = interpolate_sinc(y, y_len, offset, nx, x, sinc_depth);
            return is_max ? out : -out;
        };
        double result = 0.;
        *xmid = brent_maximize(f, ixmid - 1., ixmid + 1., 1e-10, &result, n_evals);
        *ymid = result;
    }
    }
}

// ----------------------------------------------------------------------------
// periodic.rs:306-318,362-456  Pitch, LocalMaxima, Pitched::pitch::<Hanning>
// ----------------------------------------------------------------------------
struct Pitch {
    double frequency, strength;
};
struct PitchDebug {
    std::vector<double> lag;       // normalised, window-divided, zero-extended r (2N)
    std::vector<int> maxima;       // indices of local maxima in [0, ixmax)
    std::vector<Pitch> first_pass; // parabolic freq + sinc-30 strength, before the range filter
    int brent_evals = 0;
};
// x: already-windowed frame (f64 values); returns candidates sorted by strength
// descending, always containing {0, threshold}.  local_peak/global_peak are
// ignored by the reference (periodic.rs:396) and therefore not parameters here.
// Rounding-level variants of the reference's autocorrelation fold, for the SENSITIVITY experiments of
// tools/pitch_sensitivity.py only (never a parity target): the same mathematical sums with a different rounding,
// i.e. what any re-implementation that does not reproduce the reference's operation order bit for bit produces.
//   1: terms added in descending i;   2: fused multiply-add (one rounding per term instead of two).
inline void autocorrelate_rounding_variant(const double* x, size_t n, double* r, size_t n_lags, int variant) {
    for (size_t lag = 0; lag < n_lags; ++lag) {
        double acc = 0.0;
        if (variant == 1) {
            for (size_t i = n - lag; i-- > 1;) acc = acc + x[i] * x[i + lag];
            acc = acc + x[0];
        } else {
            acc = x[0];
            for (size_t i = 1; i + lag < n; ++i) acc = std::fma(x[i], x[i + lag], acc);
        }
        r[lag] = acc;
    }
}

inline int pitch(const double* x, size_t n, double fs, double threshold, double fmin, double fmax,
                 std::vector<Pitch>& out, PitchDebug* dbg = nullptr, int acf_variant = 0) {
    out.clear();
    if (n < 2) return ERR_BADARG;
    std::vector<double> window_lag = hanning_lag_window(n);           // :400
    std::vector<double> self_lag(n);
    if (acf_variant == 0) autocorrelate(x, n, self_lag.data(), n);    // :403
    else autocorrelate_rounding_variant(x, n, self_lag.data(), n, acf_variant);
    normalize(self_lag.data(), n);                                    // :404
    for (size_t i = 0; i < n; ++i) self_lag[i] = self_lag[i] / window_lag[i];  // :406-408
    self_lag.resize(n * 2, 0.0);                                      // :411
    size_t ixmax = (size_t)std::floor(0.5 * double(n));               // :413-414
    long long offset = -(long long)ixmax - 1;
    size_t nx = (size_t)((long long)ixmax - offset);
    std::vector<Pitch> maxima;
    // local maxima of self_lag[0..ixmax): windows(3), strict both sides   :370-374,417
    for (size_t c = 1; c + 1 < ixmax; ++c) {
        if (!(self_lag[c - 1] < self_lag[c] && self_lag[c + 1] < self_lag[c])) continue;
        double peak = self_lag[c], peak_rev = self_lag[c - 1], peak_fwd = self_lag[c + 1];
        double dr = 0.5 * (peak_fwd - peak_rev);
        double d2r = 2. * peak - (peak_rev - peak_fwd);                // sign quirk :424
        double freq = fs / (double(c) + dr / d2r);
        double nn = fs / freq - double(offset);
        double strn = interpolate_sinc(self_lag.data(), self_lag.size(), offset, nx, nn, 30);
        if (strn > 1.) strn = 1. / strn;
        if (dbg) { dbg->maxima.push_back((int)c); dbg->first_pass.push_back({freq, strn}); }
        if (!((freq == 0.) || (freq > fmin && freq < fmax))) continue;  // :439
        double nref = fs / freq - double(offset);
        double xmid, ymid;
        int ev = 0;
        improve_extremum(self_lag.data(), self_lag.size(), offset, nx, nref, INTERP_SINC, 1200, true, &xmid, &ymid, &ev);
        if (dbg) dbg->brent_evals += ev;
        xmid += double(offset);
        if (ymid > 1.) ymid = 1. / ymid;
        maxima.push_back({fs / xmid, ymid});
    }
    maxima.push_back({0., threshold});                                // :452
    for (auto& p : maxima)
        if (std::isnan(p.strength)) { out = maxima; return ERR_PITCH; }  // partial_cmp().unwrap() panics
    std::stable_sort(maxima.begin(), maxima.end(), [](const Pitch& a, const Pitch& b) { return b.strength < a.strength; });
    if (dbg) dbg->lag = self_lag;
    out = maxima;
    return OK;
}

// ----------------------------------------------------------------------------
// spectrum.rs:375-441  hz_to_mel, mel_to_hz, dct, MFCC::mfcc
// ----------------------------------------------------------------------------
inline double hz_to_mel(double hz) { return 1125. * std::log1p(hz / 700.); }
inline double mel_to_hz(double mel) { return 700. * (std::exp(mel / 1125.) - 1.); }
// spectrum.rs:391-398 dct_mut: direct DCT-II ×2, f64 accumulation
template <class T> void dct(const T* signal, size_t n, T* coeffs) {
    for (size_t k = 0; k < n; ++k) {
        double acc = 0.;
        for (size_t m = 0; m < n; ++m)
            acc = acc + double(signal[m]) * std::cos(kPi * double(k) * (2. * double(m) + 1.) / (2. * double(n)));
        coeffs[k] = T(2. * acc);
    }
}

// rustfft 1.0 FFT::new(len,false).process — restated as the transform it
// computes (unnormalised forward DFT).  PARITY UNPINNED at rounding level: no
// reference test asserts an FFT/MFCC value.  Mixed-radix
// decimation-in-time with exact-angle twiddles; prime factors > 5 fall back to
// an O(p²) butterfly.
namespace detail {
template <class T> void fft_rec(const Cx<T>* in, size_t stride, Cx<T>* out, size_t n, size_t N, const Cx<T>* tw) {
    if (n == 1) { out[0] = in[0]; return; }
    size_t radix = n;
    for (size_t r : {4, 2, 3, 5})
        if (n % r == 0) { radix = r; break; }
    if (radix == n && n > 5) {
        for (size_t r = 7; r * r <= n; r += 2)
            if (n % r == 0) { radix = r; break; }
    }
    size_t m = n / radix;
    for (size_t q = 0; q < radix; ++q) fft_rec(in + q * stride, stride * radix, out + q * m, m, N, tw);
    std::vector<Cx<T>> tmp(radix);
    size_t twstep = N / n;
    for (size_t k = 0; k < m; ++k) {
        for (size_t q = 0; q < radix; ++q) tmp[q] = out[q * m + k] * tw[(q * k * twstep) % N];
        for (size_t s = 0; s < radix; ++s) {
            Cx<T> acc = tmp[0];
            for (size_t q = 1; q < radix; ++q) acc = acc + tmp[q] * tw[((q * s * m) % n) * twstep];
            out[s * m + k] = acc;
        }
    }
}
}  // namespace detail
template <class T> void fft_forward(const Cx<T>* in, Cx<T>* out, size_t n) {
    std::vector<Cx<T>> tw(n);
    for (size_t k = 0; k < n; ++k) {
        double ang = -2.0 * kPi * double(k) / double(n);
        tw[k] = Cx<T>(T(std::cos(ang)), T(std::sin(ang)));
    }
    detail::fft_rec(in, 1, out, n, n, tw.data());
}
// naive O(N²) DFT with long-double accumulation: cross-check for fft_forward
template <class T> void dft_naive(const Cx<T>* in, Cx<T>* out, size_t n) {
    for (size_t k = 0; k < n; ++k) {
        long double re = 0, im = 0;
        for (size_t j = 0; j < n; ++j) {
            long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)((k * j) % n) / (long double)n;
            long double c = std::cos(ang), s = std::sin(ang);
            re += (long double)in[j].re * c - (long double)in[j].im * s;
            im += (long double)in[j].re * s + (long double)in[j].im * c;
        }
        out[k] = Cx<T>(T(re), T(im));
    }
}
// spectrum.rs:411-414 filter-bank bin edges: num_coeffs+2 points
inline int mfcc_bins(size_t n, size_t num_coeffs, double f_lo, double f_hi, double fs, std::vector<size_t>& bins) {
    double mel_range = hz_to_mel(f_hi) - hz_to_mel(f_lo);
    bins.resize(num_coeffs + 2);
    for (size_t i = 0; i < num_coeffs + 2; ++i) {
        double point = (double(i) / double(num_coeffs)) * mel_range + hz_to_mel(f_lo);
        double b = std::floor(double(n + 1) * mel_to_hz(point) / fs);
        bins[i] = (b > 0.0) ? (size_t)b : 0;
    }
    return OK;
}
// spectrum.rs:410-440 mfcc.  x: windowed frame; out: num_coeffs values.
template <class T>
int mfcc(const T* x, size_t n, size_t num_coeffs, double f_lo, double f_hi, double fs, T* out,
         std::vector<T>* energies_out = nullptr, bool use_naive_dft = false) {
    std::vector<size_t> bins;
    mfcc_bins(n, num_coeffs, f_lo, f_hi, fs, bins);
    std::vector<Cx<T>> sig(n), spec(n);
    for (size_t i = 0; i < n; ++i) sig[i] = Cx<T>(x[i]);
    if (use_naive_dft) dft_naive(sig.data(), spec.data(), n);
    else fft_forward(sig.data(), spec.data(), n);
    std::vector<T> energies(num_coeffs);
    for (size_t wdx = 0; wdx < num_coeffs; ++wdx) {
        size_t b0 = bins[wdx], b1 = bins[wdx + 1], b2 = bins[wdx + 2];
        if (b1 < b0 || b2 < b1) return ERR_BADARG;   // usize subtraction underflow panics
        size_t up = b1 - b0;
        double up_sum = 0.;
        for (size_t bin = b0, i = 0; bin < b1; ++bin, ++i) {
            if (bin >= n) return ERR_BADARG;          // index panic
            double multiplier = double(i) / double(up);
            up_sum = up_sum + std::fabs(double(norm_sqr(spec[bin]))) * multiplier;
        }
        size_t down = b2 - b1;
        double down_sum = 0.;
        for (size_t bin = b1, i = 0; bin < b2; ++bin, ++i) {
            if (bin >= n) return ERR_BADARG;
            double multiplier = double(i) / double(down);
            down_sum = down_sum + std::fabs(double(norm(spec[bin]))) * multiplier;   // |X| with a RISING weight (quirk)
        }
        double e = std::log10(up_sum + down_sum);
        // f64::max(1e-10): NaN ⇒ 1e-10
        e = (e > 1.0e-10) ? e : 1.0e-10;
        energies[wdx] = T(e);
    }
    if (energies_out) *energies_out = energies;
    dct(energies.data(), num_coeffs, out);
    return OK;
}

// ----------------------------------------------------------------------------
// Frame-chain helpers used by tests and the CPU baseline: the caller-side
// loops of the reference drivers (examples/pitch_detection.rs:23-30,
// tests/lib.rs:71-83, spectrum.rs:471-487) expressed per frame.
// ----------------------------------------------------------------------------
// Windower::hanning frame → autocorrelate(p+1) → lpc(p)   (north-star C2 chain)
template <class T>
void frame_lpc(const float* frame, size_t n, const double* window /*n or null*/, size_t p, T* r /*p+1*/, T* ac /*p+1*/,
               T* kc /*p*/) {
    std::vector<T> xw(n), tmp(p);
    for (size_t i = 0; i < n; ++i) xw[i] = window ? T(double(frame[i]) * window[i]) : T(frame[i]);
    autocorrelate(xw.data(), n, r, p + 1);
    lpc_levinson(r, p, ac, kc, tmp.data());
}

}  // namespace vbo
