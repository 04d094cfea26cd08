// oracle_capi.cpp — C entry points over vox_box_oracle.hpp for ctypes.
//
// TEST INFRASTRUCTURE ONLY (see the header of vox_box_oracle.hpp): loaded by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs; never by the product path.
//
// Complex arrays are interleaved (re, im).  Resonance / Pitch arrays are
// interleaved pairs (frequency, bandwidth|strength).  All "batch" functions
// run the reference's serial per-frame loop, optionally spread over frames
// with OpenMP (the stand-in for the "rayon frame-parallel wrapper" named by
// BASELINE.json); n_threads <= 1 is the reference as shipped.
#include "vox_box_oracle.hpp"

#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace vbo;
typedef Cx<double> Cd;
typedef Cx<float> Cf;

#define VBO_API extern "C" __attribute__((visibility("default")))

namespace {
enum { WIN_NONE = 0, WIN_HANN_SYMMETRIC = 1, WIN_HANN_PERIODIC = 2 };
std::vector<double> make_window(int kind, size_t n) {
    if (kind == WIN_HANN_SYMMETRIC) return hanning_window(n);
    if (kind == WIN_HANN_PERIODIC) return hanning_periodic(n, n);
    return std::vector<double>();
}
int threads_or_max(int n_threads) {
#ifdef _OPENMP
    return n_threads > 0 ? n_threads : omp_get_max_threads();
#else
    (void)n_threads;
    return 1;
#endif
}
}  // namespace

VBO_API int vbo_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---- crate-semantics helpers -------------------------------------------------
VBO_API void vbo_hanning_window(int64_t n, double* out) { auto w = hanning_window(n); std::memcpy(out, w.data(), n * 8); }
VBO_API void vbo_hanning_lag_window(int64_t n, double* out) { auto w = hanning_lag_window(n); std::memcpy(out, w.data(), n * 8); }
VBO_API void vbo_hanning_periodic(int64_t count, int64_t len, double* out) { auto w = hanning_periodic(count, len); std::memcpy(out, w.data(), count * 8); }
VBO_API void vbo_sine_signal(double fs, double hz, int64_t n, double* out) { auto s = sine_signal(fs, hz, n); std::memcpy(out, s.data(), n * 8); }
VBO_API int64_t vbo_windower_count(int64_t len, int64_t bin, int64_t hop) { return (int64_t)windower_count(len, bin, hop); }

// ---- waves.rs ------------------------------------------------------------------
VBO_API double vbo_rms(const double* x, int64_t n) { return rms(x, n); }
VBO_API double vbo_max_amplitude(const double* x, int64_t n) { return max_amplitude(x, n); }
VBO_API void vbo_normalize(double* x, int64_t n) { normalize(x, n); }
VBO_API void vbo_normalize_with_max(double* x, int64_t n, double max) { normalize_with_max(x, n, true, max); }
VBO_API void vbo_preemphasis(double* x, int64_t n, double factor) { preemphasis(x, n, factor); }

// ---- periodic.rs / spectrum.rs / polynomial.rs scalars -------------------------
VBO_API int vbo_autocorrelate(const double* x, int64_t n, double* r, int64_t n_lags) { return autocorrelate(x, n, r, n_lags); }
VBO_API int vbo_autocorrelate_f32(const float* x, int64_t n, float* r, int64_t n_lags) { return autocorrelate(x, n, r, n_lags); }
VBO_API void vbo_lpc_levinson(const double* r, int64_t p, double* ac, double* kc) {
    std::vector<double> tmp(p);
    lpc_levinson(r, p, ac, kc, tmp.data());
}
VBO_API int vbo_lpc_burg(const double* x, int64_t n, int64_t p, double* coeffs) {
    std::vector<double> work(2 * n + p);
    return lpc_burg(x, n, p, coeffs, work.data());
}
VBO_API int64_t vbo_poly_degree(const double* c, int64_t len) { return poly_degree((const Cd*)c, len); }
VBO_API int64_t vbo_poly_off_low(const double* c, int64_t len) { return poly_off_low((const Cd*)c, len); }
VBO_API void vbo_laguerre(const double* c, int64_t len, double sre, double sim, double* out, int* iters) {
    Cd z = laguerre((const Cd*)c, len, Cd(sre, sim), iters);
    out[0] = z.re; out[1] = z.im;
}
VBO_API void vbo_laguerre_f32(const float* c, int64_t len, float sre, float sim, float* out, int* iters) {
    Cf z = laguerre((const Cf*)c, len, Cf(sre, sim), iters);
    out[0] = z.re; out[1] = z.im;
}
// find_roots_mut: in place over c[0..len); iters (len ints, optional)
VBO_API int vbo_find_roots_mut(double* c, int64_t len, int* iters) { return find_roots_mut((Cd*)c, len, iters); }
VBO_API int vbo_find_roots_mut_f32(float* c, int64_t len, int* iters) { return find_roots_mut((Cf*)c, len, iters); }
// find_roots: allocating form; roots_out has len complex slots; *n_out = count after popping zeros
VBO_API int vbo_find_roots(const double* c, int64_t len, double* roots_out, int64_t* n_out) {
    std::vector<Cd> out;
    int st = find_roots((const Cd*)c, len, out);
    if (st != OK) { *n_out = 0; return st; }
    *n_out = (int64_t)out.size();
    std::memcpy(roots_out, out.data(), out.size() * sizeof(Cd));
    return OK;
}
VBO_API int vbo_find_roots_f32(const float* c, int64_t len, float* roots_out, int64_t* n_out) {
    std::vector<Cf> out;
    int st = find_roots((const Cf*)c, len, out);
    if (st != OK) { *n_out = 0; return st; }
    *n_out = (int64_t)out.size();
    std::memcpy(roots_out, out.data(), out.size() * sizeof(Cf));
    return OK;
}
VBO_API int vbo_div_polynomial(double* self, int64_t len, double ore, double oim, double* rem) {
    return div_polynomial((Cd*)self, len, Cd(ore, oim), (Cd*)rem);
}
VBO_API int vbo_from_root(double re, double im, double fs, double* out) {
    Resonance<double> r;
    if (!from_root(Cd(re, im), fs, &r)) return 0;
    out[0] = r.frequency; out[1] = r.bandwidth;
    return 1;
}
VBO_API int64_t vbo_to_resonance(const double* roots, int64_t n, double fs, double* out) {
    auto res = to_resonance((const Cd*)roots, n, fs);
    std::memcpy(out, res.data(), res.size() * 16);
    return (int64_t)res.size();
}
VBO_API void vbo_estimate_formants(double* est, int64_t n_est, const double* res, int64_t n_res) {
    estimate_formants((Resonance<double>*)est, n_est, (const Resonance<double>*)res, n_res);
}
// FormantExtractor (spectrum.rs:336-369): frames of n_res resonances each; tracks_out [F][n_est]
VBO_API void vbo_formant_extractor(double* est, int64_t n_est, const double* res, int64_t n_frames, int64_t n_res, double* tracks_out) {
    for (int64_t f = 0; f < n_frames; ++f) {
        estimate_formants((Resonance<double>*)est, n_est, (const Resonance<double>*)(res + f * n_res * 2), n_res);
        std::memcpy(tracks_out + f * n_est * 2, est, n_est * 16);
    }
}
VBO_API int64_t vbo_find_formants_real_work_size(int64_t buf_len, int64_t p) { return find_formants_real_work_size(buf_len, p); }
VBO_API int64_t vbo_find_formants_complex_work_size(int64_t p) { return find_formants_complex_work_size(p); }
// find_formants (lib.rs:40).  dbg_* optional: lpc[p], roots[(p+1)*2], resonances[64], n_res
VBO_API int vbo_find_formants(const double* buf, int64_t n, double fs, double ratio, double* resampled_buf, int64_t resampled_buf_len,
                              int64_t p, int64_t work_len, double* formants, int64_t n_formants,
                              double* dbg_lpc, double* dbg_roots, double* dbg_res, int* dbg_nres) {
    FormantDebug<double> dbg;
    int st = find_formants(buf, n, fs, ratio, resampled_buf, resampled_buf_len, p, work_len,
                           (Resonance<double>*)formants, n_formants, &dbg);
    if (st == OK) {
        if (dbg_lpc) std::memcpy(dbg_lpc, dbg.lpc.data(), p * 8);
        if (dbg_roots) std::memcpy(dbg_roots, dbg.roots.data(), (p + 1) * 16);
        if (dbg_res) std::memcpy(dbg_res, dbg.resonances, MAX_RESONANCES * 16);
        if (dbg_nres) *dbg_nres = dbg.n_resonances;
    }
    return st;
}
VBO_API double vbo_interpolate_sinc(const double* y, int64_t y_len, int64_t offset, int64_t nx, double x, int64_t max_depth) {
    return interpolate_sinc(y, y_len, offset, nx, x, max_depth);
}
VBO_API void vbo_improve_extremum(const double* y, int64_t y_len, int64_t offset, int64_t nx, double ixmid, int interp,
                                  int64_t depth, int is_max, double* out_xy, int* n_evals) {
    improve_extremum(y, y_len, offset, nx, ixmid, interp, depth, is_max != 0, &out_xy[0], &out_xy[1], n_evals);
}
// pitch: x windowed frame; cand_out [max_cand][2]; returns status; *n_cand total candidates (may exceed max_cand)
VBO_API int vbo_pitch(const double* x, int64_t n, double fs, double threshold, double fmin, double fmax,
                      double* cand_out, int64_t max_cand, int64_t* n_cand, double* lag_out /*2n or null*/, int* brent_evals) {
    std::vector<Pitch> out;
    PitchDebug dbg;
    int st = pitch(x, n, fs, threshold, fmin, fmax, out, &dbg);
    *n_cand = (int64_t)out.size();
    for (size_t i = 0; i < out.size() && (int64_t)i < max_cand; ++i) { cand_out[2 * i] = out[i].frequency; cand_out[2 * i + 1] = out[i].strength; }
    if (lag_out && st == OK) std::memcpy(lag_out, dbg.lag.data(), 2 * n * 8);
    if (brent_evals) *brent_evals = dbg.brent_evals;
    return st;
}
VBO_API double vbo_hz_to_mel(double hz) { return hz_to_mel(hz); }
VBO_API double vbo_mel_to_hz(double mel) { return mel_to_hz(mel); }
VBO_API void vbo_dct(const double* s, int64_t n, double* out) { dct(s, n, out); }
VBO_API int vbo_mfcc_bins(int64_t n, int64_t num_coeffs, double f_lo, double f_hi, double fs, int64_t* bins) {
    std::vector<size_t> b;
    mfcc_bins(n, num_coeffs, f_lo, f_hi, fs, b);
    for (size_t i = 0; i < b.size(); ++i) bins[i] = (int64_t)b[i];
    return OK;
}
VBO_API int vbo_mfcc(const double* x, int64_t n, int64_t num_coeffs, double f_lo, double f_hi, double fs, double* out,
                     double* energies /*optional*/, int naive_dft) {
    std::vector<double> e;
    int st = mfcc(x, n, num_coeffs, f_lo, f_hi, fs, out, &e, naive_dft != 0);
    if (st == OK && energies) std::memcpy(energies, e.data(), num_coeffs * 8);
    return st;
}
VBO_API void vbo_fft_forward(const double* in, double* out, int64_t n, int naive) {
    if (naive) dft_naive((const Cd*)in, (Cd*)out, n);
    else fft_forward((const Cd*)in, (Cd*)out, n);
}

// ---- batched frame loops (drivers of examples/ and tests/lib.rs) ------------------
// Frames are a strided view over fp32 audio: frame f = base[f*stride .. f*stride+n).
// Samples are widened to f64 (the parity instantiation).

// C2 chain: window → autocorrelate(p+1) → lpc(p).  r_out [F][p+1], ac_out [F][p+1], kc_out [F][p] (any may be null)
VBO_API int vbo_batch_lpc(const float* base, int64_t n_frames, int64_t n, int64_t stride, int window_kind, int64_t p,
                          double* r_out, double* ac_out, double* kc_out, int n_threads) {
    std::vector<double> win = make_window(window_kind, n);
    const double* w = win.empty() ? nullptr : win.data();
    int nt = threads_or_max(n_threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static) if (nt > 1)
    for (int64_t f = 0; f < n_frames; ++f) {
        std::vector<double> r(p + 1), ac(p + 1), kc(p);
        frame_lpc<double>(base + f * stride, n, w, p, r.data(), ac.data(), kc.data());
        if (r_out) std::memcpy(r_out + f * (p + 1), r.data(), (p + 1) * 8);
        if (ac_out) std::memcpy(ac_out + f * (p + 1), ac.data(), (p + 1) * 8);
        if (kc_out) std::memcpy(kc_out + f * p, kc.data(), p * 8);
    }
    return OK;
}
// autocorrelate only: r_out [F][n_lags]
VBO_API int vbo_batch_autocorrelate(const float* base, int64_t n_frames, int64_t n, int64_t stride, int window_kind,
                                    int64_t n_lags, double* r_out, int n_threads) {
    std::vector<double> win = make_window(window_kind, n);
    int nt = threads_or_max(n_threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static) if (nt > 1)
    for (int64_t f = 0; f < n_frames; ++f) {
        std::vector<double> xw(n);
        for (int64_t i = 0; i < n; ++i) xw[i] = win.empty() ? double(base[f * stride + i]) : double(base[f * stride + i]) * win[i];
        autocorrelate(xw.data(), n, r_out + f * n_lags, n_lags);
    }
    return OK;
}
// Burg over frames: coeffs_out [F][p], status_out [F]
VBO_API int vbo_batch_burg(const float* base, int64_t n_frames, int64_t n, int64_t stride, int window_kind, int64_t p,
                           double* coeffs_out, uint8_t* status_out, int n_threads) {
    std::vector<double> win = make_window(window_kind, n);
    int nt = threads_or_max(n_threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(static) if (nt > 1)
    for (int64_t f = 0; f < n_frames; ++f) {
        std::vector<double> xw(n), work(2 * n + p);
        for (int64_t i = 0; i < n; ++i) xw[i] = win.empty() ? double(base[f * stride + i]) : double(base[f * stride + i]) * win[i];
        int st = lpc_burg(xw.data(), n, p, coeffs_out + f * p, work.data());
        if (status_out) status_out[f] = (uint8_t)st;
    }
    return OK;
}

// Formant chain over utterances.  method 0 = Burg (find_formants, lib.rs:40: periodic
// Hann inside), method 1 = "path A" of SURVEY §8d C3 (window_kind → autocorrelate(p+1)
// → lpc(p) → find_roots → from_root(im>0) → sort/pad 32 → estimate_formants).
// utt_frame_offsets [n_utts+1] (frame index ranges); each utterance starts its
// tracker from est_init [n_formants][2].  Outputs (optional): tracks_out
// [F][n_formants][2], res_out [F][32][2], nres_out [F], lpc_out [F][p(+1)], status_out [F].
VBO_API int vbo_batch_formants(const float* base, int64_t n_frames, int64_t n, int64_t stride, int window_kind, int method,
                               double fs, int64_t p, const int64_t* utt_frame_offsets, int64_t n_utts,
                               const double* est_init, int64_t n_formants,
                               double* tracks_out, double* res_out, int32_t* nres_out, double* lpc_out,
                               uint8_t* status_out, int n_threads) {
    std::vector<double> win = (method == 1) ? make_window(window_kind, n) : std::vector<double>();
    int nt = threads_or_max(n_threads);
    (void)nt;
    (void)n_frames;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 1) if (nt > 1)
    for (int64_t u = 0; u < n_utts; ++u) {
        std::vector<Resonance<double>> est(n_formants);
        for (int64_t k = 0; k < n_formants; ++k) est[k] = {est_init[2 * k], est_init[2 * k + 1]};
        std::vector<double> buf(n), rbuf(n);
        for (int64_t f = utt_frame_offsets[u]; f < utt_frame_offsets[u + 1]; ++f) {
            const float* x = base + f * stride;
            int st = OK;
            Resonance<double> resonances[MAX_RESONANCES];
            for (auto& r : resonances) r = {0., 0.};
            int nres = 0;
            if (method == 0) {
                for (int64_t i = 0; i < n; ++i) buf[i] = double(x[i]);
                std::fill(rbuf.begin(), rbuf.end(), 0.0);
                FormantDebug<double> dbg;
                st = find_formants(buf.data(), n, fs, 1.0, rbuf.data(), n, p, find_formants_real_work_size(n, p),
                                   est.data(), n_formants, &dbg);
                if (st == OK) {
                    for (size_t k = 0; k < MAX_RESONANCES; ++k) resonances[k] = dbg.resonances[k];
                    nres = dbg.n_resonances;
                    if (lpc_out) std::memcpy(lpc_out + f * p, dbg.lpc.data(), p * 8);
                }
            } else {
                std::vector<double> r(p + 1), ac(p + 1), kc(p);
                frame_lpc<double>(x, n, win.empty() ? nullptr : win.data(), p, r.data(), ac.data(), kc.data());
                if (lpc_out) std::memcpy(lpc_out + f * (p + 1), ac.data(), (p + 1) * 8);
                // ascending powers: c[k] = ac[p-k]  (same construction as lib.rs:83-90 with ac[0]=1)
                std::vector<Cd> c(p + 1);
                for (int64_t k = 0; k <= p; ++k) c[k] = Cd(ac[p - k]);
                st = find_roots_mut(c.data(), p + 1);
                if (st == OK) {
                    size_t count = 0;
                    for (int64_t k = 0; k <= p; ++k) {
                        Resonance<double> rr;
                        if (c[k].im > 0. && from_root(c[k], fs, &rr) && count < MAX_RESONANCES) resonances[count++] = rr;
                    }
                    size_t rpos = 0;
                    for (size_t k = MAX_RESONANCES; k-- > 0;)
                        if (resonances[k].frequency != 0.) { rpos = k; break; }
                    std::stable_sort(resonances, resonances + rpos + 1,
                                     [](const Resonance<double>& a, const Resonance<double>& b) { return a.frequency < b.frequency; });
                    nres = (int)count;
                    estimate_formants(est.data(), n_formants, resonances, MAX_RESONANCES);
                }
            }
            if (status_out) status_out[f] = (uint8_t)st;
            if (res_out) std::memcpy(res_out + f * MAX_RESONANCES * 2, resonances, MAX_RESONANCES * 16);
            if (nres_out) nres_out[f] = nres;
            if (tracks_out) std::memcpy(tracks_out + f * n_formants * 2, est.data(), n_formants * 16);
        }
    }
    return OK;
}

// Pitch over frames: cand_out [F][max_cand][2] (sorted by strength desc, zero padded), ncand_out [F]
VBO_API int vbo_batch_pitch(const float* base, int64_t n_frames, int64_t n, int64_t stride, int window_kind,
                            double fs, double threshold, double fmin, double fmax,
                            double* cand_out, int64_t max_cand, int32_t* ncand_out, uint8_t* status_out, int n_threads) {
    std::vector<double> win = make_window(window_kind, n);
    int nt = threads_or_max(n_threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 16) if (nt > 1)
    for (int64_t f = 0; f < n_frames; ++f) {
        std::vector<double> xw(n);
        for (int64_t i = 0; i < n; ++i) xw[i] = win.empty() ? double(base[f * stride + i]) : double(base[f * stride + i]) * win[i];
        std::vector<Pitch> out;
        int st = pitch(xw.data(), n, fs, threshold, fmin, fmax, out);
        if (status_out) status_out[f] = (uint8_t)st;
        if (ncand_out) ncand_out[f] = (int32_t)out.size();
        for (int64_t k = 0; k < max_cand; ++k) {
            cand_out[(f * max_cand + k) * 2] = k < (int64_t)out.size() ? out[k].frequency : 0.;
            cand_out[(f * max_cand + k) * 2 + 1] = k < (int64_t)out.size() ? out[k].strength : 0.;
        }
    }
    return OK;
}

// Laguerre statistics since the last reset: out[0] solves, out[1] solves that ran all 20 iterations, out[2] of those the ones
// whose last update was still > 1e-8·max(1, |z|) (not converged).  reset != 0 clears the counters after reading.
VBO_API void vbo_laguerre_stats(int64_t* out, int reset) {
    LaguerreStats& st = laguerre_stats();
    out[0] = st.solves.load();
    out[1] = st.capped.load();
    out[2] = st.unconverged.load();
    if (reset) { st.solves = 0; st.capped = 0; st.unconverged = 0; }
}

// the same loop with a rounding-level variant of the autocorrelation fold (tools/pitch_sensitivity.py)
VBO_API int vbo_batch_pitch_variant(const float* base, int64_t n_frames, int64_t n, int64_t stride, int window_kind,
                                    double fs, double threshold, double fmin, double fmax, int acf_variant,
                                    double* cand_out, int64_t max_cand, int32_t* ncand_out, uint8_t* status_out, int n_threads) {
    std::vector<double> win = make_window(window_kind, n);
    int nt = threads_or_max(n_threads);
    (void)nt;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 16) if (nt > 1)
    for (int64_t f = 0; f < n_frames; ++f) {
        std::vector<double> xw(n);
        for (int64_t i = 0; i < n; ++i) xw[i] = win.empty() ? double(base[f * stride + i]) : double(base[f * stride + i]) * win[i];
        std::vector<Pitch> out;
        int st = pitch(xw.data(), n, fs, threshold, fmin, fmax, out, nullptr, acf_variant);
        if (status_out) status_out[f] = (uint8_t)st;
        if (ncand_out) ncand_out[f] = (int32_t)out.size();
        for (int64_t k = 0; k < max_cand; ++k) {
            cand_out[(f * max_cand + k) * 2] = k < (int64_t)out.size() ? out[k].frequency : 0.;
            cand_out[(f * max_cand + k) * 2 + 1] = k < (int64_t)out.size() ? out[k].strength : 0.;
        }
    }
    return OK;
}

// MFCC over frames: out [F][n_keep] (first n_keep of num_coeffs DCT rows)
VBO_API int vbo_batch_mfcc(const float* base, int64_t n_frames, int64_t n, int64_t stride, int window_kind,
                           int64_t num_coeffs, double f_lo, double f_hi, double fs, int64_t n_keep,
                           double* out, int naive_dft, int n_threads) {
    std::vector<double> win = make_window(window_kind, n);
    int nt = threads_or_max(n_threads);
    (void)nt;
    int rc = OK;
#pragma omp parallel for num_threads(nt) schedule(static) if (nt > 1)
    for (int64_t f = 0; f < n_frames; ++f) {
        std::vector<double> xw(n), c(num_coeffs);
        for (int64_t i = 0; i < n; ++i) xw[i] = win.empty() ? double(base[f * stride + i]) : double(base[f * stride + i]) * win[i];
        int st = mfcc(xw.data(), n, num_coeffs, f_lo, f_hi, fs, c.data(), (std::vector<double>*)nullptr, naive_dft != 0);
        if (st != OK) {
#pragma omp critical
            rc = st;
        }
        for (int64_t k = 0; k < n_keep; ++k) out[f * n_keep + k] = c[k];
    }
    return rc;
}
