"""ctypes loader for the CPU parity oracle (oracle/_build/libvoxbox_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(vox_box.rs_b200/) never imports this module.

The oracle restates the reference crate's algorithms in f64 (see
oracle/vox_box_oracle.hpp for the file:line citations).
"""
import ctypes as C
import os
import subprocess
import wave

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvoxbox_oracle.so")

WIN_NONE, WIN_HANN_SYMMETRIC, WIN_HANN_PERIODIC = 0, 1, 2
OK, ERR_LPC, ERR_PITCH, ERR_POLYNOMIAL, ERR_WORKSPACE, ERR_BADARG = 0, 1, 2, 3, 4, 6
INTERP_NONE, INTERP_PARABOLIC, INTERP_SINC = 0, 1, 2
MAX_RESONANCES = 32
MALE_FORMANT_ESTIMATES = (320.0, 1440.0, 2760.0, 3200.0)
FEMALE_FORMANT_ESTIMATES = (480.0, 1760.0, 3200.0, 3520.0)


def build(force=False):
    """Compile the oracle with oracle/Makefile (g++, OpenMP)."""
    if force or not os.path.exists(_SO) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
        for f in ("oracle_capi.cpp", "vox_box_oracle.hpp", "Makefile")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_i64 = C.c_int64
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)


def _declare(L):
    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("vbo_max_threads", C.c_int)
    sig("vbo_hanning_window", None, _i64, _dp)
    sig("vbo_hanning_lag_window", None, _i64, _dp)
    sig("vbo_hanning_periodic", None, _i64, _i64, _dp)
    sig("vbo_sine_signal", None, C.c_double, C.c_double, _i64, _dp)
    sig("vbo_windower_count", _i64, _i64, _i64, _i64)
    sig("vbo_rms", C.c_double, _dp, _i64)
    sig("vbo_max_amplitude", C.c_double, _dp, _i64)
    sig("vbo_normalize", None, _dp, _i64)
    sig("vbo_normalize_with_max", None, _dp, _i64, C.c_double)
    sig("vbo_preemphasis", None, _dp, _i64, C.c_double)
    sig("vbo_autocorrelate", C.c_int, _dp, _i64, _dp, _i64)
    sig("vbo_autocorrelate_f32", C.c_int, _fp, _i64, _fp, _i64)
    sig("vbo_lpc_levinson", None, _dp, _i64, _dp, _dp)
    sig("vbo_lpc_burg", C.c_int, _dp, _i64, _i64, _dp)
    sig("vbo_poly_degree", _i64, _dp, _i64)
    sig("vbo_poly_off_low", _i64, _dp, _i64)
    sig("vbo_laguerre", None, _dp, _i64, C.c_double, C.c_double, _dp, _ip)
    sig("vbo_laguerre_f32", None, _fp, _i64, C.c_float, C.c_float, _fp, _ip)
    sig("vbo_find_roots_mut", C.c_int, _dp, _i64, _ip)
    sig("vbo_find_roots_mut_f32", C.c_int, _fp, _i64, _ip)
    sig("vbo_find_roots", C.c_int, _dp, _i64, _dp, _i64p)
    sig("vbo_find_roots_f32", C.c_int, _fp, _i64, _fp, _i64p)
    sig("vbo_div_polynomial", C.c_int, _dp, _i64, C.c_double, C.c_double, _dp)
    sig("vbo_from_root", C.c_int, C.c_double, C.c_double, C.c_double, _dp)
    sig("vbo_to_resonance", _i64, _dp, _i64, C.c_double, _dp)
    sig("vbo_estimate_formants", None, _dp, _i64, _dp, _i64)
    sig("vbo_formant_extractor", None, _dp, _i64, _dp, _i64, _i64, _dp)
    sig("vbo_find_formants_real_work_size", _i64, _i64, _i64)
    sig("vbo_find_formants_complex_work_size", _i64, _i64)
    sig("vbo_find_formants", C.c_int, _dp, _i64, C.c_double, C.c_double, _dp, _i64, _i64, _i64, _dp, _i64,
        _dp, _dp, _dp, _ip)
    sig("vbo_interpolate_sinc", C.c_double, _dp, _i64, _i64, _i64, C.c_double, _i64)
    sig("vbo_improve_extremum", None, _dp, _i64, _i64, _i64, C.c_double, C.c_int, _i64, C.c_int, _dp, _ip)
    sig("vbo_pitch", C.c_int, _dp, _i64, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _i64, _i64p, _dp, _ip)
    sig("vbo_hz_to_mel", C.c_double, C.c_double)
    sig("vbo_mel_to_hz", C.c_double, C.c_double)
    sig("vbo_dct", None, _dp, _i64, _dp)
    sig("vbo_mfcc_bins", C.c_int, _i64, _i64, C.c_double, C.c_double, C.c_double, _i64p)
    sig("vbo_mfcc", C.c_int, _dp, _i64, _i64, C.c_double, C.c_double, C.c_double, _dp, _dp, C.c_int)
    sig("vbo_fft_forward", None, _dp, _dp, _i64, C.c_int)
    sig("vbo_batch_lpc", C.c_int, _fp, _i64, _i64, _i64, C.c_int, _i64, _dp, _dp, _dp, C.c_int)
    sig("vbo_batch_autocorrelate", C.c_int, _fp, _i64, _i64, _i64, C.c_int, _i64, _dp, C.c_int)
    sig("vbo_batch_burg", C.c_int, _fp, _i64, _i64, _i64, C.c_int, _i64, _dp, _u8p, C.c_int)
    sig("vbo_batch_formants", C.c_int, _fp, _i64, _i64, _i64, C.c_int, C.c_int, C.c_double, _i64, _i64p, _i64,
        _dp, _i64, _dp, _dp, _i32p, _dp, _u8p, C.c_int)
    sig("vbo_batch_pitch", C.c_int, _fp, _i64, _i64, _i64, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
        _dp, _i64, _i32p, _u8p, C.c_int)
    sig("vbo_laguerre_stats", None, _i64p, C.c_int)
    sig("vbo_batch_pitch_variant", C.c_int, _fp, _i64, _i64, _i64, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
        _dp, _i64, _i32p, _u8p, C.c_int)
    sig("vbo_batch_mfcc", C.c_int, _fp, _i64, _i64, _i64, C.c_int, _i64, C.c_double, C.c_double, C.c_double, _i64,
        _dp, C.c_int, C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return a.ctypes.data_as(_fp)


def _f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _f32(x):
    a = np.asarray(x)
    assert a.dtype == np.float32 and a.flags.c_contiguous, "audio must be contiguous float32"
    return a


def _cplx(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.complex128))


# ---- crate semantics -------------------------------------------------------------
def max_threads():
    return lib().vbo_max_threads()


def hanning_window(n):
    out = np.empty(n)
    lib().vbo_hanning_window(n, _d(out))
    return out


def hanning_lag_window(n):
    out = np.empty(n)
    lib().vbo_hanning_lag_window(n, _d(out))
    return out


def hanning_periodic(count, length=None):
    out = np.empty(count)
    lib().vbo_hanning_periodic(count, count if length is None else length, _d(out))
    return out


def sine_signal(fs, hz, n):
    out = np.empty(n)
    lib().vbo_sine_signal(fs, hz, n, _d(out))
    return out


def sine(length):
    """The `sine(len)` helper of the reference's unit tests (periodic.rs:470-473)."""
    return sine_signal(float(length), 1.0, length)


def windower_count(length, bin_, hop):
    return lib().vbo_windower_count(length, bin_, hop)


def read_wav(path):
    """PCM16 mono WAV → f64 samples scaled by 1/32767 (tests/lib.rs:17-19). Returns (samples, fs)."""
    with wave.open(path) as w:
        assert w.getnchannels() == 1 and w.getsampwidth() == 2
        raw = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")
        return raw.astype(np.float64) / 32767.0, float(w.getframerate())


# ---- waves.rs ----------------------------------------------------------------------
def rms(x):
    x = _f64(x)
    return lib().vbo_rms(_d(x), x.size)


def max_amplitude(x):
    x = _f64(x)
    return lib().vbo_max_amplitude(_d(x), x.size)


def normalize(x, max_=None):
    x = _f64(x).copy()
    if max_ is None:
        lib().vbo_normalize(_d(x), x.size)
    else:
        lib().vbo_normalize_with_max(_d(x), x.size, max_)
    return x


def preemphasis(x, factor):
    x = _f64(x).copy()
    lib().vbo_preemphasis(_d(x), x.size, factor)
    return x


# ---- periodic.rs / spectrum.rs / polynomial.rs ------------------------------------------
def autocorrelate(x, n_lags):
    x = _f64(x)
    r = np.empty(n_lags)
    st = lib().vbo_autocorrelate(_d(x), x.size, _d(r), n_lags)
    if st:
        raise ValueError("autocorrelate: lag out of range (reference panics)")
    return r


def lpc(r, p, with_kc=False):
    r = _f64(r)
    ac, kc = np.zeros(p + 1), np.zeros(p)
    lib().vbo_lpc_levinson(_d(r), p, _d(ac), _d(kc))
    return (ac, kc) if with_kc else ac


def lpc_praat(x, p):
    x = _f64(x)
    co = np.zeros(p)
    st = lib().vbo_lpc_burg(_d(x), x.size, p, _d(co))
    return st, co


def laguerre(coeffs, start, f32=False):
    it = C.c_int(0)
    if f32:
        c = np.ascontiguousarray(np.asarray(coeffs, dtype=np.complex64))
        out = np.empty(2, dtype=np.float32)
        lib().vbo_laguerre_f32(c.view(np.float32).ctypes.data_as(_fp), c.size, np.float32(start.real),
                               np.float32(start.imag), out.ctypes.data_as(_fp), C.byref(it))
    else:
        c = _cplx(coeffs)
        out = np.empty(2)
        lib().vbo_laguerre(_d(c.view(np.float64)), c.size, start.real, start.imag, _d(out), C.byref(it))
    return complex(out[0], out[1]), it.value


def find_roots(coeffs, f32=False):
    """Allocating find_roots (polynomial.rs:79-89). Returns (status, roots)."""
    n = C.c_int64(0)
    if f32:
        c = np.ascontiguousarray(np.asarray(coeffs, dtype=np.complex64))
        out = np.zeros(c.size, dtype=np.complex64)
        st = lib().vbo_find_roots_f32(c.view(np.float32).ctypes.data_as(_fp), c.size,
                                      out.view(np.float32).ctypes.data_as(_fp), C.byref(n))
    else:
        c = _cplx(coeffs)
        out = np.zeros(c.size, dtype=np.complex128)
        st = lib().vbo_find_roots(_d(c.view(np.float64)), c.size, _d(out.view(np.float64)), C.byref(n))
    return st, out[: n.value]


def find_roots_mut(coeffs):
    """In-place form: returns (status, buffer after write-back, laguerre iteration counts)."""
    c = _cplx(coeffs).copy()
    iters = np.zeros(c.size, dtype=np.int32)
    st = lib().vbo_find_roots_mut(_d(c.view(np.float64)), c.size, iters.ctypes.data_as(_ip))
    return st, c, iters


def div_polynomial(coeffs, other):
    c = _cplx(coeffs).copy()
    rem = np.zeros_like(c)
    st = lib().vbo_div_polynomial(_d(c.view(np.float64)), c.size, other.real, other.imag, _d(rem.view(np.float64)))
    return st, c, rem


def poly_degree(coeffs):
    c = _cplx(coeffs)
    return lib().vbo_poly_degree(_d(c.view(np.float64)), c.size)


def poly_off_low(coeffs):
    c = _cplx(coeffs)
    return lib().vbo_poly_off_low(_d(c.view(np.float64)), c.size)


def from_root(root, fs):
    out = np.empty(2)
    ok = lib().vbo_from_root(root.real, root.imag, fs, _d(out))
    return (out[0], out[1]) if ok else None


def to_resonance(roots, fs):
    r = _cplx(roots)
    out = np.zeros((max(r.size, 1), 2))
    n = lib().vbo_to_resonance(_d(r.view(np.float64)), r.size, fs, _d(out))
    return out[:n]


def estimate_formants(estimates, resonances):
    est = _f64(estimates).reshape(-1, 2).copy()
    res = _f64(resonances).reshape(-1, 2)
    lib().vbo_estimate_formants(_d(est), est.shape[0], _d(res), res.shape[0])
    return est


def formant_extractor(estimates, frames):
    """FormantExtractor over frames [F][n_res][2]. Returns (tracks [F][k][2], final estimates)."""
    est = _f64(estimates).reshape(-1, 2).copy()
    res = _f64(frames)
    F, n_res = res.shape[0], res.shape[1]
    tracks = np.zeros((F, est.shape[0], 2))
    lib().vbo_formant_extractor(_d(est), est.shape[0], _d(res), F, n_res, _d(tracks))
    return tracks, est


def find_formants(buf, fs, p, formants, resample_ratio=1.0, resampled_buf=None, work_len=None):
    """lib.rs:40 find_formants. Returns dict(status, formants, lpc, roots, resonances, n_res)."""
    buf = _f64(buf)
    n = buf.size
    rlen = int(np.ceil(resample_ratio * n))
    if resampled_buf is None:
        resampled_buf = np.zeros(rlen)
    if work_len is None:
        work_len = lib().vbo_find_formants_real_work_size(max(rlen, resampled_buf.size), p)
    fm = _f64(formants).reshape(-1, 2).copy()
    dl, dr, dres = np.zeros(p), np.zeros(p + 1, dtype=np.complex128), np.zeros((MAX_RESONANCES, 2))
    nres = C.c_int(0)
    st = lib().vbo_find_formants(_d(buf), n, fs, resample_ratio, _d(resampled_buf), resampled_buf.size, p, work_len,
                                 _d(fm), fm.shape[0], _d(dl), _d(dr.view(np.float64)), _d(dres), C.byref(nres))
    return dict(status=st, formants=fm, lpc=dl, roots=dr, resonances=dres, n_res=nres.value)


def interpolate_sinc(y, offset, nx, x, max_depth):
    y = _f64(y)
    return lib().vbo_interpolate_sinc(_d(y), y.size, offset, nx, x, max_depth)


def improve_extremum(y, offset, nx, ixmid, interp=INTERP_SINC, depth=1200, is_max=True):
    y = _f64(y)
    out = np.empty(2)
    ev = C.c_int(0)
    lib().vbo_improve_extremum(_d(y), y.size, offset, nx, ixmid, interp, depth, int(is_max), _d(out), C.byref(ev))
    return out[0], out[1], ev.value


def pitch(x, fs, threshold, fmin, fmax, max_cand=256, want_lag=False):
    """Pitched::pitch::<Hanning> on one windowed frame. Returns (status, candidates [k][2], extras)."""
    x = _f64(x)
    cand = np.zeros((max_cand, 2))
    n = C.c_int64(0)
    lag = np.zeros(2 * x.size) if want_lag else None
    ev = C.c_int(0)
    st = lib().vbo_pitch(_d(x), x.size, fs, threshold, fmin, fmax, _d(cand), max_cand, C.byref(n),
                         _d(lag) if want_lag else None, C.byref(ev))
    return st, cand[: min(n.value, max_cand)], dict(lag=lag, brent_evals=ev.value, n_cand=n.value)


def hz_to_mel(hz):
    return lib().vbo_hz_to_mel(hz)


def mel_to_hz(mel):
    return lib().vbo_mel_to_hz(mel)


def dct(x):
    x = _f64(x)
    out = np.empty_like(x)
    lib().vbo_dct(_d(x), x.size, _d(out))
    return out


def mfcc_bins(n, num_coeffs, f_lo, f_hi, fs):
    b = np.zeros(num_coeffs + 2, dtype=np.int64)
    lib().vbo_mfcc_bins(n, num_coeffs, f_lo, f_hi, fs, b.ctypes.data_as(_i64p))
    return b


def mfcc(x, num_coeffs, f_lo, f_hi, fs, naive_dft=False, want_energies=False):
    x = _f64(x)
    out, e = np.zeros(num_coeffs), np.zeros(num_coeffs)
    st = lib().vbo_mfcc(_d(x), x.size, num_coeffs, f_lo, f_hi, fs, _d(out), _d(e), int(naive_dft))
    if st:
        raise ValueError("mfcc: filter-bank bin out of range (reference panics)")
    return (out, e) if want_energies else out


def fft_forward(x, naive=False):
    x = _cplx(x)
    out = np.empty_like(x)
    lib().vbo_fft_forward(_d(x.view(np.float64)), _d(out.view(np.float64)), x.size, int(naive))
    return out


# ---- batched frame loops over strided fp32 audio ------------------------------------------
def n_frames_of(n_samples, frame_len, hop):
    return int(windower_count(n_samples, frame_len, hop))


def batch_lpc(audio, n_frames, frame_len, stride, window, p, n_threads=1, want_kc=False):
    a = _f32(audio)
    r = np.empty((n_frames, p + 1))
    ac = np.empty((n_frames, p + 1))
    kc = np.empty((n_frames, p)) if want_kc else None
    lib().vbo_batch_lpc(_f(a), n_frames, frame_len, stride, window, p, _d(r), _d(ac), _d(kc) if want_kc else None,
                        n_threads)
    return (r, ac, kc) if want_kc else (r, ac)


def batch_autocorrelate(audio, n_frames, frame_len, stride, window, n_lags, n_threads=1):
    a = _f32(audio)
    r = np.empty((n_frames, n_lags))
    lib().vbo_batch_autocorrelate(_f(a), n_frames, frame_len, stride, window, n_lags, _d(r), n_threads)
    return r


def batch_burg(audio, n_frames, frame_len, stride, window, p, n_threads=1):
    a = _f32(audio)
    co = np.zeros((n_frames, p))
    st = np.zeros(n_frames, dtype=np.uint8)
    lib().vbo_batch_burg(_f(a), n_frames, frame_len, stride, window, p, _d(co), st.ctypes.data_as(_u8p), n_threads)
    return co, st


def batch_formants(audio, n_frames, frame_len, stride, window, method, fs, p, utt_frame_offsets, est_init,
                   n_threads=1):
    """method 0: find_formants (Burg, periodic Hann); 1: Hann→autocorr→Levinson path. Returns dict."""
    a = _f32(audio)
    offs = np.ascontiguousarray(np.asarray(utt_frame_offsets, dtype=np.int64))
    est = _f64(est_init).reshape(-1, 2)
    k = est.shape[0]
    tracks = np.zeros((n_frames, k, 2))
    res = np.zeros((n_frames, MAX_RESONANCES, 2))
    nres = np.zeros(n_frames, dtype=np.int32)
    lpc_out = np.zeros((n_frames, p + (1 if method == 1 else 0)))
    st = np.zeros(n_frames, dtype=np.uint8)
    lib().vbo_batch_formants(_f(a), n_frames, frame_len, stride, window, method, fs, p,
                             offs.ctypes.data_as(_i64p), offs.size - 1, _d(est), k, _d(tracks), _d(res),
                             nres.ctypes.data_as(_i32p), _d(lpc_out), st.ctypes.data_as(_u8p), n_threads)
    return dict(tracks=tracks, resonances=res, n_res=nres, lpc=lpc_out, status=st)


def batch_pitch(audio, n_frames, frame_len, stride, window, fs, threshold, fmin, fmax, max_cand=16, n_threads=1):
    a = _f32(audio)
    cand = np.zeros((n_frames, max_cand, 2))
    nc = np.zeros(n_frames, dtype=np.int32)
    st = np.zeros(n_frames, dtype=np.uint8)
    lib().vbo_batch_pitch(_f(a), n_frames, frame_len, stride, window, fs, threshold, fmin, fmax, _d(cand), max_cand,
                          nc.ctypes.data_as(_i32p), st.ctypes.data_as(_u8p), n_threads)
    return cand, nc, st


def laguerre_stats(reset=True):
    """(solves, solves that ran all 20 iterations, of those not converged) since the last reset — polynomial.rs:34-72."""
    out = np.zeros(3, dtype=np.int64)
    lib().vbo_laguerre_stats(out.ctypes.data_as(_i64p), int(reset))
    return tuple(int(v) for v in out)


def batch_pitch_variant(audio, n_frames, frame_len, stride, window, fs, threshold, fmin, fmax, acf_variant, max_cand=16, n_threads=1):
    """batch_pitch with a rounding-level variant of the autocorrelation fold (1: descending i, 2: fused multiply-add):
    sensitivity experiments only (tools/pitch_sensitivity.py), never a parity target."""
    a = _f32(audio)
    cand = np.zeros((n_frames, max_cand, 2))
    nc = np.zeros(n_frames, dtype=np.int32)
    st = np.zeros(n_frames, dtype=np.uint8)
    lib().vbo_batch_pitch_variant(_f(a), n_frames, frame_len, stride, window, fs, threshold, fmin, fmax, acf_variant, _d(cand), max_cand,
                                  nc.ctypes.data_as(_i32p), st.ctypes.data_as(_u8p), n_threads)
    return cand, nc, st


def batch_mfcc(audio, n_frames, frame_len, stride, window, num_coeffs, f_lo, f_hi, fs, n_keep=None, naive_dft=False,
               n_threads=1):
    a = _f32(audio)
    n_keep = num_coeffs if n_keep is None else n_keep
    out = np.zeros((n_frames, n_keep))
    st = lib().vbo_batch_mfcc(_f(a), n_frames, frame_len, stride, window, num_coeffs, f_lo, f_hi, fs, n_keep, _d(out),
                              int(naive_dft), n_threads)
    if st:
        raise ValueError("mfcc: filter-bank bin out of range (reference panics)")
    return out
