"""GPU parity of MFCC (spectrum.rs:371-441) and the waves.rs helpers through the C ABI vs the f64 oracle.

Tolerance (BASELINE.json north_star): MFCC within 1e-5 relative, norm-wise per frame vector
(max|gpu − ref| <= 1e-5 · max|ref|); clamp flips (a band whose sum is ~1 crossing the log10 clamp
under fp32 FFT error) are counted and reported, never masked."""
import numpy as np
import pytest

from gpu_util import ctx, normwise, synth, vb

pytestmark = pytest.mark.gpu
TOL_MFCC = 1e-5


def _three_tone(oracle, n=400, fs=16000.0):  # SURVEY Appendix B14 signal (already windowed)
    i = np.arange(n)
    x = np.sin(2 * np.pi * 440 * i / fs) + 0.5 * np.sin(2 * np.pi * 1230 * i / fs) + 0.25 * np.sin(2 * np.pi * 3100 * i / fs)
    return x * oracle.hanning_window(n)


def test_mel_and_dct_kats(oracle):
    """spectrum.rs:570-577 (hz_to_mel(300) ≈ 401.25), :605-613 test_dct [.2,.3,.4,.3] → [2.4,−.26131,−.28284,.10823]."""
    assert abs(vb.hz_to_mel(300.0) - 401.25) < 1e-2 and abs(vb.mel_to_hz(401.25) - 300.0) < 1e-2
    assert vb.hz_to_mel(1234.5) == oracle.hz_to_mel(1234.5) and vb.mel_to_hz(987.6) == oracle.mel_to_hz(987.6)
    c = ctx()
    got = c.dct(np.array([0.2, 0.3, 0.4, 0.3]))[0]
    assert np.max(np.abs(got - [2.4, -0.26131259, -0.28284271, 0.10823922])) < 1e-7
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 40))
    got = c.dct(x)
    exp = np.stack([oracle.dct(r) for r in x])
    assert np.max(np.abs(got - exp)) < 1e-12
    got32 = c.dct(x.astype(np.float32))
    assert got32.dtype == np.float32 and np.max(np.abs(got32 - exp)) < 1e-4


def test_mfcc_three_tone_appendix_b(oracle):
    """SURVEY B16/B17 cross-check values (restatement) + the oracle on the same frame."""
    c = ctx()
    x = _three_tone(oracle).astype(np.float32)
    d = c.to_device(x)
    fr = c.frames(d.ptr, 1, 400, 400, vb.WINDOW_NONE)
    out, en = c.mfcc(fr, 13, 100.0, 8000.0, 16000.0, want_energies=True)
    ref, ref_e = oracle.mfcc(x.astype(np.float64), 13, 100.0, 8000.0, 16000.0, want_energies=True)
    assert normwise(out.to_host()[0], ref) < TOL_MFCC
    assert np.max(np.abs(en.to_host()[0] - ref_e)) < 1e-5
    b16 = [33.757459, 16.213206, 0.059104, 2.009578, -6.614071, -7.392607, 6.835764, 0.870308, -10.827949, -4.538998,
           -1.100201, 0.266661, 6.357958]
    # B16 was computed from f64 samples; this frame is the fp32-rounded signal, hence the looser bound
    assert np.max(np.abs(out.to_host()[0] - b16)) < 2e-3
    out40 = c.mfcc(fr, 40, 100.0, 7000.0, 16000.0, n_keep=13).to_host()[0]
    ref40 = oracle.mfcc(x.astype(np.float64), 40, 100.0, 7000.0, 16000.0)[:13]
    assert normwise(out40, ref40) < TOL_MFCC


def test_mfcc_not_nan_on_zeros(oracle):
    """spectrum.rs:593-602 test_mfcc_not_nan: mfcc(13,(100,8000),22050) on 512 zeros is finite
    (every band clamps to 1e-10; SURVEY B19: row 0 = 2·13·1e-10)."""
    c = ctx()
    d = c.to_device(np.zeros(512, dtype=np.float32))
    out = c.mfcc(c.frames(d.ptr, 1, 512, 512, vb.WINDOW_NONE), 13, 100.0, 8000.0, 22050.0).to_host()[0]
    assert np.all(np.isfinite(out)) and abs(out[0] - 2.6e-9) < 1e-15
    assert np.max(np.abs(out - oracle.mfcc(np.zeros(512), 13, 100.0, 8000.0, 22050.0))) < 1e-18


@pytest.mark.parametrize("fs,N,hop,M,keep,lo,hi,window", [
    (16000, 400, 160, 40, 13, 133.0, 6855.0, vb.WINDOW_HANN_SYMMETRIC),   # C5 (real-packed 200 = 4·2·5·5)
    (16000, 512, 256, 13, 13, 100.0, 7000.0, vb.WINDOW_HANN_SYMMETRIC),   # power of two
    (22050, 360, 180, 26, 26, 50.0, 10000.0, vb.WINDOW_HANN_SYMMETRIC),   # radix 3 (180 = 4·5·3·3)
    (16000, 405, 135, 20, 12, 100.0, 7000.0, vb.WINDOW_HANN_SYMMETRIC),   # odd N: complex FFT (3^4·5)
    (16000, 398, 160, 20, 20, 100.0, 7000.0, vb.WINDOW_HANN_SYMMETRIC),   # 199 is prime: direct DFT
    (16000, 400, 160, 40, 40, 133.0, 7600.0, vb.WINDOW_NONE),
])
def test_mfcc_synthetic(oracle, fs, N, hop, M, keep, lo, hi, window):
    audio = synth.utterance(31, fs, seconds=2.0)
    # two gain levels: band sums straddle the log10 clamp at 1
    audio = np.concatenate([audio, 0.05 * audio]).astype(np.float32)
    c = ctx()
    F = min(c.n_frames_of(audio.size, N, hop), 300)
    d = c.to_device(audio)
    out, en = c.mfcc(c.frames(d.ptr, F, N, hop, window), M, lo, hi, float(fs), n_keep=keep, want_energies=True)
    ref = oracle.batch_mfcc(audio, F, N, hop, window, M, lo, hi, float(fs), n_keep=keep, n_threads=0)
    ref_e = np.stack([oracle.mfcc(audio[f * hop: f * hop + N].astype(np.float64) *
                                  (oracle.hanning_window(N) if window == vb.WINDOW_HANN_SYMMETRIC else 1.0),
                                  M, lo, hi, float(fs), want_energies=True)[1] for f in range(F)])
    e = en.to_host()
    flips = np.count_nonzero((e == 1e-10) != (ref_e == 1e-10))
    err = normwise(out.to_host(), ref)
    print(f"mfcc N={N} M={M}: max norm-wise err {err.max():.2e}, median {np.median(err):.2e}, clamp flips {flips}/{e.size}")
    assert flips == 0
    assert err.max() < TOL_MFCC
    # the opt-in fp32 transform: reported, bounded loosely (quiet frames sit on the log10 clamp)
    c.mfcc_set_fft_precision(vb.F32)
    try:
        out32, en32 = c.mfcc(c.frames(d.ptr, F, N, hop, window), M, lo, hi, float(fs), n_keep=keep, want_energies=True)
    finally:
        c.mfcc_set_fft_precision(vb.F64)
    err32 = normwise(out32.to_host(), ref)
    flips32 = np.count_nonzero((en32.to_host() == 1e-10) != (ref_e == 1e-10))
    print(f"   fp32 FFT: max {err32.max():.2e}, median {np.median(err32):.2e}, p99 {np.quantile(err32, 0.99):.2e}, flips {flips32}")
    assert np.median(err32) < 1e-6 and err32.max() < 1e-3


def test_mfcc_host_twin_f32_and_bad_bins(oracle):
    c = ctx()
    audio = synth.utterance(32, 16000, seconds=1.0)
    F = c.n_frames_of(audio.size, 400, 160)
    d = c.to_device(audio)
    dev = c.mfcc(c.frames(d.ptr, F, 400, 160, vb.WINDOW_HANN_SYMMETRIC), 40, 133.0, 6855.0, 16000.0, n_keep=13,
                 out_dtype=vb.F32).to_host()
    host = c.mfcc_host(audio, F, 400, 160, vb.WINDOW_HANN_SYMMETRIC, 40, 133.0, 6855.0, 16000.0, n_keep=13, out_dtype=vb.F32)
    assert dev.dtype == np.float32 and np.array_equal(dev, host)
    ref = oracle.batch_mfcc(audio, F, 400, 160, oracle.WIN_HANN_SYMMETRIC, 40, 133.0, 6855.0, 16000.0, n_keep=13, n_threads=0)
    assert normwise(dev, ref).max() < 2e-5  # fp32 output rounding on top
    # a filter bank reaching past the spectrum: the reference panics on the index
    with pytest.raises(vb.VoxBoxError) as ei:
        c.mfcc(c.frames(d.ptr, 1, 400, 160, vb.WINDOW_NONE), 13, 100.0, 20000.0, 16000.0)
    assert ei.value.status == vb.ERR_BADARG


# ------------------------------------------------------------------------------------------ waves.rs
def test_waves_kats(oracle):
    c = ctx()
    # waves.rs:139-144 test_rms: rms(sine(64)) ≈ 0.707
    s = np.sin(2 * np.pi * np.arange(64) / 64)
    assert abs(c.rms(s)[0] - 0.707) < 1e-3
    # SURVEY A.2 worked example
    got = c.preemphasis(np.array([1.0, 0, 0, 0, 1.0]), 0.1)[0]
    exp = [1.1558545456544038, 0.24805021344239853, 0.3947841760435743, 0.6283185307179586, 1.0]
    assert np.max(np.abs(got - exp)) < 1e-15


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 400, 1000])
def test_waves_vs_oracle(oracle, n):
    c = ctx()
    rng = np.random.default_rng(n)
    x = rng.standard_normal((7, n))
    x[3, 0] = -9.0  # the maximum sits at index 0
    assert np.allclose(c.rms(x), [oracle.rms(r) for r in x], rtol=1e-14, atol=0)
    assert np.array_equal(c.max_amplitude(x), [oracle.max_amplitude(r) for r in x])
    assert np.allclose(c.normalize(x), np.stack([oracle.normalize(r) for r in x]), rtol=1e-15, atol=0)
    assert np.allclose(c.normalize(x, np.full(7, 2.5)), np.stack([oracle.normalize(r, 2.5) for r in x]), rtol=1e-15, atol=0)
    got = c.preemphasis(x, 0.05)
    exp = np.stack([oracle.preemphasis(r, 0.05) for r in x])
    assert np.max(np.abs(got - exp)) <= 1e-13 * max(1.0, np.max(np.abs(exp)))
    x32 = x.astype(np.float32)
    assert np.allclose(c.rms(x32), [oracle.rms(r) for r in x32.astype(np.float64)], rtol=1e-6)


def test_waves_nan_semantics(oracle):
    """max_amplitude folds from |x[0]| with `>`: a NaN at index 0 sticks, a NaN elsewhere never wins (waves.rs:45-58)."""
    c = ctx()
    x = np.array([[1.0, np.nan, -3.0, 2.0], [np.nan, 1.0, 5.0, 2.0]])
    got = c.max_amplitude(x)
    assert got[0] == 3.0 and np.isnan(got[1])
    assert got[0] == oracle.max_amplitude(x[0]) and np.isnan(oracle.max_amplitude(x[1]))


@pytest.mark.parametrize("N", [160, 200, 256, 320, 400, 480, 512, 640, 800, 1024])
def test_mfcc_fast_kernels_every_specialised_length(oracle, N, monkeypatch):
    """Every frame length with a specialised warp-per-frame kernel (vbx_mfcc_fast.cuh), fp64 and fp32 transforms, a frame
    count that leaves a ragged last group, against the oracle and against the any-length CTA kernel."""
    fs = 16000
    audio = synth.utterance(60 + N % 7, fs, seconds=1.2)
    c = ctx()
    hop = N // 2
    F = min(c.n_frames_of(audio.size, N, hop), 37)
    d = c.to_device(audio)
    fr = c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC)
    hi = 7000.0 if N >= 256 else 6000.0
    out = c.mfcc(fr, 26, 100.0, hi, float(fs), n_keep=13).to_host()
    ref = oracle.batch_mfcc(audio, F, N, hop, oracle.WIN_HANN_SYMMETRIC, 26, 100.0, hi, float(fs), n_keep=13, n_threads=0)
    assert normwise(out, ref).max() < 1e-9
    monkeypatch.setenv("VBX_MFCC_GENERIC", "1")
    gen = c.mfcc(fr, 26, 100.0, hi, float(fs), n_keep=13).to_host()
    monkeypatch.delenv("VBX_MFCC_GENERIC")
    assert normwise(out, gen).max() < 1e-12
    c.mfcc_set_fft_precision(vb.F32)
    try:
        out32 = c.mfcc(fr, 26, 100.0, hi, float(fs), n_keep=13).to_host()
    finally:
        c.mfcc_set_fft_precision(vb.F64)
    assert np.median(normwise(out32, ref)) < 1e-6


@pytest.mark.parametrize("M,keep,lo,hi,window", [
    (40, 13, 133.0, 6855.0, vb.WINDOW_HANN_SYMMETRIC),   # C5: num_coeffs divisible by 4 → the folded DCT
    (26, 13, 100.0, 7000.0, vb.WINDOW_HANN_PERIODIC),    # plain DCT loop
    (13, 13, 100.0, 8000.0, vb.WINDOW_NONE),             # odd num_coeffs (padded energy row); bins up to the Nyquist slot
    (44, 20, 20.0, 7900.0, vb.WINDOW_HANN_SYMMETRIC),    # empty low-frequency intervals, bins folded above N/2, two rows per twin
])
@pytest.mark.parametrize("dtype", ["f32", "i16"])
def test_mfcc_lane5_kernel(oracle, monkeypatch, M, keep, lo, hi, window, dtype):
    """mfcc_lane5_kernel (400-sample frames: five lanes per frame, twin warps per group of six frames): several utterances as
    segments, a frame count that leaves a ragged last group, energies, f32 and PCM input, f64 and f32 output — against the
    oracle, against mfcc_warp_kernel, and the kernel that ran is checked by name."""
    fs, N, hop, U = 16000, 400, 160, 5
    ns = fs + 48   # even: every frame starts on a sample pair
    audio = np.stack([synth.utterance(300 + u, fs, seconds=ns / fs)[:ns] for u in range(U)])
    c = ctx()
    J = c.n_frames_of(ns, N, hop)
    F = U * J
    assert F % 6 != 0   # the last group of six frames is ragged
    if dtype == "i16":
        pcm = np.round(audio * 32767.0).astype(np.int16)
        d = c.to_device(pcm)
        host = (pcm.astype(np.float64) / 32767.0).astype(np.float32)   # the oracle's input type: 6e-8 away from the PCM path's samples
        fr = c.frames(d.ptr, F, N, hop, window, dtype=vb.I16, frames_per_segment=J, segment_stride=ns)
    else:
        d = c.to_device(audio.astype(np.float32))
        host = audio.astype(np.float32)
        fr = c.frames(d.ptr, F, N, hop, window, frames_per_segment=J, segment_stride=ns)
    c.profile_begin()
    out, en = c.mfcc(fr, M, lo, hi, float(fs), n_keep=keep, want_energies=True)
    assert "mfcc_lane5_kernel" in c.profile_end()
    out, en = out.to_host(), en.to_host()
    monkeypatch.setenv("VBX_MFCC_LANE5", "0")
    c.profile_begin()
    old, old_en = c.mfcc(fr, M, lo, hi, float(fs), n_keep=keep, want_energies=True)
    assert "mfcc_warp_kernel" in c.profile_end()
    monkeypatch.delenv("VBX_MFCC_LANE5")
    assert normwise(out, old.to_host()).max() < 1e-12 and normwise(en, old_en.to_host()).max() < 1e-12
    win = {vb.WINDOW_HANN_SYMMETRIC: oracle.WIN_HANN_SYMMETRIC, vb.WINDOW_HANN_PERIODIC: oracle.WIN_HANN_PERIODIC,
           vb.WINDOW_NONE: oracle.WIN_NONE}[window]
    ref = np.concatenate([oracle.batch_mfcc(host[u], J, N, hop, win, M, lo, hi, float(fs), n_keep=keep, n_threads=0) for u in range(U)])
    assert normwise(out, ref).max() < (1e-9 if dtype == "f32" else TOL_MFCC)
    out32 = c.mfcc(fr, M, lo, hi, float(fs), n_keep=keep, out_dtype=vb.F32).to_host()
    assert out32.dtype == np.float32 and normwise(out32, out).max() < 1e-6


def test_mfcc_lane5_kernel_falls_back_on_odd_offsets(oracle):
    """Frames that do not start on a sample pair (odd hop, odd base offset) cannot use the paired loads: the warp kernel runs."""
    fs, N = 16000, 400
    audio = synth.utterance(77, fs, seconds=1.0)
    c = ctx()
    d = c.to_device(audio)
    for base_off, hop in ((1, 160), (0, 161)):
        F = c.n_frames_of(audio.size - base_off, N, hop)
        c.profile_begin()
        out = c.mfcc(c.frames(d.ptr + 4 * base_off, F, N, hop, vb.WINDOW_HANN_SYMMETRIC), 40, 133.0, 6855.0, float(fs), n_keep=13).to_host()
        assert "mfcc_warp_kernel" in c.profile_end()
        ref = oracle.batch_mfcc(audio[base_off:], F, N, hop, oracle.WIN_HANN_SYMMETRIC, 40, 133.0, 6855.0, float(fs), n_keep=13, n_threads=0)
        assert normwise(out, ref).max() < 1e-9
