"""GPU parity on REAL speech at 44.1 kHz: examples/formant_extraction/sample-two_vowels.wav of the reference (124 928 samples,
16-bit mono, MIT data; tests/fixtures/), every path through the C ABI vs the f64 oracle on the same samples.

Tolerances (BASELINE.json north_star): autocorrelation / LPC / MFCC 1e-5 relative (norm-wise per frame), formant frequencies and
bandwidths 0.5 Hz with identical resonance counts and track assignments, pitch 0.1 Hz on EVERY list position with identical
voiced/unvoiced decisions."""
import os

import numpy as np
import pytest

from gpu_util import ctx, normwise, vb

pytestmark = pytest.mark.gpu
MALE = np.array([[f, 1.0] for f in (320., 1440., 2760., 3200.)])


@pytest.fixture(scope="module")
def speech(oracle, fixtures_dir):
    x, fs = oracle.read_wav(os.path.join(fixtures_dir, "sample-two_vowels.wav"))
    assert fs == 44100.0 and x.size == 124928
    return x.astype(np.float32), fs


@pytest.mark.parametrize("method,window", [(vb.LPC_BURG, vb.WINDOW_HANN_PERIODIC), (vb.LPC_AUTOCORR, vb.WINDOW_HANN_SYMMETRIC)])
def test_real_speech_formants(oracle, speech, method, window):
    """25 ms / 10 ms frames (N = 1102, hop = 441, the C3 framing), order 12, tracker from the MALE estimates."""
    x, fs = speech
    c = ctx()
    N, hop = 1102, 441
    F = c.n_frames_of(x.size, N, hop)
    d = c.to_device(x)
    out = c.find_formants(c.frames(d.ptr, F, N, hop, window), fs, 12, method, MALE[None])
    ref = oracle.batch_formants(x, F, N, hop, window, 0 if method == vb.LPC_BURG else 1, fs, 12, np.array([0, F]), MALE)
    assert np.array_equal(out["status"], ref["status"])
    assert np.array_equal(out["n_res"], ref["n_res"]), f"resonance-count mismatches in {np.count_nonzero(out['n_res'] != ref['n_res'])} frames"
    assert np.max(np.abs(out["resonances"] - ref["resonances"])) < 0.5
    assert np.max(np.abs(out["tracks"] - ref["tracks"])) < 0.5  # identical slot assignment: a different resonance is >> 0.5 Hz away
    assert np.max(np.abs(out["estimates"][0] - ref["tracks"][-1])) < 0.5


def test_real_speech_formants_10khz_example(oracle, speech):
    """examples/formant_extraction/src/main.rs:40-60: find_formants with resample_ratio = 10000 / 44100 on 1024-sample frames."""
    x, fs = speech
    c = ctx()
    N, hop, ratio = 1024, 512, 10000.0 / 44100.0
    F = min(c.n_frames_of(x.size, N, hop), 120)
    d = c.to_device(x)
    out = c.find_formants_resampled(c.frames(d.ptr, F, N, hop, vb.WINDOW_NONE), 10000.0, ratio, 10, MALE[None])
    est = MALE.copy()
    for f in range(F):
        o = oracle.find_formants(x[f * hop:f * hop + N].astype(np.float64), 10000.0, 10, est, resample_ratio=ratio)
        assert o["status"] == out["status"][f]
        if o["status"] == 0:
            est = o["formants"]
        assert np.max(np.abs(out["tracks"][f] - est)) < 0.5, f


def test_real_speech_pitch(oracle, speech):
    """Boersma candidates, 75-600 Hz, three periods of the floor per frame (N = 1764 at 44.1 kHz), 10 ms hop."""
    x, fs = speech
    c = ctx()
    N, hop, K = 1764, 441, 48
    F = c.n_frames_of(x.size, N, hop)
    d = c.to_device(x)
    res = c.pitch(c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC), fs, 0.45, 75.0, 600.0, K)
    cand, n, st = res["candidates"].to_host(), res["n_cand"].to_host(), res["status"].to_host()
    rc, rn, rs = oracle.batch_pitch(x, F, N, hop, oracle.WIN_HANN_SYMMETRIC, fs, 0.45, 75.0, 600.0, K, n_threads=0)
    assert np.array_equal(st, rs) and np.array_equal(n, rn)
    assert np.array_equal(cand[:, 0, 0] != 0, rc[:, 0, 0] != 0)
    assert np.count_nonzero(cand[:, 0, 0] != 0) > F // 4  # the vowels are voiced
    mask = np.arange(K)[None, :] < np.minimum(n, K)[:, None]
    df = np.abs(cand[..., 0] - rc[..., 0])[mask]
    ds = np.abs(cand[..., 1] - rc[..., 1])[mask]
    assert df.max() < 0.1, f"{np.count_nonzero(df > 0.1)} list positions off by > 0.1 Hz (max {df.max():.3e})"
    assert ds.max() < 1e-6


def test_real_speech_mfcc_and_lpc(oracle, speech):
    x, fs = speech
    c = ctx()
    N, hop = 1102, 441
    F = c.n_frames_of(x.size, N, hop)
    d = c.to_device(x)
    fr = c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC)
    m = c.mfcc(fr, 40, 133.0, 6855.0, fs, n_keep=13).to_host()
    mo = oracle.batch_mfcc(x, F, N, hop, oracle.WIN_HANN_SYMMETRIC, 40, 133.0, 6855.0, fs, n_keep=13)
    assert np.max(normwise(m, mo)) < 1e-5
    r, ac, _ = c.lpc(fr, 12)
    rr, ra = oracle.batch_lpc(x, F, N, hop, oracle.WIN_HANN_SYMMETRIC, 12)
    assert np.max(normwise(r.to_host(), rr)) < 1e-5 and np.max(normwise(ac.to_host(), ra)) < 1e-5
    # the same file as 16-bit PCM (what the reference's WAV readers hold before scaling by 1 / 32767 IN f64, tests/lib.rs:17-19):
    # the oracle gets the f64 samples k / 32767 (the fp32 array above rounds them, which the LPC conditioning turns into 1e-5)
    pcm = np.round(x.astype(np.float64) * 32767.0).astype(np.int16)
    dp = c.to_device(pcm)
    r2, ac2, _ = c.lpc(c.frames(dp.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, dtype=vb.I16), 12)
    x64 = pcm.astype(np.float64) / 32767.0
    w = oracle.hanning_window(N)
    r64 = np.stack([oracle.autocorrelate(x64[f * hop:f * hop + N] * w, 13) for f in range(F)])
    a64 = np.stack([oracle.lpc(r64[f], 12) for f in range(F)])
    assert np.max(normwise(r2.to_host(), r64)) < 1e-12 and np.max(normwise(ac2.to_host(), a64)) < 1e-7
