"""int16 PCM input: samples are scaled by 1/32767 on load exactly as the reference's WAV
drivers do (tests/lib.rs:17-19: `sample as f64 / (i32::MAX >> 16) as f64`), fused into the window table.
Every entry point that takes a vbx_frames view must give the same result for int16 samples as for the
float samples s/32767 (the fp32 copy of those is not exact, so the comparison is against the f64 oracle on
s/32767 in f64 for LPC, and against the float path with a tolerance elsewhere)."""
import os

import numpy as np
import pytest

from gpu_util import ctx, normwise, synth, vb

pytestmark = pytest.mark.gpu


def _pcm(seconds=1.0, fs=16000, seed=50):
    x = synth.utterance(seed, fs, seconds=seconds)
    return np.round(x / np.max(np.abs(x)) * 30000).astype(np.int16)


def test_pcm16_lpc_matches_oracle_on_scaled_samples(oracle):
    c = ctx()
    pcm = _pcm()
    N, hop, p = 400, 160, 12
    F = c.n_frames_of(pcm.size, N, hop)
    d = c.to_device(pcm)
    r, ac, _ = c.lpc(c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, dtype=vb.I16), p)
    # the oracle's batch loop takes fp32 samples: feed the scaled samples rounded to fp32 and allow that rounding
    xs = (pcm.astype(np.float64) / 32767.0)
    win = oracle.hanning_window(N)
    ref_r = np.stack([oracle.autocorrelate(xs[f * hop: f * hop + N] * win, p + 1) for f in range(F)])
    ref_ac = np.stack([oracle.lpc(rr, p) for rr in ref_r])
    assert normwise(r.to_host(), ref_r).max() < 1e-12
    assert normwise(ac.to_host(), ref_ac).max() < 1e-8
    # unaligned start (odd sample offset) takes the scalar staging path
    d2 = c.to_device(pcm[1:])
    r2, _, _ = c.lpc(c.frames(d2.ptr, F - 1, N, hop, vb.WINDOW_HANN_SYMMETRIC, dtype=vb.I16), p)
    ref2 = np.stack([oracle.autocorrelate(xs[1 + f * hop: 1 + f * hop + N] * win, p + 1) for f in range(F - 1)])
    assert normwise(r2.to_host(), ref2).max() < 1e-12


def test_pcm16_wav_fixture_find_formants(oracle, fixtures_dir):
    """tests/lib.rs:44-90 test_formant_calculation shape on the raw PCM of short_sample.wav: bin 1024 / hop 512,
    order 10, MALE estimates; int16 input vs the f64 oracle on sample/32767."""
    import wave
    with wave.open(os.path.join(fixtures_dir, "short_sample.wav")) as w:
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).copy()
        fs = float(w.getframerate())
    c = ctx()
    N, hop, p = 1024, 512, 10
    F = c.n_frames_of(pcm.size, N, hop)
    est = np.array([[[f, 1.0] for f in (320., 1440., 2760., 3200.)]])
    d = c.to_device(pcm)
    out = c.find_formants(c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_PERIODIC, dtype=vb.I16), fs, p, vb.LPC_BURG, est)
    x = pcm.astype(np.float64) / 32767.0
    state = est[0].copy()
    for f in range(F):
        res = oracle.find_formants(x[f * hop: f * hop + N].copy(), fs, p, state)
        state = res["formants"]
        assert res["status"] == 0
        assert np.max(np.abs(out["tracks"][f] - state)) < 0.5, f
    b12 = np.array([[1030.918, 264.413], [2724.528, 320.901], [3719.483, 114.118], [3200.0, 1.0]])
    assert np.max(np.abs(out["tracks"][0] - b12)) < 5e-3


def test_pcm16_pitch_mfcc_match_float_path():
    c = ctx()
    pcm = _pcm(seconds=1.5, seed=51)
    xf = (pcm.astype(np.float64) / 32767.0).astype(np.float32)
    fs = 16000.0
    d16, d32 = c.to_device(pcm), c.to_device(xf)
    Fp = c.n_frames_of(pcm.size, 640, 160)
    a = c.pitch(c.frames(d16.ptr, Fp, 640, 160, vb.WINDOW_HANN_SYMMETRIC, dtype=vb.I16), fs, 0.45, 75.0, 600.0, 16)
    b = c.pitch(c.frames(d32.ptr, Fp, 640, 160, vb.WINDOW_HANN_SYMMETRIC), fs, 0.45, 75.0, 600.0, 16)
    ca, cb = a["candidates"].to_host(), b["candidates"].to_host()
    assert np.array_equal(a["n_cand"].to_host(), b["n_cand"].to_host())
    assert np.array_equal(ca[:, 0, 0] != 0, cb[:, 0, 0] != 0) and np.max(np.abs(ca[:, 0, 0] - cb[:, 0, 0])) < 0.1
    Fm = c.n_frames_of(pcm.size, 400, 160)
    ma = c.mfcc(c.frames(d16.ptr, Fm, 400, 160, vb.WINDOW_HANN_SYMMETRIC, dtype=vb.I16), 40, 133.0, 6855.0, fs, n_keep=13).to_host()
    mb = c.mfcc(c.frames(d32.ptr, Fm, 400, 160, vb.WINDOW_HANN_SYMMETRIC), 40, 133.0, 6855.0, fs, n_keep=13).to_host()
    assert normwise(ma, mb).max() < 1e-5  # the float copy of s/32767 is rounded to fp32, the int16 path is not


def test_pcm16_host_twin():
    c = ctx()
    pcm = _pcm(seconds=0.5, seed=52)
    F = c.n_frames_of(pcm.size, 400, 160)
    d = c.to_device(pcm)
    r, ac, kc = c.lpc(c.frames(d.ptr, F, 400, 160, vb.WINDOW_HANN_SYMMETRIC, dtype=vb.I16), 12)
    hr, hac, hkc = c.lpc_host(pcm, F, 400, 160, vb.WINDOW_HANN_SYMMETRIC, 12)
    assert np.array_equal(hr, r.to_host()) and np.array_equal(hac, ac.to_host())
