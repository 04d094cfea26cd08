"""Non-finite input samples: the reference has no guards on the hot path ("NaN/inf propagate silently"), so
the CUDA path must neither hang nor invent errors — it propagates them the way the oracle does."""
import numpy as np
import pytest

from gpu_util import ctx, synth, vb

pytestmark = pytest.mark.gpu


def _poisoned(fs=16000):
    x = synth.utterance(90, fs, seconds=0.5).copy()
    x[1000] = np.nan      # frames touching sample 1000 are poisoned, the others are clean
    x[5000] = np.inf
    return x


def test_nan_inputs_lpc_formants(oracle):
    c = ctx()
    x = _poisoned()
    N, hop, p = 400, 160, 12
    F = c.n_frames_of(x.size, N, hop)
    d = c.to_device(x)
    r, ac, _ = c.lpc(c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC), p)
    rr, ra = oracle.batch_lpc(x, F, N, hop, oracle.WIN_HANN_SYMMETRIC, p)
    bad = ~np.isfinite(rr).all(axis=1)
    assert bad.any() and (~bad).any()
    assert np.array_equal(~np.isfinite(r.to_host()).all(axis=1), bad)
    assert np.allclose(ac.to_host()[~bad], ra[~bad], rtol=1e-7, atol=1e-9)
    est = np.array([[[f, 1.0] for f in (320., 1440., 2760., 3200.)]])
    for method, win in ((vb.LPC_BURG, vb.WINDOW_HANN_PERIODIC), (vb.LPC_AUTOCORR, vb.WINDOW_HANN_SYMMETRIC)):
        out = c.find_formants(c.frames(d.ptr, F, N, hop, win), 16000.0, p, method, est)
        o = oracle.batch_formants(x, F, N, hop, win, 0 if method == vb.LPC_BURG else 1, 16000.0, p, np.array([0, F]), est[0])
        assert np.array_equal(out["status"], o["status"]) and np.array_equal(out["n_res"], o["n_res"])
        assert np.allclose(out["tracks"], o["tracks"], rtol=0, atol=0.5)


def test_nan_inputs_pitch_mfcc(oracle):
    c = ctx()
    x = _poisoned()
    d = c.to_device(x)
    Fp = c.n_frames_of(x.size, 640, 160)
    res = c.pitch(c.frames(d.ptr, Fp, 640, 160, vb.WINDOW_HANN_SYMMETRIC), 16000.0, 0.45, 75.0, 600.0, 16)
    rc, rn, rs = oracle.batch_pitch(x, Fp, 640, 160, oracle.WIN_HANN_SYMMETRIC, 16000.0, 0.45, 75.0, 600.0, 16)
    assert np.array_equal(res["n_cand"].to_host(), rn) and np.array_equal(res["status"].to_host(), rs)
    assert np.allclose(res["candidates"].to_host()[:, 0, :], rc[:, 0, :], rtol=0, atol=0.1, equal_nan=True)
    Fm = c.n_frames_of(x.size, 400, 160)
    m = c.mfcc(c.frames(d.ptr, Fm, 400, 160, vb.WINDOW_HANN_SYMMETRIC), 40, 133.0, 6855.0, 16000.0, n_keep=13).to_host()
    mo = oracle.batch_mfcc(x, Fm, 400, 160, oracle.WIN_HANN_SYMMETRIC, 40, 133.0, 6855.0, 16000.0, n_keep=13)
    assert np.all(np.isfinite(m)) == np.all(np.isfinite(mo))
    fin = np.isfinite(mo).all(axis=1)
    assert np.allclose(m[fin], mo[fin], rtol=1e-7, atol=1e-7) and np.array_equal(np.isfinite(m).all(axis=1), fin)
