"""Large batches of the C3 / C4 / C5 shapes (hundreds of utterances tiled from a few distinct ones, so that the oracle only
has to run on the distinct ones): size-independent properties that must hold for every frame — determinism across CTAs,
waves and work-list orders (identical inputs → identical outputs, bit for bit), the invariants of each output — plus
oracle parity on the distinct utterances.  The LPC chain's full-size test (C2, 1 h of audio) is in test_gpu_lpc.py."""
import numpy as np
import pytest

from gpu_util import ctx, normwise, synth, vb

pytestmark = pytest.mark.gpu

MALE = np.array([[f, 1.0] for f in (320., 1440., 2760., 3200.)])


def _tiled(n_distinct, reps, fs, seconds, first):
    base = synth.corpus(n_distinct, fs, seconds, first=first)
    return base, np.ascontiguousarray(np.tile(base, (reps, 1)))


@pytest.mark.parametrize("method,window", [(vb.LPC_AUTOCORR, vb.WINDOW_HANN_SYMMETRIC), (vb.LPC_BURG, vb.WINDOW_HANN_PERIODIC)])
def test_formants_large_batch(oracle, method, window):
    """C3 shape (44.1 kHz, N=1102, hop=441, order 12), 240 utterances of 5 s."""
    fs, N, hop, p = 44100, 1102, 441, 12
    base, audio = _tiled(4, 60, fs, 5.0, first=300)
    U, ns = audio.shape
    c = ctx()
    J = c.n_frames_of(ns, N, hop)
    d = c.to_device(audio)
    fr = c.frames(d.ptr, U * J, N, hop, window, frames_per_segment=J, segment_stride=ns)
    out = c.find_formants(fr, float(fs), p, method, np.tile(MALE, (U, 1, 1)))
    trk, nres, res, status = out["tracks"], out["n_res"], out["resonances"], out["status"]
    assert trk.shape[0] == U * J and np.all(status == 0)
    # determinism: utterance u and u + 4 are the same audio
    for name in ("tracks", "n_res", "resonances", "estimates"):
        a = out[name].reshape((U // 4, 4) + out[name].shape[1:]) if name == "estimates" else out[name].reshape((U // 4, 4 * J) + out[name].shape[1:])
        assert np.array_equal(a, np.broadcast_to(a[:1], a.shape)), name
    # invariants: resonance counts <= p/2, kept frequencies inside (50, fs/2 - 50) and ascending, zero padding after them
    assert np.all((nres >= 0) & (nres <= p // 2))
    f_res = res[..., 0]
    k = np.arange(f_res.shape[1])[None, :]
    live = k < nres[:, None]
    assert np.all(f_res[live] > 50.0) and np.all(f_res[live] < fs / 2 - 50.0)
    assert np.all(np.diff(f_res, axis=1)[live[:, 1:]] >= 0)  # a live slot's left neighbour is live too
    assert np.all(res[~live] == 0)
    assert np.all(np.isfinite(trk))
    # oracle parity on the distinct utterances
    for u in range(4):
        ref = oracle.batch_formants(base[u], J, N, hop, window, 0 if method == vb.LPC_BURG else 1, float(fs), p, [0, J], MALE,
                                    n_threads=0)
        sl = slice(u * J, (u + 1) * J)
        assert np.array_equal(nres[sl], ref["n_res"])
        assert np.max(np.abs(trk[sl] - ref["tracks"])) < 0.5
        assert np.max(np.abs(res[sl] - ref["resonances"])) < 0.5


def test_pitch_large_batch(oracle):
    """C4 shape (16 kHz, N=640, hop=160, 75-600 Hz), 96 utterances of 10 s: 95 712 frames, ~2.4 M candidates through the
    tile-sorted, queue-fed refinement (whose grouping must not change any result)."""
    fs, N, hop, K, thr = 16000, 640, 160, 40, 0.45
    base, audio = _tiled(3, 32, fs, 10.0, first=400)
    U, ns = audio.shape
    c = ctx()
    J = c.n_frames_of(ns, N, hop)
    d = c.to_device(audio)
    fr = c.frames(d.ptr, U * J, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
    res = c.pitch(fr, float(fs), thr, 75.0, 600.0, K, out_dtype=vb.F64)
    cand, ncand, status = res["candidates"].to_host(), res["n_cand"].to_host(), res["status"].to_host()
    assert np.all(status == 0) and np.all(ncand >= 1)
    for a in (cand, ncand):
        b = a.reshape((U // 3, 3 * J) + a.shape[1:])
        assert np.array_equal(b, np.broadcast_to(b[:1], b.shape))
    # invariants: strengths descending over the reported candidates, exactly one unvoiced candidate (0 Hz, threshold),
    # every other frequency inside (75, 600), zero padding after the list
    kk = np.arange(K)[None, :]
    live = kk < np.minimum(ncand, K)[:, None]
    assert np.all(np.diff(cand[..., 1], axis=1)[live[:, 1:]] <= 0)
    unv = live & (cand[..., 0] == 0.0)
    assert np.all(unv.sum(axis=1)[ncand <= K] == 1) and np.all(cand[..., 1][unv] == thr)
    voiced = live & ~unv
    assert np.all(cand[..., 0][voiced] > 75.0) and np.all(cand[..., 0][voiced] < 600.0)
    assert np.all(cand[~live] == 0)
    for u in range(3):
        rc, rn, rs = oracle.batch_pitch(base[u], J, N, hop, oracle.WIN_HANN_SYMMETRIC, float(fs), thr, 75.0, 600.0, K, n_threads=0)
        sl = slice(u * J, (u + 1) * J)
        assert np.array_equal(ncand[sl], rn)
        assert np.array_equal(cand[sl, 0, 0] == 0, rc[:, 0, 0] == 0)          # voiced / unvoiced
        assert np.max(np.abs(cand[sl, 0, 0] - rc[:, 0, 0])) < 0.1               # top candidate


def test_mfcc_large_batch(oracle):
    """C5 shape (16 kHz, N=400, hop=160, 40 bands, 13 kept), 360 utterances of 10 s = 359 280 frames."""
    fs, N, hop = 16000, 400, 160
    base, audio = _tiled(6, 60, fs, 10.0, first=500)
    U, ns = audio.shape
    c = ctx()
    J = c.n_frames_of(ns, N, hop)
    d = c.to_device(audio)
    fr = c.frames(d.ptr, U * J, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
    out, en = c.mfcc(fr, 40, 133.0, 6855.0, float(fs), n_keep=13, want_energies=True, out_dtype=vb.F64)
    m, e = out.to_host(), en.to_host()
    assert m.shape == (U * J, 13) and e.shape == (U * J, 40)
    assert np.all(np.isfinite(m)) and np.all(e >= 1e-10)           # the log10 clamp (spectrum.rs:435)
    b = m.reshape(U // 6, 6 * J, 13)
    assert np.array_equal(b, np.broadcast_to(b[:1], b.shape))
    # DCT row 0 is twice the sum of the band energies (cos 0 = 1, spectrum.rs:395-397)
    assert np.max(np.abs(m[:, 0] - 2.0 * e.sum(axis=1)) / np.maximum(np.abs(m[:, 0]), 1.0)) < 1e-12
    for u in range(6):
        ref = oracle.batch_mfcc(base[u], J, N, hop, oracle.WIN_HANN_SYMMETRIC, 40, 133.0, 6855.0, float(fs), n_keep=13, n_threads=0)
        assert np.max(normwise(m[u * J:(u + 1) * J], ref)) < 1e-5
