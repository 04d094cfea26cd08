"""The `_host` entry points run a chunked H2D / kernels / D2H pipeline (csrc/vbx_pipeline.cuh).  These tests force
many small chunks (VBX_HOST_CHUNK_MB=0 ⇒ one utterance — or one frame — per chunk) and check that every host twin
returns exactly what the device-pointer entry point returns on the same data, for two-level (utterance) views and for
single-segment views (where the McCandless tracker state has to carry across chunk boundaries)."""
import numpy as np
import pytest

from gpu_util import ctx, synth, vb

pytestmark = pytest.mark.gpu


def _male(n):
    return np.tile(np.array([[f, 1.0] for f in (320., 1440., 2760., 3200.)]), (n, 1, 1))


@pytest.mark.parametrize("chunk_mb", ["0", None])
def test_host_twins_match_device_segmented(monkeypatch, chunk_mb):
    if chunk_mb is not None:
        monkeypatch.setenv("VBX_HOST_CHUNK_MB", chunk_mb)
    c = ctx()
    fs, N, hop, p = 16000, 400, 160, 12
    audio = synth.corpus(5, fs, seconds=0.5, first=40)
    U, ns = audio.shape
    J = c.n_frames_of(ns, N, hop)
    F = U * J
    d = c.to_device(audio)
    fr = c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
    # LPC
    r, ac, kc = c.lpc(fr, p)
    hfr = c.frames(audio.ctypes.data, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, frames_per_segment=J, segment_stride=ns)
    import ctypes as C
    hr, hac, hkc = np.zeros((F, p + 1)), np.zeros((F, p + 1)), np.zeros((F, p))
    c._check(c.lib.vbx_lpc_host(c.h, C.byref(hfr), p, hr.ctypes.data, hac.ctypes.data, hkc.ctypes.data, vb.F64), "vbx_lpc_host")
    assert np.array_equal(hr, r.to_host()) and np.array_equal(hac, ac.to_host()) and np.array_equal(hkc, kc.to_host())
    # formants (both LPC methods), tracker state per utterance
    for method, win in ((vb.LPC_BURG, vb.WINDOW_HANN_PERIODIC), (vb.LPC_AUTOCORR, vb.WINDOW_HANN_SYMMETRIC)):
        frw = c.frames(d.ptr, F, N, hop, win, frames_per_segment=J, segment_stride=ns)
        dev = c.find_formants(frw, float(fs), p, method, _male(U))
        host = c.find_formants_host(audio, F, N, hop, win, float(fs), p, method, _male(U), frames_per_segment=J, segment_stride=ns)
        for k in ("tracks", "estimates", "resonances", "n_res", "status"):
            assert np.array_equal(dev[k], host[k]), (method, k)
    # pitch
    frp = c.frames(d.ptr, U * c.n_frames_of(ns, 640, 160), 640, 160, vb.WINDOW_HANN_SYMMETRIC,
                   frames_per_segment=c.n_frames_of(ns, 640, 160), segment_stride=ns)
    dev = c.pitch(frp, float(fs), 0.45, 75.0, 600.0, 12)
    host = c.pitch_host(audio, frp.n_frames, 640, 160, vb.WINDOW_HANN_SYMMETRIC, float(fs), 0.45, 75.0, 600.0, 12,
                        frames_per_segment=frp.frames_per_segment, segment_stride=ns)
    assert np.array_equal(dev["candidates"].to_host(), host["candidates"]) and np.array_equal(dev["n_cand"].to_host(), host["n_cand"])
    assert np.array_equal(dev["status"].to_host(), host["status"])
    # MFCC
    dev = c.mfcc(fr, 40, 133.0, 6855.0, float(fs), n_keep=13).to_host()
    host = c.mfcc_host(audio, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, 40, 133.0, 6855.0, float(fs), n_keep=13,
                       frames_per_segment=J, segment_stride=ns)
    assert np.array_equal(dev, host)


def test_host_twins_single_segment_frame_chunks(monkeypatch):
    """One long utterance split into per-frame chunks: overlapping frames are re-uploaded per chunk and the
    tracker state carries from chunk to chunk."""
    monkeypatch.setenv("VBX_HOST_CHUNK_MB", "0")
    c = ctx()
    fs, N, hop, p = 16000, 400, 160, 10
    audio = synth.utterance(44, fs, seconds=0.6)
    F = c.n_frames_of(audio.size, N, hop)
    d = c.to_device(audio)
    dev = c.find_formants(c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_PERIODIC), float(fs), p, vb.LPC_BURG, _male(1))
    host = c.find_formants_host(audio, F, N, hop, vb.WINDOW_HANN_PERIODIC, float(fs), p, vb.LPC_BURG, _male(1))
    for k in ("tracks", "estimates", "resonances", "n_res", "status"):
        assert np.array_equal(dev[k], host[k]), k
    r = c.autocorrelate(c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC), 13).to_host()
    assert np.array_equal(r, c.autocorrelate_host(audio, F, N, hop, vb.WINDOW_HANN_SYMMETRIC, 13))
