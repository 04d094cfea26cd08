"""Committed golden fixtures (tests/golden/):
  reference_kats.json — known answers asserted by the reference's own tests (transcribed data);
  oracle_vectors.npz  — oracle outputs on a seeded input (tests/golden/make_golden.py).
CPU tests: the oracle reproduces both (pins the oracle, detects drift).  GPU tests (-m gpu): the CUDA path
through the C ABI reproduces both, with no oracle in the loop."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import make_golden  # noqa: E402

KATS = json.load(open(os.path.join(GOLDEN, "reference_kats.json")))
VEC = np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))
CASES = make_golden.CASES


# ---------------------------------------------------------------------------------------------- CPU: oracle vs golden
def test_oracle_regenerates_golden_vectors(oracle):
    fresh = make_golden.build()
    assert set(fresh) == set(VEC.files)
    for k in VEC.files:
        assert np.array_equal(np.asarray(fresh[k]), VEC[k], equal_nan=True), k


def test_oracle_vs_reference_kats(oracle):
    k = KATS["lpc_burg"]
    assert np.max(np.abs(oracle.lpc_praat(np.array(k["input"], dtype=float), k["order"])[1] - k["coeffs"])) < k["tol"]
    k = KATS["lpc_levinson"]
    auto = oracle.normalize(oracle.autocorrelate(oracle.sine(8), 8))
    assert np.max(np.abs(auto - k["autocorr"])) < k["tol"] and np.max(np.abs(oracle.lpc(auto, k["order"]) - k["lpc"])) < k["tol"]
    k = KATS["roots_cubic"]
    assert np.max(np.abs(oracle.find_roots(np.array(k["coeffs"], dtype=complex))[1] - k["roots"])) < k["tol"]
    k = KATS["dct"]
    assert np.max(np.abs(oracle.dct(k["input"]) - k["output"])) < k["tol"]
    k = KATS["formant_extractor"]
    est = np.array([[f, 1.0] for f in k["estimates"]])
    frames = np.array([[[f, 1.0] for f in fr] for fr in k["frames"]])
    assert np.array_equal(oracle.formant_extractor(est, frames)[0][:, :, 0], np.array(k["tracks"]))


# ---------------------------------------------------------------------------------------------- GPU: CUDA path vs golden
def _gpu():
    from gpu_util import ctx, vb
    return ctx(), vb


@pytest.mark.gpu
def test_gpu_vs_reference_kats():
    c, vb = _gpu()
    k = KATS["lpc_burg"]
    d = c.to_device(np.array(k["input"], dtype=np.float32))
    co, st = c.lpc_burg(c.frames(d.ptr, 1, len(k["input"]), len(k["input"]), vb.WINDOW_NONE), k["order"])
    assert st.to_host()[0] == 0 and np.max(np.abs(co.to_host()[0] - k["coeffs"])) < k["tol"]
    k = KATS["lpc_levinson"]
    s = np.sin(2 * np.pi * np.arange(8) / 8).astype(np.float32)  # sine(8), spectrum.rs:456-459
    d = c.to_device(s)
    r = c.autocorrelate(c.frames(d.ptr, 1, 8, 8, vb.WINDOW_NONE), 8).to_host()
    rn = c.normalize(r)
    assert np.max(np.abs(rn[0] - k["autocorr"])) < k["tol"]
    ac, _ = c.lpc_levinson(c.to_device(rn), k["order"])
    assert np.max(np.abs(ac.to_host()[0] - k["lpc"])) < k["tol"]
    k = KATS["resonance_from_root"]
    res, n = c.roots_to_resonances(np.array([[complex(*z) for z in k["roots"]]]), k["fs"])
    assert n[0] == 1 and np.max(np.abs(res[0, 0] - k["resonances"][0])) < k["tol"]
    k = KATS["resonances_from_lpc"]
    lpc = c.to_device(np.array([k["lpc"]], dtype=np.float64))
    out = c.lpc_to_resonances(lpc, len(k["lpc"]), False, k["fs"], strict_im=True)
    f = out["resonances"].to_host()[0, :4, 0]
    assert out["n_res"].to_host()[0] == 4 and np.max(np.abs(f - k["frequencies"])) < k["tol_hz"]
    k = KATS["formant_extractor"]
    frames = np.array([[[f, 1.0] for f in fr] for fr in k["frames"]])
    tracks, _ = c.estimate_formants(frames, np.array([[f, 1.0] for f in k["estimates"]]))
    assert np.array_equal(tracks[:, :, 0], np.array(k["tracks"]))
    k = KATS["dct"]
    assert np.max(np.abs(c.dct(np.array(k["input"]))[0] - k["output"])) < k["tol"]
    k = KATS["mel"]
    assert abs(vb.hz_to_mel(k["hz"]) - k["mel"]) < k["tol"] and abs(vb.mel_to_hz(k["mel"]) - k["hz"]) < k["tol"]
    k = KATS["laguerre"]
    z = c.laguerre(np.array([k["coeffs"]], dtype=np.complex128), complex(*k["start"]))[0]
    assert abs(z.real - k["root"][0]) < k["tol"] and abs(z.imag - k["root"][1]) < k["tol"]
    for name in ("roots_cubic", "roots_quadratic_real", "roots_quadratic_complex", "roots_linear"):
        k = KATS[name]
        roots, st = c.find_roots(np.array([k["coeffs"]], dtype=np.complex128))
        exp = np.array([complex(*z) if isinstance(z, list) else complex(z) for z in k["roots"]])
        assert st[0] == 0 and np.max(np.abs(roots[0, :exp.size] - exp)) < k["tol"], name
    k = KATS["pitch_150hz"]
    phase = np.mod(np.cumsum(np.full(k["n_samples"], k["hz"] / k["fs"])) - k["hz"] / k["fs"], 1.0)
    x = np.sin(2 * np.pi * phase).astype(np.float32)
    d = c.to_device(x)
    res = c.pitch(c.frames(d.ptr, 1, k["bin"], k["hop"], vb.WINDOW_HANN_SYMMETRIC), k["fs"], k["threshold"], k["min"], k["max"], 8)
    assert abs(res["candidates"].to_host()[0, 0, 0] - k["top_frequency"]) < k["tol_hz"]
    k = KATS["rms"]
    assert abs(c.rms(np.sin(2 * np.pi * np.arange(64) / 64))[0] - k["rms"]) < k["tol"]


@pytest.mark.gpu
def test_gpu_vs_golden_oracle_vectors():
    """Tolerances of BASELINE.json north_star: r/LPC/MFCC 1e-5 norm-wise, formants 0.5 Hz + identical tracks' slot
    assignment, pitch 0.1 Hz + identical voiced/unvoiced."""
    from gpu_util import normwise
    c, vb = _gpu()
    x, fs = VEC["audio"], float(CASES["fs"])
    d = c.to_device(x)
    p = CASES["lpc"]
    F = c.n_frames_of(x.size, p["n"], p["hop"])
    r, ac, kc = c.lpc(c.frames(d.ptr, F, p["n"], p["hop"], vb.WINDOW_HANN_SYMMETRIC), p["p"])
    assert normwise(r.to_host(), VEC["lpc_r"]).max() < 1e-5 and normwise(ac.to_host(), VEC["lpc_ac"]).max() < 1e-5
    assert normwise(kc.to_host(), VEC["lpc_kc"]).max() < 1e-5
    co, st = c.lpc_burg(c.frames(d.ptr, F, p["n"], p["hop"], vb.WINDOW_HANN_PERIODIC), p["p"])
    assert np.array_equal(st.to_host(), VEC["burg_status"]) and normwise(co.to_host(), VEC["burg_coeffs"]).max() < 1e-5
    est = np.array([[[f, 1.0] for f in (320., 1440., 2760., 3200.)]])
    for name, method, win in (("formants_burg", vb.LPC_BURG, vb.WINDOW_HANN_PERIODIC),
                              ("formants_autocorr", vb.LPC_AUTOCORR, vb.WINDOW_HANN_SYMMETRIC)):
        out = c.find_formants(c.frames(d.ptr, F, p["n"], p["hop"], win), fs, p["p"], method, est)
        assert np.array_equal(out["n_res"], VEC[name + "_nres"]), name
        assert np.max(np.abs(out["resonances"] - VEC[name + "_res"])) < 0.5, name
        assert np.max(np.abs(out["tracks"] - VEC[name + "_tracks"])) < 0.5, name
    q = CASES["pitch"]
    Fp = c.n_frames_of(x.size, q["n"], q["hop"])
    res = c.pitch(c.frames(d.ptr, Fp, q["n"], q["hop"], vb.WINDOW_HANN_SYMMETRIC), fs, q["thr"], q["fmin"], q["fmax"], q["k"])
    cand, n = res["candidates"].to_host(), res["n_cand"].to_host()
    assert np.array_equal(n, VEC["pitch_n"]) and np.array_equal(res["status"].to_host(), VEC["pitch_status"])
    assert np.array_equal(cand[:, 0, 0] != 0, VEC["pitch_cand"][:, 0, 0] != 0)
    assert np.max(np.abs(cand[:, 0, 0] - VEC["pitch_cand"][:, 0, 0])) < 0.1
    m = CASES["mfcc"]
    out = c.mfcc(c.frames(d.ptr, F, m["n"], m["hop"], vb.WINDOW_HANN_SYMMETRIC), m["m"], m["lo"], m["hi"], fs, n_keep=m["keep"])
    assert normwise(out.to_host(), VEC["mfcc"]).max() < 1e-5
    w = VEC["waves_in"]
    assert np.allclose(c.rms(w), VEC["waves_rms"], rtol=1e-14) and np.array_equal(c.max_amplitude(w), VEC["waves_max"])
    assert np.allclose(c.normalize(w), VEC["waves_norm"], rtol=1e-15) and np.max(np.abs(c.preemphasis(w, 0.05) - VEC["waves_pre"])) < 1e-13
