#!/usr/bin/env python
"""Regenerates tests/golden/oracle_vectors.npz: outputs of the CPU oracle (oracle/, the f64 restatement of the
reference) on small seeded inputs, one entry per hot-path operator.  The reference itself cannot be run in this
image (Rust crate, no cargo), so these vectors are ORACLE outputs, not reference outputs; the reference's own
asserted known answers live next to them in reference_kats.json.  They serve two purposes: the GPU tests check
the CUDA path against committed numbers (no oracle in the loop), and the CPU tests detect drift of the oracle.

    python tests/golden/make_golden.py        # rewrites oracle_vectors.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vox_box.rs_b200", "python"))

CASES = dict(fs=16000, seed_utt=77, seconds=0.4, lpc=dict(n=400, hop=160, p=12), burg=dict(n=400, hop=160, p=12),
             formants=dict(n=400, hop=160, p=12), pitch=dict(n=640, hop=160, thr=0.45, fmin=75.0, fmax=600.0, k=24),
             mfcc=dict(n=400, hop=160, m=40, keep=13, lo=133.0, hi=6855.0))


def build():
    import oracle
    from voxbox_b200 import synth
    oracle.build()
    fs = CASES["fs"]
    x = synth.utterance(CASES["seed_utt"], fs, seconds=CASES["seconds"])
    out = {"audio": x}
    c = CASES["lpc"]
    F = oracle.n_frames_of(x.size, c["n"], c["hop"])
    r, ac, kc = oracle.batch_lpc(x, F, c["n"], c["hop"], oracle.WIN_HANN_SYMMETRIC, c["p"], want_kc=True)
    out.update(lpc_r=r, lpc_ac=ac, lpc_kc=kc)
    c = CASES["burg"]
    co, st = oracle.batch_burg(x, F, c["n"], c["hop"], oracle.WIN_HANN_PERIODIC, c["p"])
    out.update(burg_coeffs=co, burg_status=st)
    c = CASES["formants"]
    est = np.array([[f, 1.0] for f in (320., 1440., 2760., 3200.)])
    for name, method, win in (("formants_burg", 0, oracle.WIN_HANN_PERIODIC), ("formants_autocorr", 1, oracle.WIN_HANN_SYMMETRIC)):
        d = oracle.batch_formants(x, F, c["n"], c["hop"], win, method, float(fs), c["p"], np.array([0, F]), est)
        out.update({name + "_tracks": d["tracks"], name + "_res": d["resonances"], name + "_nres": d["n_res"]})
    c = CASES["pitch"]
    Fp = oracle.n_frames_of(x.size, c["n"], c["hop"])
    cand, nc, st = oracle.batch_pitch(x, Fp, c["n"], c["hop"], oracle.WIN_HANN_SYMMETRIC, float(fs), c["thr"], c["fmin"], c["fmax"], c["k"])
    out.update(pitch_cand=cand, pitch_n=nc, pitch_status=st)
    c = CASES["mfcc"]
    out["mfcc"] = oracle.batch_mfcc(x, F, c["n"], c["hop"], oracle.WIN_HANN_SYMMETRIC, c["m"], c["lo"], c["hi"], float(fs), n_keep=c["keep"])
    rng = np.random.default_rng(5)
    w = rng.standard_normal((3, 50))
    out.update(waves_in=w, waves_rms=np.array([oracle.rms(v) for v in w]), waves_max=np.array([oracle.max_amplitude(v) for v in w]),
               waves_norm=np.stack([oracle.normalize(v) for v in w]), waves_pre=np.stack([oracle.preemphasis(v, 0.05) for v in w]))
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **build())
    print("wrote", os.path.join(HERE, "oracle_vectors.npz"))
