"""GPU parity of the formant path through the C ABI vs the f64 oracle: Burg LPC, polynomial roots,
resonances, the McCandless tracker and the composed find_formants.

Tolerances (BASELINE.json north_star): LPC within 1e-5 relative (norm-wise per frame), formant
frequencies and bandwidths within 0.5 Hz, identical formant-track assignments."""
import os

import numpy as np
import pytest

from gpu_util import ctx, normwise, synth, vb

pytestmark = pytest.mark.gpu
TOL_LPC = 1e-5
TOL_HZ = 0.5


# ------------------------------------------------------------------------------------------ Burg
def test_burg_kat(oracle):  # spectrum.rs:515-525 test_lpc_praat, 1e-10
    c = ctx()
    src = np.array(list(range(1, 11)) + list(range(10, 0, -1)), dtype=np.float32)
    d = c.to_device(src)
    co, st = c.lpc_burg(c.frames(d.ptr, 1, 20, 20, vb.WINDOW_NONE), 5)
    exp = [-2.529731754197289, 2.6138925001574935, -1.6951059551991234, 0.7776548472652218, -0.15008712022777612]
    assert st.to_host()[0] == 0
    assert np.all(np.abs(co.to_host()[0] - exp) < 1e-10)


@pytest.mark.parametrize("N,hop,p,window", [(400, 160, 12, vb.WINDOW_HANN_PERIODIC), (1102, 441, 12, vb.WINDOW_HANN_PERIODIC),
                                             (1024, 512, 10, vb.WINDOW_HANN_PERIODIC), (64, 64, 4, vb.WINDOW_NONE),
                                             (257, 100, 16, vb.WINDOW_HANN_SYMMETRIC), (1153, 400, 8, vb.WINDOW_NONE)])
def test_burg_synthetic(oracle, N, hop, p, window):
    """Warp-register kernel (frame_len <= 1153) incl. a length that needs the last lane's padding."""
    audio = synth.utterance(5, 16000, seconds=1.0)
    c = ctx()
    F = min(c.n_frames_of(audio.size, N, hop), 60)
    d = c.to_device(audio)
    co, st = c.lpc_burg(c.frames(d.ptr, F, N, hop, window), p)
    ref, ref_st = oracle.batch_burg(audio, F, N, hop, window, p)
    assert np.array_equal(st.to_host(), ref_st)
    assert np.max(normwise(co.to_host(), ref)) < TOL_LPC
    assert np.max(normwise(co.to_host(), ref)) < 1e-9  # fp64 throughout: only summation order differs


def test_burg_block_kernel_and_long_frame(oracle, fixtures_dir, monkeypatch):
    """The any-length CTA-per-frame kernel: forced on a short frame, and natural on the 31 232-sample
    whole-file frame of tests/lib.rs:13-42 (order 13)."""
    c = ctx()
    audio = synth.utterance(6, 16000, seconds=0.5)
    d = c.to_device(audio)
    monkeypatch.setenv("VBX_BURG_FORCE_BLOCK", "1")
    co, st = c.lpc_burg(c.frames(d.ptr, 20, 400, 160, vb.WINDOW_HANN_PERIODIC), 12)
    monkeypatch.delenv("VBX_BURG_FORCE_BLOCK")
    ref, _ = oracle.batch_burg(audio, 20, 400, 160, oracle.WIN_HANN_PERIODIC, 12)
    assert np.max(normwise(co.to_host(), ref)) < 1e-9
    x, fs = oracle.read_wav(os.path.join(fixtures_dir, "down_sampled.wav"))
    xf = x.astype(np.float32)
    d = c.to_device(xf)
    co, st = c.lpc_burg(c.frames(d.ptr, 1, xf.size, xf.size, vb.WINDOW_HANN_PERIODIC), 13)
    ref, _ = oracle.batch_burg(xf, 1, xf.size, xf.size, oracle.WIN_HANN_PERIODIC, 13)
    assert st.to_host()[0] == 0 and np.max(normwise(co.to_host(), ref)) < 1e-8


def test_burg_denum_error(oracle):  # spectrum.rs:123-125: all-zero frame → Err(LPC("Denum was <= 0.0"))
    c = ctx()
    audio = np.zeros(800, dtype=np.float32)
    audio[400:] = synth.utterance(1, 16000, seconds=0.1)[:400]
    d = c.to_device(audio)
    co, st = c.lpc_burg(c.frames(d.ptr, 2, 400, 400, vb.WINDOW_NONE), 8)
    ref, ref_st = oracle.batch_burg(audio, 2, 400, 400, oracle.WIN_NONE, 8)
    assert st.to_host().tolist() == [vb.ERR_LPC, 0] == ref_st.tolist()
    assert np.all(np.isnan(co.to_host()[0])) and normwise(co.to_host()[1], ref[1]) < 1e-9


# ------------------------------------------------------------------------------------------ polynomial.rs
def test_find_roots_kats(oracle):  # polynomial.rs:295-377
    c = ctx()
    for dt, tol in ((np.complex128, 1e-12), (np.complex64, 1e-6)):
        r, st = c.find_roots(np.array([[1.0, 2.5]], dtype=dt))
        assert st[0] == 0 and abs(r[0, 0] - (-0.4)) < tol and r[0, 1] == 0
        r, st = c.find_roots(np.array([[1.0, 2.5, -2.0]], dtype=dt))
        assert st[0] == 0 and abs(r[0, 0] - (-0.31872930440884)) < tol and abs(r[0, 1] - 1.5687293044088) < tol
        r, st = c.find_roots(np.array([[1.0, -2.5, 2.0]], dtype=dt))
        assert abs(r[0, 0] - complex(0.625, -0.33071891388307)) < tol and abs(r[0, 1] - complex(0.625, 0.33071891388307)) < tol
        r, st = c.find_roots(np.array([[1.0, 2.5, -2.0, -3.0]], dtype=dt))  # test_hi_d_roots: order-sensitive
        exp = [-1.1409835232292, -0.35308705904629, 0.82740391560878]
        assert st[0] == 0 and np.all(np.abs(r[0, :3] - exp) < 1e-6) and r[0, 3] == 0


def test_find_roots_errors_and_order(oracle):
    c = ctx()
    r, st = c.find_roots(np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 0.0]], dtype=np.complex128))
    assert st.tolist() == [vb.ERR_POLYNOMIAL, vb.ERR_POLYNOMIAL]  # "Zero degree polynomial"
    # order-8 LPC polynomial of spectrum.rs:616-633: same roots in the same order as the oracle
    co = [-0.80098309, 1.20869679, -1.61846677, 0.86630291, -1.44203292, 0.93621726, -0.58772811, 0.65949051]
    poly = np.concatenate([[1.0], co])[::-1].astype(np.complex128)
    r, st = c.find_roots(poly[None, :])
    _, ref, _ = oracle.find_roots_mut(poly)
    assert st[0] == 0 and np.max(np.abs(r[0] - ref)) < 1e-9
    res, n = c.roots_to_resonances(r[:, :8], 11025.0)
    assert n[0] == 4 and np.all(np.abs(res[0, :4, 0] - [251.770, 2289.634, 3037.846, 4045.196]) < 1e-2)


def test_laguerre_kat(oracle):  # polynomial.rs:282-292
    c = ctx()
    z = c.laguerre(np.array([[1.0, 2.5, 2.0, 3.0]], dtype=np.complex128), complex(-64.0, -64.0))
    assert abs(z[0] - complex(-0.1070229535872, -0.8514680262155)) < 1e-8
    coef = [1.0, -0.99640256, 0.25383306, -0.25471634, 0.5084799, -0.0685858, -0.35042483, 0.07676613, -0.12874511,
            0.11829436, 0.023972526]
    z = c.laguerre(np.array([coef], dtype=np.complex64), complex(-64.0, -64.0))  # polynomial.rs:380-386: finite
    assert np.isfinite(z[0].real) and np.isfinite(z[0].imag)


def test_div_polynomial(oracle):  # polynomial.rs:155-195
    c = ctx()
    q, rem, st = c.div_polynomial(np.array([[2.0, 3.0, 1.0]], dtype=np.complex128), complex(1.0, 0.0))
    assert st[0] == 0 and np.allclose(q[0], [2.0, 1.0, 0.0]) and abs(rem[0, 0]) < 1e-15
    _, _, st = c.div_polynomial(np.array([[2.0, 3.0, 1.0]], dtype=np.complex128), complex(0.0, 0.0))
    assert st[0] == vb.ERR_POLYNOMIAL


def test_resonance_kat(oracle):  # spectrum.rs:462-468
    c = ctx()
    roots = np.array([[complex(-0.5, 0.86602540378444), complex(-0.5, -0.86602540378444)]])
    res, n = c.roots_to_resonances(roots, 300.0)
    assert n[0] == 1 and abs(res[0, 0, 0] - 100.0) < 1e-8 and abs(res[0, 0, 1]) < 1e-8


# ------------------------------------------------------------------------------------------ McCandless
def test_formant_extractor_kat(oracle):  # spectrum.rs:528-567, exact
    c = ctx()
    frames = np.array([[100.0, 150.0, 200.0, 240.0, 300.0], [110.0, 180.0, 210.0, 230.0, 310.0],
                       [230.0, 270.0, 290.0, 350.0, 360.0]])
    res = np.stack([frames, np.ones_like(frames)], axis=-1)
    est = np.array([[140.0, 1.0], [230.0, 1.0], [320.0, 1.0]])
    tracks, final = c.estimate_formants(res, est)
    assert tracks[0, :, 0].tolist() == [150.0, 240.0, 300.0]
    assert tracks[1, :, 0].tolist() == [180.0, 230.0, 310.0]
    assert tracks[2, :, 0].tolist() == [230.0, 270.0, 290.0]
    assert np.array_equal(final[0], tracks[2])


def test_mccandless_random_vs_oracle(oracle):
    """Random resonance frames incl. zero padding, duplicates and ties: the slot logic must be identical."""
    rng = np.random.default_rng(11)
    c = ctx()
    n_seg, J, slots = 40, 25, 32
    res = np.zeros((n_seg * J, slots, 2))
    for f in range(n_seg * J):
        k = rng.integers(0, 8)
        fr = np.sort(rng.choice(np.arange(60.0, 5000.0, 10.0), k, replace=False))  # coarse grid ⇒ ties happen
        res[f, :k, 0] = fr
        res[f, :k, 1] = rng.uniform(20, 400, k)
    for n_est in (1, 3, 4, 6, 7):
        est0 = np.tile(np.stack([np.linspace(300, 3500, n_est), np.ones(n_est)], -1), (n_seg, 1, 1))
        tracks, final = c.estimate_formants(res, est0, n_segments=n_seg)
        for u in range(n_seg):
            t_ref, f_ref = oracle.formant_extractor(est0[u], res[u * J:(u + 1) * J])
            assert np.array_equal(tracks[u * J:(u + 1) * J], t_ref), (n_est, u)
            assert np.array_equal(final[u], f_ref)


# ------------------------------------------------------------------------------------------ find_formants
def _male():
    return np.array([[f, 1.0] for f in (320.0, 1440.0, 2760.0, 3200.0)])


def test_find_formants_short_sample(oracle, fixtures_dir):
    """tests/lib.rs:44-90 through the GPU: bin 1024 / hop 512, order 10, MALE estimates."""
    x, fs = oracle.read_wav(os.path.join(fixtures_dir, "short_sample.wav"))
    xf = x.astype(np.float32)
    c = ctx()
    F = c.n_frames_of(xf.size, 1024, 512)
    assert F == 4
    out = c.find_formants_host(xf, F, 1024, 512, vb.WINDOW_HANN_PERIODIC, fs, 10, vb.LPC_BURG, _male())
    ref = oracle.batch_formants(xf, F, 1024, 512, oracle.WIN_HANN_PERIODIC, 0, fs, 10, [0, F], _male())
    assert np.all(out["status"] == 0) and np.array_equal(out["n_res"], ref["n_res"])
    assert np.max(np.abs(out["resonances"] - ref["resonances"])) < TOL_HZ
    assert np.max(np.abs(out["tracks"] - ref["tracks"])) < TOL_HZ
    exp0 = [(1030.918, 264.413), (2724.528, 320.901), (3719.483, 114.118), (3200.0, 1.0)]
    exp2 = [(1025.91, 332.98), (2695.679, 277.572), (2695.679, 277.572), (3709.671, 116.011)]
    assert np.all(np.abs(out["tracks"][0] - np.array(exp0)) < 5e-3)  # fp32 samples vs the f64 WAV scaling
    assert np.all(np.abs(out["tracks"][2] - np.array(exp2)) < 5e-3)
    assert np.array_equal(out["estimates"][0], out["tracks"][3])


def test_find_formants_against_praat_frame(oracle, fixtures_dir):
    """tests/lib.rs:13-42: the whole 31 232-sample file as ONE frame, order 13."""
    x, fs = oracle.read_wav(os.path.join(fixtures_dir, "down_sampled.wav"))
    xf = x.astype(np.float32)
    c = ctx()
    out = c.find_formants_host(xf, 1, xf.size, xf.size, vb.WINDOW_HANN_PERIODIC, fs, 13, vb.LPC_BURG, _male())
    exp = [(179.102, 472.0368), (998.382, 273.192), (2358.3585, 660.5514), (3082.1675, 221.7892)]
    assert out["status"][0] == 0 and np.all(np.abs(out["tracks"][0] - np.array(exp)) < 2e-2)


@pytest.mark.parametrize("fs,N,hop,method,window", [
    (16000, 400, 160, vb.LPC_AUTOCORR, vb.WINDOW_HANN_SYMMETRIC),
    (44100, 1102, 441, vb.LPC_AUTOCORR, vb.WINDOW_HANN_SYMMETRIC),   # C3 path A
    (44100, 1102, 441, vb.LPC_BURG, vb.WINDOW_HANN_PERIODIC),        # C3 path B (find_formants parity)
    (16000, 400, 160, vb.LPC_BURG, vb.WINDOW_HANN_PERIODIC),
])
def test_find_formants_synthetic_utterances(oracle, fs, N, hop, method, window):
    """Several utterances in one call (two-level view), tracker state per utterance, order 12."""
    n_utts, seconds, p = 6, 1.5, 12
    audio = synth.corpus(n_utts, fs, seconds)
    n_samp = audio.shape[1]
    c = ctx()
    J = c.n_frames_of(n_samp, N, hop)
    F = n_utts * J
    d = c.to_device(audio)
    fr = c.frames(d.ptr, F, N, hop, window, frames_per_segment=J, segment_stride=n_samp)
    out = c.find_formants(fr, float(fs), p, method, np.tile(_male(), (n_utts, 1, 1)))
    worst_res = worst_trk = 0.0
    mism_count = mism_assign = 0
    for u in range(n_utts):
        ref = oracle.batch_formants(audio[u], J, N, hop, window, 0 if method == vb.LPC_BURG else 1, float(fs), p, [0, J],
                                    _male(), n_threads=0)
        sl = slice(u * J, (u + 1) * J)
        assert np.array_equal(out["status"][sl], ref["status"])
        mism_count += int(np.sum(out["n_res"][sl] != ref["n_res"]))
        same = out["n_res"][sl] == ref["n_res"]
        worst_res = max(worst_res, float(np.max(np.abs(out["resonances"][sl][same] - ref["resonances"][same]))))
        d_trk = np.abs(out["tracks"][sl] - ref["tracks"])
        mism_assign += int(np.sum(np.max(d_trk, axis=(1, 2)) > TOL_HZ))
        worst_trk = max(worst_trk, float(np.max(np.where(d_trk > TOL_HZ, 0.0, d_trk))))
        assert np.max(np.abs(out["estimates"][u] - ref["tracks"][-1])) < TOL_HZ or mism_assign
    # report, never mask: every frame must have the same resonance count and the same track assignment
    assert mism_count == 0, f"{mism_count} frames with a different number of resonances"
    assert mism_assign == 0, f"{mism_assign} frames whose formant tracks differ by more than {TOL_HZ} Hz"
    assert worst_res < TOL_HZ and worst_trk < TOL_HZ
    assert worst_res < 1e-3  # fp64-polished roots: far inside the tolerance


@pytest.mark.parametrize("fs,N,hop", [(16000, 400, 160), (44100, 1102, 441)])
@pytest.mark.parametrize("p", [2, 3, 4, 8, 11, 12, 13, 16, 24])
def test_pair_deflation_matches_one_at_a_time(fs, N, hop, p):
    """The fused fp32 path divides conjugate pairs out of the real LPC polynomial (lpc_roots_pair_kernel); a call that asks
    for the roots keeps the reference's one-root-at-a-time order (lpc_roots_rt_kernel).  Same polynomial, same fp64
    polish: the resonances must be the same set — identical counts, values far inside the 0.5 Hz tolerance — for even and
    odd orders (an odd order has at least one real root: the linear-factor branch), both `im` filters, Levinson and Burg."""
    audio = synth.utterance(31 + p, fs, seconds=2.0)
    c = ctx()
    F = c.n_frames_of(audio.size, N, hop)
    d = c.to_device(audio)
    _, ac, _ = c.lpc(c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC), p)
    burg, _ = c.lpc_burg(c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_PERIODIC), p)
    for lpc, has_one in ((ac, True), (burg, False)):
        for strict in (True, False):
            a = c.lpc_to_resonances(lpc, p, has_one, float(fs), strict_im=strict)                    # pair deflation, fp32 + polish
            b = c.lpc_to_resonances(lpc, p, has_one, float(fs), strict_im=strict, want_roots=True)   # one at a time, fp32 + polish
            t = c.lpc_to_resonances(lpc, p, has_one, float(fs), strict_im=strict, want_roots=True, precision=1)  # one at a time, fp64
            for x in (a, b):
                assert np.array_equal(x["status"].to_host(), t["status"].to_host())
                assert np.array_equal(x["n_res"].to_host(), t["n_res"].to_host()), (has_one, strict)
            # against the fp64 path: the pair kernel (real-coefficient deflation) stays at rounding level even at order 24,
            # where the fp32 one-at-a-time deflation with complex coefficients drifts by 0.05 Hz at 16 kHz and 1.6 Hz at
            # 44.1 kHz (a caller who wants the root list at such orders asks for precision = 1)
            assert np.max(np.abs(a["resonances"].to_host() - t["resonances"].to_host())) < 1e-3, (has_one, strict)
            if p <= 16:
                assert np.max(np.abs(b["resonances"].to_host() - t["resonances"].to_host())) < 0.5, (has_one, strict)


def test_roots_precisions_agree(oracle):
    """fp32 Laguerre + fp64 polish (default) and the fp64 Laguerre path give the same resonances."""
    audio = synth.utterance(9, 16000, seconds=2.0)
    c = ctx()
    N, hop, p = 400, 160, 12
    F = c.n_frames_of(audio.size, N, hop)
    d = c.to_device(audio)
    _, ac, _ = c.lpc(c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_SYMMETRIC), p)
    a = c.lpc_to_resonances(ac, p, True, 16000.0, precision=0, want_roots=True)
    b = c.lpc_to_resonances(ac, p, True, 16000.0, precision=1, want_roots=True)
    assert np.array_equal(a["n_res"].to_host(), b["n_res"].to_host())
    assert np.max(np.abs(a["resonances"].to_host() - b["resonances"].to_host())) < 1e-4
    # both find the same multiset of roots as the oracle's f64 find_roots
    ach = ac.to_host()
    rb = b["roots"].to_host()
    for f in range(0, F, 17):
        _, ref, _ = oracle.find_roots_mut(ach[f][::-1].astype(np.complex128))
        got = rb[f, :, 0] + 1j * rb[f, :, 1]
        left = list(ref[:p])
        for z in got:  # same multiset: every root pairs off with a distinct oracle root
            j = int(np.argmin([abs(z - w) for w in left]))
            assert abs(z - left[j]) < 1e-6
            left.pop(j)


def test_find_formants_empty_and_bad_args(oracle):
    c = ctx()
    d = c.to_device(np.zeros(2048, dtype=np.float32))
    out = c.find_formants(c.frames(d.ptr, 0, 400, 160, vb.WINDOW_HANN_PERIODIC), 16000.0, 12, vb.LPC_BURG,
                          np.zeros((0, 4, 2)))
    assert out["tracks"].shape == (0, 4, 2)
    with pytest.raises(vb.VoxBoxError) as e:
        c.find_formants(c.frames(d.ptr, 1, 400, 160, vb.WINDOW_HANN_PERIODIC), 16000.0, 40, vb.LPC_BURG, _male()[None])
    assert e.value.status == vb.ERR_BADARG
    # silent frame: Burg fails (denum <= 0) → status LPC, estimates untouched (lib.rs:75 `?`)
    out = c.find_formants(c.frames(d.ptr, 2, 400, 160, vb.WINDOW_HANN_PERIODIC, frames_per_segment=2, segment_stride=0),
                          16000.0, 12, vb.LPC_BURG, _male()[None])
    assert out["status"].tolist() == [vb.ERR_LPC, vb.ERR_LPC]
    assert np.array_equal(out["tracks"][1], _male())


@pytest.mark.parametrize("fs,N,hop,ratio,p", [(44100, 2205, 441, 10000.0 / 44100.0, 13),   # examples/formant_extraction: 50 ms / 10 ms → 10 kHz
                                               (16000, 400, 160, 1.5, 10),                  # up-sampling
                                               (16000, 512, 256, 0.5, 8)])
def test_find_formants_resampled(oracle, fs, N, hop, ratio, p):
    """lib.rs:57-61: linear resampling inside find_formants (parity vs the oracle's restatement of
    sample::interpolate::{Linear, Converter}; the reference itself never tests resample_ratio != 1)."""
    c = ctx()
    audio = synth.utterance(13, fs, seconds=0.6)
    F = min(c.n_frames_of(audio.size, N, hop), 24)
    fs_res = fs * ratio
    for arr, dt in ((audio, vb.F32), (audio.astype(np.float64), vb.F64)):
        d = c.to_device(arr)
        out = c.find_formants_resampled(c.frames(d.ptr, F, N, hop, vb.WINDOW_NONE, dtype=dt), fs_res, ratio, p, _male()[None])
        state = _male().copy()
        for f in range(F):
            o = oracle.find_formants(arr[f * hop: f * hop + N].astype(np.float64), fs_res, p, state, resample_ratio=ratio)
            assert o["status"] == out["status"][f] == 0
            state = o["formants"]
            assert o["n_res"] == out["n_res"][f], f
            assert np.max(np.abs(out["resonances"][f] - o["resonances"])) < TOL_HZ, f
            assert np.max(np.abs(out["tracks"][f] - state)) < TOL_HZ, f
    # ratio 1 takes the plain-copy branch (lib.rs:62-64)
    d = c.to_device(audio)
    a = c.find_formants_resampled(c.frames(d.ptr, F, N, hop, vb.WINDOW_NONE), float(fs), 1.0, p, _male()[None])
    b = c.find_formants(c.frames(d.ptr, F, N, hop, vb.WINDOW_HANN_PERIODIC), float(fs), p, vb.LPC_BURG, _male()[None])
    assert np.array_equal(a["tracks"], b["tracks"])


def test_burg_f64_samples(oracle):
    """lpc_praat on f64 slices (the reference's own instantiation): VBX_F64 samples, no fp32 rounding of the input."""
    c = ctx()
    rng = np.random.default_rng(4)
    x = rng.standard_normal(3 * 300) * np.hanning(900)
    d = c.to_device(x)
    co, st = c.lpc_burg(c.frames(d.ptr, 3, 300, 300, vb.WINDOW_NONE, dtype=vb.F64), 10)
    for f in range(3):
        s_, ref = oracle.lpc_praat(x[f * 300:(f + 1) * 300], 10)
        assert s_ == 0 and st.to_host()[f] == 0
        assert np.max(np.abs(co.to_host()[f] - ref)) < 1e-9 * max(1.0, np.max(np.abs(ref)))
