"""Deterministic synthetic speech-like audio (SURVEY.md §8d "Synthetic audio").

Per utterance u (seed 0x5EED ^ u): a piecewise-linear f0 contour (80–300 Hz, one
knot per 0.25 s, 30 % of the segments unvoiced); the source is a unit impulse train
at the running period when voiced and N(0,1)·0.1 when unvoiced; it is shaped by a
cascade of 4 (fs <= 16 kHz) or 5 two-pole resonators with per-utterance centres
and bandwidths; the result is scaled to peak 0.5 and white noise at −40 dB re peak
is added (mandatory: without a noise floor LPC is ill-posed even in f64,
SURVEY §7.3); samples are stored as fp32.  Host-side (numpy/scipy): data
generation is outside every timed region.
"""
import numpy as np
from scipy.signal import lfilter

_CENTRES = [(300., 900.), (900., 2200.), (2200., 3200.), (3200., 4200.), (4200., 5500.)]


def utterance(u, fs, seconds=10.0, noise_db=-40.0):
    rng = np.random.default_rng(0x5EED ^ int(u))
    n = int(round(fs * seconds))
    seg = int(round(0.25 * fs))
    n_knots = n // seg + 2
    f0_knots = rng.uniform(80.0, 300.0, n_knots)
    voiced_seg = rng.uniform(0, 1, n_knots) >= 0.30
    t = np.arange(n) / seg
    f0 = np.interp(t, np.arange(n_knots), f0_knots)
    voiced = voiced_seg[np.minimum((t).astype(np.int64), n_knots - 1)]
    phase = np.cumsum(f0 / fs)
    pulses = np.diff(np.floor(phase), prepend=0.0) > 0
    src = np.where(voiced, pulses.astype(np.float64), 0.1 * rng.standard_normal(n))
    n_res = 4 if fs <= 16000 else 5
    y = src
    for lo, hi in _CENTRES[:n_res]:
        fc = rng.uniform(lo, hi)
        bw = rng.uniform(50.0, 300.0)
        r = np.exp(-np.pi * bw / fs)
        th = 2 * np.pi * fc / fs
        y = lfilter([1.0], [1.0, -2 * r * np.cos(th), r * r], y)
    y = y * (0.5 / max(np.max(np.abs(y)), 1e-30))
    y = y + 0.5 * 10 ** (noise_db / 20.0) * rng.standard_normal(n)
    return y.astype(np.float32)


def corpus(n_utts, fs, seconds=10.0, first=0):
    """[n_utts, n_samples] fp32, utterances first..first+n_utts-1."""
    return np.stack([utterance(first + u, fs, seconds) for u in range(n_utts)])
