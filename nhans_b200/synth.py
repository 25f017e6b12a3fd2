"""Synthetic 16 kHz int16 inputs for tests and benchmarks (SURVEY.md §8 d).

There is no network for datasets, so every measured configuration uses these deterministic
signals: a harmonic 'speech-like' sweep with 3 Hz amplitude modulation plus pink-ish noise at
0 dB, scaled to peak 20 000.  Context clips are 3.0 s (>= 32 240 samples are needed for the 200
context frames the reference slices, SN/apply.py:381-382)."""
from __future__ import annotations

import numpy as np

FS = 16000


def _pinkish(rng, n):
    white = rng.standard_normal(n)
    out = np.empty(n)
    acc = 0.0
    # y[n] = x[n] + 0.9 y[n-1]
    try:
        from scipy.signal import lfilter
        return lfilter([1.0], [1.0, -0.9], white)
    except Exception:  # pragma: no cover
        for i in range(n):
            acc = white[i] + 0.9 * acc
            out[i] = acc
        return out


def speech_like(seconds, seed):
    n = int(round(seconds * FS))
    t = np.arange(n) / FS
    f0 = 110.0 + (220.0 - 110.0) * t / max(seconds, 1e-9)
    phase = 2 * np.pi * np.cumsum(f0) / FS
    s = sum((1.0 / h) * np.sin(h * phase) for h in range(1, 9))
    return s * 0.5 * (1 + np.sin(2 * np.pi * 3 * t))


def _to_int16(x, peak=20000.0):
    x = x / (np.max(np.abs(x)) + 1e-12) * peak
    return np.round(x).astype(np.int16)


def mixture(seconds, u):
    """Utterance ``u``: speech-like + noise at 0 dB (seed 1000+u)."""
    rng = np.random.default_rng(1000 + u)
    s = speech_like(seconds, 1000 + u)
    nz = _pinkish(rng, len(s))
    nz *= np.sqrt(np.mean(s * s) / (np.mean(nz * nz) + 1e-12))
    return _to_int16(s + nz)


def noise_clip(u, kind="neg", seconds=3.0):
    """Context clip: seed 2000+u (neg) / 3000+u (pos)."""
    seed = (2000 if kind == "neg" else 3000) + u
    rng = np.random.default_rng(seed)
    return _to_int16(_pinkish(rng, int(round(seconds * FS))))


def speaker_clip(u, kind="target", seconds=3.0):
    """Separator context: a second speech-like signal (seed 4000+u target / 5000+u interference)."""
    seed = (4000 if kind == "target" else 5000) + u
    rng = np.random.default_rng(seed)
    s = speech_like(seconds, seed) * (1.0 + 0.1 * rng.standard_normal())
    s = np.roll(s, int(rng.integers(0, len(s))))
    return _to_int16(s + 0.01 * rng.standard_normal(len(s)))


def silence(seconds=3.0):
    """Digital silence: what ``apply_denoiser`` feeds as --pos (Silent.wav, SN/apply.py:478-481)."""
    return np.zeros(int(round(seconds * FS)), dtype=np.int16)


def batch(n_utts, seconds, with_pos=False, first=0):
    """Lists of int16 arrays: (mix, ctx_pos_or_None, ctx_neg)."""
    mix = [mixture(seconds, first + u) for u in range(n_utts)]
    neg = [noise_clip(first + u, "neg") for u in range(n_utts)]
    pos = [noise_clip(first + u, "pos") for u in range(n_utts)] if with_pos else None
    return mix, pos, neg
