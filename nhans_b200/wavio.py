"""16 kHz 16-bit PCM wav I/O for the CLIs (the reference uses scipy.io.wavfile the same way,
N_HANS___Selective_Noise/apply.py:23-25, 46-53, 201-202)."""
from __future__ import annotations

import numpy as np
from scipy.io.wavfile import read as _wavread, write as _wavwrite

FS = 16000


def read_wav(in_path):
    """read_wav of SN/apply.py:46-53: asserts 16 kHz and int16.  Mono files come back as int16 (what the C ABI
    takes).  Stereo files are averaged exactly like the reference (`samples.mean(axis=1)`, float64 with
    half-integer values): the caller routes such clips through the float entry points (Engine.enhance_float),
    so nothing is rounded back to int16."""
    rate, samples = _wavread(in_path)
    assert rate == FS, "%s: sample rate %d, expected %d" % (in_path, rate, FS)
    assert samples.dtype == np.int16, "%s: dtype %s, expected int16" % (in_path, samples.dtype)
    if samples.ndim > 1:
        samples = samples.mean(axis=1)
    assert samples.ndim == 1
    return np.ascontiguousarray(samples)


def is_pcm16(x):
    return np.asarray(x).dtype == np.int16


def normalise_host(samples):
    """handle_signals' normalisation on the host (SN/apply.py:150-152): samples / (max|samples| + 1e-6) in float64,
    then float32.  For int16 input `abs` wraps -32768 exactly like the reference's numpy call does."""
    x = np.asarray(samples)
    if len(x) == 0:
        return np.zeros(0, np.float32)
    return (x / (float(max(abs(x))) + 0.000001)).astype(np.float32)


def write_wav(path, samples, rate=FS):
    _wavwrite(path, rate, np.asarray(samples))
