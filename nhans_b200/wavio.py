"""16 kHz 16-bit PCM wav I/O for the CLIs (the reference uses scipy.io.wavfile the same way,
N_HANS___Selective_Noise/apply.py:23-25, 46-53, 201-202)."""
from __future__ import annotations

import numpy as np
from scipy.io.wavfile import read as _wavread, write as _wavwrite

FS = 16000


def read_wav(in_path):
    """read_wav of SN/apply.py:46-53: asserts 16 kHz and int16.  Stereo files are averaged like the
    reference (`samples.mean(axis=1)`); because the C ABI takes int16 PCM the mean is rounded half to even,
    a <= 0.5 LSB deviation from the reference's float64 mean (exact for Silent.wav, which is all zeros)."""
    rate, samples = _wavread(in_path)
    assert rate == FS, "%s: sample rate %d, expected %d" % (in_path, rate, FS)
    assert samples.dtype == np.int16, "%s: dtype %s, expected int16" % (in_path, samples.dtype)
    if samples.ndim > 1:
        samples = np.rint(samples.mean(axis=1)).astype(np.int16)
    assert samples.ndim == 1
    return np.ascontiguousarray(samples)


def write_wav(path, samples, rate=FS):
    _wavwrite(path, rate, np.asarray(samples))
