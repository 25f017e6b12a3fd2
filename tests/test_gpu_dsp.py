"""CUDA STFT / iSTFT kernels through the C ABI against the oracle (bit-exact framing, 1e-3 spectra)."""
import numpy as np
import pytest

from nhans_b200 import synth
from oracle import nhans_oracle as O

pytestmark = pytest.mark.gpu

LOGMAG_TOL = 1e-3          # |d log-magnitude| == relative magnitude error (north_star: 1e-3 relative)
WAVE_TOL = 1e-5


def _clips():
    loud = synth.mixture(0.3, 3).copy()
    loud[100] = -32768                                   # numpy abs(int16) wrap case (SN/apply.py:150)
    return [synth.mixture(1.0, 0), synth.mixture(0.537, 1), synth.mixture(0.03, 2)[:450],
            np.zeros(1234, np.int16), loud, synth.mixture(0.02, 5)[:300]]


def test_normalise_and_trim_bit_exact(engine_sn):
    clips = _clips()
    got = engine_sn.normalise(clips, trim=True)
    for c, g in zip(clips, got):
        ref = O.normalise(c)
        ref = ref[:O.trim_len(len(ref))]
        assert len(g) == len(ref)
        assert np.array_equal(g.view(np.uint32), ref.view(np.uint32))
    got = engine_sn.normalise(clips[:2], trim=False)
    assert [len(g) for g in got] == [len(c) for c in clips[:2]]


def test_stft_framing_and_values(engine_sn):
    clips = _clips()
    lm, ph, fo, peak = engine_sn.stft(clips)
    assert fo.tolist() == np.cumsum([0] + [O.frame_index(len(c)).shape[0] for c in clips]).tolist()   # bit-exact framing
    assert fo[-1] - fo[-2] == 0                                                                       # 300 samples -> no frame
    for u, c in enumerate(clips):
        assert int(peak[u]) == int(O.peak_of(c)) if len(c) else True
        rl, rp = O.logmag_phase(O.normalise(c))
        g, gp = lm[fo[u]:fo[u + 1]], ph[fo[u]:fo[u + 1]]
        if not g.size:
            continue
        assert np.abs(g - rl).max() < LOGMAG_TOL
        mag = np.exp(rl.astype(np.float64))
        perr = np.abs(np.exp(1j * gp.astype(np.float64)) - np.exp(1j * rp.astype(np.float64))) * mag
        assert perr.max() < 1e-3 * max(1.0, mag.max())
    # all-zero clip: log(0 + 1e-5) everywhere (A.1)
    z = lm[fo[3]:fo[4]]
    assert np.allclose(z, np.log(np.float32(1e-5)), atol=1e-6)


def test_istft_matches_oracle_and_roundtrips(engine_sn):
    clips = _clips()[:3]
    lm, ph, fo, peak = engine_sn.stft(clips)
    y, i16, oo = engine_sn.istft(lm, ph, fo, peak=peak, want_i16=True)
    for u, c in enumerate(clips):
        rl, rp = O.logmag_phase(O.normalise(c))
        ry = O.istft(rl, rp)
        gy = y[oo[u]:oo[u + 1]]
        assert len(gy) == len(ry) == O.trim_len(len(c))
        assert np.abs(gy - ry).max() < WAVE_TOL
        ct = c[:len(gy)].astype(np.int32)
        if len(gy) > 900:                                   # interior samples come back exactly as int16
            assert np.abs(i16[oo[u]:oo[u + 1]][400:-400].astype(np.int32) - ct[400:-400]).max() == 0


def test_full_size_roundtrip_256x4s(engine_sn):
    """BASELINE config-2 sized batch (256 x 4 s): STFT -> iSTFT returns the input PCM in the interior."""
    base = [synth.mixture(4.0, u) for u in range(8)]
    clips = [base[u % 8] for u in range(256)]
    lm, ph, fo, peak = engine_sn.stft(clips)
    assert fo[-1] == 256 * 398
    y, i16, oo = engine_sn.istft(lm, ph, fo, peak=peak, want_i16=True)
    assert oo[-1] == 256 * 63920
    for u in (0, 7, 100, 255):
        c = clips[u][:63920].astype(np.int32)
        assert np.abs(i16[oo[u]:oo[u + 1]][400:-400].astype(np.int32) - c[400:-400]).max() == 0
    # every copy of the same clip gives bit-identical spectra wherever it sits in the batch
    assert np.array_equal(lm[fo[0]:fo[1]], lm[fo[248]:fo[249]])
