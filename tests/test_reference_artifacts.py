"""Pins against artefacts the REFERENCE itself produced.  DEMO_N-HANS/ holds the wav sets written by the reference's
evaluate() (N_HANS___Selective_Noise/main.py:266-353): mixed / target / posNoise / negNoise of validation
utterances mixed by reader.combine_signals -> domixing.  Excerpts + whole-file statistics are committed as
tests/golden/demo_relations.npz (made by tests/golden/make_demo_fixtures.py).  They pin two things the new
restatements (nhans_b200/selective_noise/apply.py::domixing, oracle.nhans_oracle.domixing_sn) must reproduce:

  * the scaling quirk of SN/reader.py:172-178: `mixed` is divided by the peak of the raw mixture, but target and
    the two noise signals are divided by the peak of the *already normalised* mixture (~1), so
    target + negNoise == r * mixed with r = that raw peak, not 1;
  * the SNR convention K = sqrt(Ps / Pn * 10^(-snr/10)) (SN/reader.py:160-170): speech-to-noise power ratios of the
    reference's files follow the SNR labels in their names."""
import os

import numpy as np
import pytest

from nhans_b200 import synth
from oracle import nhans_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "demo_relations.npz"))
N_SETS = len(G["names"])


def test_fixture_inventory():
    assert N_SETS == 13
    assert sum("%d_posNoise" % i in G for i in range(N_SETS)) == 6       # 6 selective-noise sets, 7 denoising sets


@pytest.mark.parametrize("i", range(N_SETS))
def test_reference_files_obey_the_domixing_scaling(i):
    m = G["%d_mixed" % i].astype(np.float64)
    s = G["%d_target" % i].astype(np.float64) + G["%d_negNoise" % i].astype(np.float64)
    r = float(G["%d_ratio" % i])
    assert float(G["%d_resid" % i]) < 2e-5                                # whole file: target + neg == r * mixed
    assert np.abs(s - r * m).max() < 2e-5                                 # and sample by sample on the excerpt (float32 files)
    assert r > 0.9


def test_the_quirk_is_real():
    """If target / noises were normalised like `mixed`, every ratio would be 1."""
    r = np.array([float(G["%d_ratio" % i]) for i in range(N_SETS)])
    assert np.sum(np.abs(r - 1) > 0.02) >= 10 and r.max() > 1.3


def _ratio(mixed, target, neg):
    m, s = mixed.astype(np.float64), target.astype(np.float64) + neg.astype(np.float64)
    r = np.dot(s, m) / np.dot(m, m)
    return r, np.linalg.norm(s - r * m) / np.linalg.norm(s)


def test_restatements_reproduce_the_scaling():
    from nhans_b200.selective_noise import apply as A
    clean = O.normalise(synth.mixture(2.0, 11))
    clean = clean[:len(clean) - (len(clean) - 400) % 160]
    pos, neg = O.normalise(synth.noise_clip(11, "pos")), O.normalise(synth.noise_clip(11, "neg"))
    for snr_p, snr_n in ((3, 8), (-3, 0), (5, 3)):
        raw_peak = None
        mixed, target, kp, kn, ps, ns = A.domixing(clean, pos, neg, snr_p, snr_n)
        raw = clean + np.float32(kp) * pos[:len(clean)] + np.float32(kn) * neg[:len(clean)]
        raw_peak = float(np.abs(raw).max())
        r, res = _ratio(mixed, target, ns)
        assert res < 2e-6 and abs(r / raw_peak - 1) < 1e-5               # r is the peak of the raw mixture, like the files
        om, ops, ons, ot = O.domixing_sn(clean, pos, neg, snr_p, snr_n, with_target=True)
        r2, res2 = _ratio(om, ot, ons)
        assert res2 < 2e-6 and abs(r2 / raw_peak - 1) < 1e-5
        # SNR convention: speech / scaled-noise power = 10^(snr / 10) over the mixed span
        sig = (target - ps).astype(np.float64)
        assert abs(10 * np.log10(np.mean(sig ** 2) / np.mean(ps.astype(np.float64) ** 2)) - snr_p) < 0.05
        assert abs(10 * np.log10(np.mean(sig ** 2) / np.mean(ns.astype(np.float64) ** 2)) - snr_n) < 0.05


def test_reference_files_follow_the_snr_labels():
    """The files cover frames [200:] only while the powers were taken over the whole utterance (noise repeated to
    length; real, non-stationary noise), so the match is statistical: correlated with the labels and close on average."""
    lab, est = [], []
    for i in range(N_SETS):
        sp, sn = G["%d_snr_labels" % i]
        if "%d_snr_pos_est" % i in G:
            lab.append(sp); est.append(float(G["%d_snr_pos_est" % i]))
        lab.append(sn); est.append(float(G["%d_snr_neg_est" % i]))
    lab, est = np.array(lab, np.float64), np.array(est)
    assert np.corrcoef(lab, est)[0, 1] > 0.7
    assert np.mean(np.abs(est - lab)) < 2.5 and abs(np.mean(est - lab)) < 2.0
    # the opposite sign convention (noise-to-speech) would put the estimates at -label
    assert np.mean(np.abs(est - lab)) < 0.5 * np.mean(np.abs(est + lab))


# The reference's own apply outputs (N_HANS___Selective_Noise/audio_examples): input length -> output length.
# exp1_noisy.wav has 63520 samples and exp1_denoised.wav 63440; exp2: 49600 -> 49520 (float32, 16 kHz, peak 2.28 /
# 1.17, i.e. unclipped).  handle_signals drops (N - 400) % 160 trailing samples (SN/apply.py:157-161), the
# inverse STFT returns (T - 1) * 160 + 400 samples.
REFERENCE_LENGTHS = [(63520, 63440), (49600, 49520)]


@pytest.mark.parametrize("n_in,n_out", REFERENCE_LENGTHS)
def test_output_length_matches_the_reference_files(n_in, n_out):
    assert O.trim_len(n_in) == n_out
    T = 1 + (O.trim_len(n_in) - 400) // 160
    assert (T - 1) * 160 + 400 == n_out
    x = np.zeros(n_in, np.int16)
    lm, ph = O.logmag_phase(O.normalise(x)[:O.trim_len(n_in)])
    assert lm.shape == (T, 201) and len(O.istft(lm, ph)) == n_out


def test_library_output_offsets_match_the_reference_files():
    """nhans_output_offsets is host arithmetic: callable without a GPU."""
    import ctypes
    from nhans_b200 import _lib
    lib = _lib.load()
    offs = np.array([0, 63520, 63520 + 49600, 63520 + 49600 + 399], np.int64)
    out = np.zeros(4, np.int64)
    rc = lib.nhans_output_offsets(offs.ctypes.data_as(ctypes.c_void_p), 3, out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0 and list(out) == [0, 63440, 63440 + 49520, 63440 + 49520]     # < 400 samples -> no frames
