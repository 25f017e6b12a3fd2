"""Pins of the CPU oracle that need no weights and no GPU (SURVEY.md §4 identities) + golden vectors."""
import os

import numpy as np
import pytest
import torch

from nhans_b200 import synth, weights as W
from oracle import nhans_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_trim_and_frame_count():
    for n, t in ((64000, 398), (128000, 798), (160000, 998)):       # 4 / 8 / 10 s
        nt = O.trim_len(n)
        assert (nt - 400) % 160 == 0 and n - nt < 160
        assert O.frame_index(nt).shape[0] == t
    assert O.frame_index(399).shape[0] == 0
    idx = O.frame_index(1000)
    assert idx[0, 0] == 0 and idx[1, 0] == 160 and idx[-1, -1] == 160 * 3 + 399


def test_normalise_matches_reference_expression():
    x = synth.mixture(0.1, 3)
    ref = (x / (max(abs(x)) + 0.000001)).astype(np.float32)          # SN/apply.py:150,153 verbatim
    assert np.array_equal(O.normalise(x), ref)
    assert np.abs(O.normalise(np.zeros(500, np.int16))).max() == 0


def test_stft_definition_against_naive_dft():
    x = O.normalise(synth.mixture(0.05, 1))
    X = O.stft(x)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 400)
    t = 2
    naive = np.array([np.sum(x[160 * t:160 * t + 400] * w * np.exp(-2j * np.pi * k * np.arange(400) / 400)) for k in (0, 1, 57, 200)])
    assert np.allclose(X[t, [0, 1, 57, 200]], naive, atol=1e-5)
    assert X.shape[1] == 201


def test_istft_of_stft_is_identity_in_the_interior():
    x = O.normalise(synth.mixture(1.0, 2))
    x = x[:O.trim_len(len(x))]
    lm, ph = O.logmag_phase(x)
    y = O.istft(lm, ph)
    assert len(y) == len(x)
    assert np.abs(y[400:-400] - x[400:-400]).max() < 2e-5             # the +1e-5 floor costs ~1e-5
    assert np.abs(y[:400] - x[:400]).max() > 1e-3                     # edges are attenuated (A.7), reference behaviour


def test_inverse_window_is_the_tf_formula():
    w = O.hann_periodic().astype(np.float64)
    d = np.array([sum(w[m + 160 * j] ** 2 for j in range(3) if m + 160 * j < 400) for m in range(160)])
    assert d.min() > 0.85 and d.max() < 1.02
    assert np.allclose(O.inverse_window(), (w / np.tile(d, 3)[:400]).astype(np.float32))


def test_windows_are_frame_shifts_with_zero_padding():
    T = 50
    spec = np.arange(T * 201, dtype=np.float32).reshape(T, 201) + 1
    win = O.strided_crop(spec, 35, 1)
    assert win.shape == (T, 35, 201)                                   # one window per frame
    for i in (0, 5, 17, 30, 49):
        for r in (0, 16, 17, 18, 34):
            f = i - 17 + r
            exp = spec[f] if 0 <= f < T else np.zeros(201, np.float32)
            assert np.array_equal(win[i, r], exp)
    assert np.array_equal(O.strided_crop(spec, 1, 1)[:, 0], spec)     # phase windows: no padding


def test_same_padding_rule_against_torch():
    # asymmetric TF 'SAME' (k=4: before 1 / after 2) vs an explicit gather
    x = torch.arange(35 * 201, dtype=torch.float32).reshape(1, 1, 35, 201)
    for n, k, s, exp in ((35, 4, 1, (1, 2)), (35, 4, 2, (1, 2)), (18, 3, 2, (0, 1)), (9, 3, 2, (1, 1)), (200, 8, 3, (3, 3)), (67, 8, 1, (3, 4))):
        assert O.same_pads(n, k, s) == exp
    net = O.Net({"c/w": np.ones((4, 4, 1, 1), np.float32)})
    y = net.conv(x, "c", (2, 2), False)
    assert y.shape[2:] == (18, 101)
    assert float(y[0, 0, 0, 0]) == float(x[0, 0, 0:3, 0:3].sum())     # top-left window sees 1 pad row/col
    assert float(y[0, 0, 17, 100]) == float(x[0, 0, 33:35, 199:201].sum())


def test_zero_last_dense_gives_mixed_processed(weights_sn):
    w = dict(weights_sn)
    w["last_dense/w"] = np.zeros_like(w["last_dense/w"])
    w["last_dense/b"] = np.zeros_like(w["last_dense/b"])
    net = O.Net(w)
    r = O.apply_arrays(net, synth.mixture(0.06, 4), synth.silence(), synth.noise_clip(4), return_all=True)
    assert np.array_equal(r["denoised"], r["logmag"])
    assert np.array_equal(r["samples"], r["mixed_processed"])


def test_faithful_equals_dedup(oracle_sn):
    mix, pos, neg = synth.mixture(0.045, 5), synth.noise_clip(5, "pos"), synth.noise_clip(5, "neg")
    a = O.apply_arrays(oracle_sn, mix, pos, neg, faithful=True)       # towers per window, mb = 100
    b = O.apply_arrays(oracle_sn, mix, pos, neg, faithful=False)
    assert np.abs(a - b).max() < 1e-4


def test_context_too_short_is_an_error():
    with pytest.raises(ValueError):
        O.context_of(O.logmag_phase(O.normalise(synth.noise_clip(0, seconds=1.0)))[0])


@pytest.mark.parametrize("tag,variant", [("sn", 0), ("ss", 1)])
def test_oracle_reproduces_golden(tag, variant, weights_sn, weights_ss):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(os.path.join(GOLD, "golden_%s.npz" % tag))
    net = O.Net(weights_sn if variant == 0 else weights_ss, variant)
    mix, a, b = mg.inputs(tag)
    r = O.apply_arrays(net, mix, a, b, return_all=True)
    assert np.abs(r["logmag"] - g["logmag"]).max() < 1e-5
    assert np.abs(r["denoised"] - g["denoised"]).max() < 5e-4          # fp32 conv reduction order may differ per CPU
    assert np.abs(r["samples"] - g["samples"]).max() < 5e-3
