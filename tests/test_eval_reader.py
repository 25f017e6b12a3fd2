"""Row n4: eval-mode reader (MD5-chosen SNRs, seed pairing, mixing) and the evaluation outputs / loss.
CPU part: host logic against the oracle's restatement of SN/reader.py:183-223, SS/reader.py:120-153.
GPU part: Engine.eval_outputs / read_seeds.get_examples / evaluate against the oracle."""
import hashlib
import os
import pickle

import numpy as np
import pytest

from nhans_b200 import synth, weights as W
from nhans_b200.wavio import write_wav
from oracle import nhans_oracle as O


def _corpus(d, n_speech=2, n_noise=4):
    sp, nz = [], []
    os.makedirs(str(d / "speech"), exist_ok=True)
    os.makedirs(str(d / "noise"), exist_ok=True)
    for i in range(n_speech):
        p = str(d / "speech" / ("s%d.wav" % i))
        write_wav(p, synth.mixture(2.7 + 0.2 * i, 30 + i))
        sp.append(p)
    for i in range(n_noise):
        p = str(d / "noise" / ("n%d.wav" % i))
        write_wav(p, synth.noise_clip(30 + i, "pos" if i % 2 == 0 else "neg"))
        nz.append(p)
    with open(str(d / "speech" / "valid.pkl"), "wb") as f:
        pickle.dump(sp, f)
    with open(str(d / "noise" / "valid.pkl"), "wb") as f:
        pickle.dump(nz, f)
    return sp, nz


def test_md5_snrs():
    from nhans_b200.selective_noise import reader as R
    from nhans_b200.source_separation import reader as S
    seen = set()
    for i in range(64):
        path = "/data/speech/valid/utt_%03d.wav" % i
        h = hashlib.md5(path.encode()).hexdigest()
        want = ([-3, 0, 3, 5, 8][int(h[:8], 16) % 5], [-3, 0, 3, 5, 8][int(h[:6], 16) % 5])
        assert R.eval_snrs(path) == want == R.eval_snrs(path.encode()) == O.eval_snrs(W.SELECTIVE_NOISE, path)
        assert S.eval_snr(path) == [-5, -3, -1, 0, 1, 3, 5][int(h[:8], 16) % 7] == O.eval_snrs(W.SEPARATOR, path)[0]
        seen.add(want)
    assert len(seen) > 8                                              # the choice really varies with the name


def test_seed_pairing_and_combine_signals(tmp_path):
    from nhans_b200.selective_noise import reader as R
    from nhans_b200.source_separation import reader as S
    from nhans_b200.wavio import read_wav
    sp, nz = _corpus(tmp_path)
    for mod in (R, S):
        mod.FLAGS.speech_wav_dir = str(tmp_path / "speech") + "/"
        mod.FLAGS.noise_wav_dir = str(tmp_path / "noise") + "/"
    er = R.read_seeds("valid").preparations()
    assert er.seed_tuples() == [(sp[0], nz[0], nz[1]), (sp[1], nz[2], nz[3])]   # one clean + two consecutive noise seeds
    assert S.read_seeds("test").preparations().seed_tuples() == [(sp[0], nz[0]), (sp[1], nz[1])]
    with pytest.raises(NotImplementedError):
        R.read_seeds("train")
    target, ps, ns, mixed, snr_p, snr_n = R.combine_signals(False, sp[1].encode(), nz[2].encode(), nz[3].encode())
    snrs = O.eval_snrs(W.SELECTIVE_NOISE, sp[1])
    clean = O.normalise(read_wav(sp[1]))
    clean = clean[:len(clean) - (len(clean) - 400) % 160]
    om, oa, ob, ot = O.domixing_sn(clean, O.normalise(read_wav(nz[2])), O.normalise(read_wav(nz[3])), snrs[0], snrs[1], with_target=True)
    assert (int(snr_p), int(snr_n)) == snrs
    assert np.array_equal(mixed, om) and np.array_equal(ps, oa) and np.array_equal(ns, ob) and np.array_equal(target, ot)


def test_oracle_eval_loss_definition(weights_sn):
    """The per-example loss is mean_k((denoised - target)^2 * linspace(2, 1, 201)) (SN/main.py:243-246)."""
    rng = np.random.default_rng(0)
    d, t = rng.standard_normal((7, 201)).astype(np.float32), rng.standard_normal((7, 201)).astype(np.float32)
    w = np.linspace(2, 1, 201, dtype=np.float32)
    want = ((d.astype(np.float64) - t) ** 2 * w).mean(axis=1)
    got = np.mean(np.square(d - t) * w.reshape(1, -1), axis=1, dtype=np.float32)
    assert np.allclose(got, want, rtol=1e-5)
    assert abs(w[0] - 2) < 1e-7 and abs(w[200] - 1) < 1e-7 and abs(w[100] - 1.5) < 1e-6


@pytest.mark.gpu
def test_eval_loss_kernel(engine_sn):
    rng = np.random.default_rng(1)
    for n in (1, 37, 1000):
        d, t = rng.standard_normal((n, 201)).astype(np.float32), rng.standard_normal((n, 201)).astype(np.float32)
        want = ((d.astype(np.float64) - t) ** 2 * np.linspace(2, 1, 201)).mean(axis=1)
        assert np.allclose(engine_sn.eval_loss(d, t), want, rtol=2e-5)
    assert engine_sn.eval_loss(np.zeros((0, 201), np.float32), np.zeros((0, 201), np.float32)).shape == (0,)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["sn", "ss"])
def test_eval_outputs_vs_oracle(tag, tmp_path, engine_sn, engine_ss, oracle_sn, oracle_ss):
    from nhans_b200.wavio import read_wav
    sp, nz = _corpus(tmp_path)
    if tag == "sn":
        from nhans_b200.selective_noise import reader as R
        got = R.model_outputs(engine_sn, sp[0], nz[0], nz[1])
        ref = O.eval_outputs_arrays(oracle_sn, W.SELECTIVE_NOISE, sp[0], read_wav(sp[0]), read_wav(nz[0]), read_wav(nz[1]))
        keys = ["mixed", "target", "pos", "neg"]
        assert got["snr_pos"][0] == ref["snr_pos"] and got["snr_neg"][0] == ref["snr_neg"]
    else:
        from nhans_b200.source_separation import reader as R
        got = R.model_outputs(engine_ss, sp[0], sp[1])
        ref = O.eval_outputs_arrays(oracle_ss, W.SEPARATOR, sp[0], read_wav(sp[0]), read_wav(sp[1]))
        keys = ["mixed", "clean"]
        assert got["snr"][0] == ref["snr"]
    n = len(ref["location"])
    assert np.array_equal(got["location"], ref["location"]) and len(got["loss"]) == n
    for k in keys:
        assert np.abs(got[k] - ref[k]).max() < 1e-3, k
    rel = np.linalg.norm(np.exp(got["denoised"]) - np.exp(ref["denoised"])) / np.linalg.norm(np.exp(ref["denoised"]))
    assert rel < 1e-3
    assert np.allclose(got["loss"], ref["loss"], rtol=2e-2, atol=1e-4)
    assert abs(got["loss"].mean() / ref["loss"].mean() - 1) < 5e-3


@pytest.mark.gpu
def test_eval_before_training_files(tmp_path, engine_sn):
    from nhans_b200.selective_noise import main as M, reader as R
    sp, nz = _corpus(tmp_path)
    R.FLAGS.speech_wav_dir = str(tmp_path / "speech") + "/"
    R.FLAGS.noise_wav_dir = str(tmp_path / "noise") + "/"
    M.FLAGS.wav_dump_folder = str(tmp_path / "dump")
    M.FLAGS.dump_results = str(tmp_path / "npy")
    losses = M.eval_before_training(("valid",), engine=engine_sn)
    assert np.isfinite(losses["valid"]) and losses["valid"] > 0
    wavs = sorted(os.listdir(str(tmp_path / "dump")))
    assert len(wavs) == 2 * 5 and sum(w.endswith("_denoised.wav") for w in wavs) == 2
    arr = np.load(str(tmp_path / "npy" / "nhans_b200_valid_0_location.npy"))
    assert (arr == 0).sum() == 2                                      # two utterances, each restarting at location 0
