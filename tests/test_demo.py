"""Row n3 (apply_demo): on-the-fly 0 dB mixing on the host, float-input STFT, processing from frame 200 on.
CPU part: the host mixing of the apply.py mirrors against the oracle's restatement of SN/apply.py:56-139 and
SS/apply.py:55-108 (bit-exact) and its defining properties.  GPU part: enhance_demo against the oracle."""
import numpy as np
import pytest

from nhans_b200 import synth, weights as W
from oracle import nhans_oracle as O


def _f32(pcm):
    return O.normalise(pcm)


def test_domixing_sn_matches_oracle_bit_exact():
    from nhans_b200.selective_noise import apply as A
    clean = _f32(synth.mixture(2.5, 0))
    clean = clean[:len(clean) - (len(clean) - 400) % 160]
    pos, neg = _f32(synth.noise_clip(0, "pos", 1.0)), _f32(synth.noise_clip(1, "neg", 4.0))   # shorter and longer
    mixed, target, kp, kn, ps, ns = A.domixing(clean, pos, neg, 0, 0)
    om, ops, ons = O.domixing_sn(clean, pos, neg)
    assert mixed.dtype == np.float32 and ps.dtype == np.float32
    assert np.array_equal(mixed, om) and np.array_equal(ps, ops) and np.array_equal(ns, ons)
    assert len(mixed) == len(clean) == len(ps) == len(ns)
    # 0 dB: the scaled noises carry the speech power; the mixture is peak normalised
    p = lambda x: np.mean(np.square(x.astype(np.float64)))
    assert abs(p(ps) / p(ns) - 1) < 1e-5
    assert abs(np.abs(mixed).max() - 1) < 1e-5
    # the short noise was repeated (SN/apply.py:60-62)
    n1 = len(pos)
    assert np.allclose(ps[n1:2 * n1], ps[:n1], rtol=0, atol=0)
    # reference quirk (SN/apply.py:99-102): target and the noise signals are divided by the peak of the already
    # normalised mixture (~1), i.e. they stay on the pre-normalisation scale
    assert np.abs(target - (clean + ps)).max() < 1e-5


def test_domixing_snr_argument():
    from nhans_b200.selective_noise import apply as A
    clean = _f32(synth.mixture(1.0, 3))
    pos, neg = _f32(synth.noise_clip(3, "pos")), _f32(synth.noise_clip(3, "neg"))
    _, _, kp0, kn0, _, _ = A.domixing(clean, pos, neg, 0, 0)
    _, _, kp1, kn1, _, _ = A.domixing(clean, pos, neg, 3, -3)
    assert abs(kp1 / kp0 - 10 ** (-3 / 20.0)) < 1e-9 and abs(kn1 / kn0 - 10 ** (3 / 20.0)) < 1e-9
    z = np.zeros_like(pos)
    _, _, kz, _, zs, _ = A.domixing(clean, z, neg, 0, 0)           # silent noise: K = 1 (SN/apply.py:82-83)
    assert kz == 1.0 and not zs.any()


def test_domixing_ss_matches_oracle_bit_exact():
    from nhans_b200.source_separation import apply as A
    clean = _f32(synth.speaker_clip(0, "target", 2.0))
    noise = _f32(synth.speaker_clip(0, "interference", 0.7))
    mixed, k = A.domixing(clean, noise, 0)
    om, ok = O.domixing_ss(clean, noise)
    assert np.array_equal(mixed, om) and k == ok


def test_combine_signals_files(tmp_path):
    from nhans_b200.selective_noise import apply as A
    from nhans_b200.source_separation import apply as S
    from nhans_b200.wavio import write_wav
    sp, po, ne = synth.mixture(2.6, 5), synth.noise_clip(5, "pos"), synth.noise_clip(5, "neg")
    for n, x in (("s", sp), ("p", po), ("n", ne)):
        write_wav(str(tmp_path / (n + ".wav")), x)
    ps, ns, mixed, snr_p, snr_n = A.combine_signals(str(tmp_path / "s.wav"), str(tmp_path / "p.wav"), str(tmp_path / "n.wav"))
    om, oa, ob = O.demo_signals(W.SELECTIVE_NOISE, sp, po, ne)
    assert np.array_equal(mixed, om) and np.array_equal(ps, oa) and np.array_equal(ns, ob)
    assert (len(mixed) - 400) % 160 == 0 and snr_p == 0 and snr_n == 0
    clean, nk, mixed, snr = S.combine_signals(str(tmp_path / "s.wav"), str(tmp_path / "p.wav"))
    om, oa, ob = O.demo_signals(W.SEPARATOR, sp, po)
    assert np.array_equal(mixed, om) and np.array_equal(nk, oa) and np.array_equal(clean, ob)


def _snr(ref, got):
    ref = ref.astype(np.float64)
    err = got.astype(np.float64) - ref
    return 10 * np.log10(np.sum(ref ** 2) / (np.sum(err ** 2) + 1e-30))


@pytest.mark.gpu
def test_stft_f32_vs_oracle(engine_sn):
    sigs = [_f32(synth.mixture(1.0, 1)) * np.float32(0.37), _f32(synth.noise_clip(2, "pos", 0.5))]
    lm, ph, fo = engine_sn.stft_f32(sigs)
    assert list(fo) == [0, 98, 98 + 48]
    for u, x in enumerate(sigs):
        olm, oph = O.logmag_phase(x)
        assert np.abs(lm[fo[u]:fo[u + 1]] - olm).max() < 1e-3           # |d log-magnitude| = relative magnitude error
    # same as the int16 entry point when the floats are the normalised PCM (the int16 path multiplies by the float64
    # reciprocal of the peak, which can differ from the exact quotient in the last float32 bit once in ~1e9 samples)
    pcm = synth.mixture(0.8, 4)
    a = engine_sn.stft([pcm])
    b = engine_sn.stft_f32([O.normalise(pcm)])
    assert np.abs(a[0] - b[0]).max() < 1e-5 and np.mean(a[0] != b[0]) < 1e-4
    d = np.abs(a[1] - b[1])
    assert np.minimum(d, 2 * np.pi - d).max() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["sn", "ss"])
def test_enhance_demo_vs_oracle(tag, engine_sn, engine_ss, oracle_sn, oracle_ss):
    if tag == "sn":
        eng, net = engine_sn, oracle_sn
        sigs = O.demo_signals(W.SELECTIVE_NOISE, synth.mixture(3.0, 7), synth.noise_clip(7, "pos"), synth.noise_clip(7, "neg"))
    else:
        eng, net = engine_ss, oracle_ss
        sigs = O.demo_signals(W.SEPARATOR, synth.speaker_clip(7, "target", 3.0), synth.speaker_clip(7, "interference", 2.5))
    y, ymix = eng.enhance_demo(*sigs)
    ry, rmix = O.apply_demo_arrays(net, *sigs)
    T = 1 + (len(sigs[0]) - 400) // 160
    assert len(y) == len(ry) == (T - 200 - 1) * 160 + 400
    assert _snr(ry, y) >= 40.0
    assert np.abs(ymix - rmix).max() < 1e-5


@pytest.mark.gpu
def test_apply_demo_files(tmp_path, monkeypatch):
    from nhans_b200 import session
    from nhans_b200.selective_noise import apply as A
    from nhans_b200.engine import NhansError
    from scipy.io import wavfile
    from nhans_b200.wavio import write_wav
    read_wav = lambda p: wavfile.read(p)[1]                              # float32 files, like the reference writes
    monkeypatch.setenv("NHANS_WIN_CAPACITY", "128")
    monkeypatch.setenv("NHANS_ROW_CAPACITY", "2")
    monkeypatch.setenv("NHANS_MODEL_DIR", str(tmp_path / "no_model"))
    monkeypatch.setenv("NHANS_ALLOW_RANDOM_INIT", "1")            # no checkpoint on the test box: seeded random init
    session.close_all()
    for n, x in (("s", synth.mixture(2.6, 5)), ("p", synth.noise_clip(5, "pos")), ("n", synth.noise_clip(5, "neg")),
                 ("short", synth.mixture(1.5, 6))):
        write_wav(str(tmp_path / (n + ".wav")), x)
    save_to = str(tmp_path / "output_demo.wav")
    y, ymix = A.apply_demo(str(tmp_path / "s.wav"), str(tmp_path / "p.wav"), str(tmp_path / "n.wav"), save_to)
    out = read_wav(save_to)
    mixed_demo = read_wav(str(tmp_path / "mixed_demo.wav"))            # save_to[:-15] + 'mixed_demo.wav'
    assert out.dtype == np.float32 and np.array_equal(out, y) and np.array_equal(mixed_demo, ymix)
    with pytest.raises(NhansError):                                     # mixture must be longer than 200 frames
        A.apply_demo(str(tmp_path / "short.wav"), str(tmp_path / "p.wav"), str(tmp_path / "n.wav"), save_to)
    session.close_all()
