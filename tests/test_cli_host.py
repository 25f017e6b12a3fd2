"""Host logic of the CLI mirrors that needs no GPU: folder pairing with a partially populated --pos folder, stereo
handling, checkpoint policy."""
import os

import numpy as np
import pytest

from nhans_b200 import synth, weights as W
from nhans_b200.selective_noise import apply as sn_apply
from nhans_b200.wavio import is_pcm16, normalise_host, read_wav, write_wav


class _FakeEngine:
    """Records what the apply mirror feeds the engine (no GPU in the CPU suite)."""

    def __init__(self):
        self.calls = []

    def enhance(self, mixes, poss, negs, **kw):
        self.calls.append(("enhance", mixes, poss, negs))
        oo = np.cumsum([0] + [max(0, len(m) - (len(m) - 400) % 160) for m in mixes])
        return {"out_offs": oo, "f32": [np.zeros(oo[i + 1] - oo[i], np.float32) for i in range(len(mixes))]}

    def postmix(self, out_offs, compensate=0.0, ac=False):
        n = [int(out_offs[i + 1] - out_offs[i]) for i in range(len(out_offs) - 1)]
        z = [np.zeros(k, np.float32) for k in n]
        return {"mixed_processed": z, "removed": z, "compensated": z, "snr_est": np.ones(len(n), np.float32)}

    def enhance_float(self, mix, a, b):
        self.calls.append(("enhance_float", mix, a, b))
        z = np.zeros(400, np.float32)
        return dict(f32=z, mixed_processed=z, peak=1.0)


def test_partial_pos_folder_and_silent_entries(tmp_path, monkeypatch):
    d = tmp_path
    for sub in ("in", "neg", "pos"):
        os.makedirs(str(d / sub))
    for n in ("a.wav", "b.wav", "c.wav"):
        write_wav(str(d / "in" / n), synth.mixture(0.2, 1))
        write_wav(str(d / "neg" / n), synth.noise_clip(1))
    write_wav(str(d / "pos" / "b.wav"), synth.noise_clip(1, "pos"))          # only ONE of the three inputs has a --pos file
    mixed, poss, negs, outs = sn_apply._pairs(str(d / "in"), str(d / "pos"), str(d / "neg"), str(d / "out"))
    assert [os.path.basename(p) for p in mixed] == ["a.wav", "b.wav", "c.wav"]
    assert poss[0] is None and poss[2] is None and poss[1].endswith("b.wav")
    fake = _FakeEngine()
    monkeypatch.setattr(sn_apply, "get_engine", lambda v: fake)
    sn_apply.apply_snc_batch(mixed, poss, negs, outs)                        # used to crash in read_wav(None)
    kind, mixes, pos_clips, neg_clips = fake.calls[0]
    assert kind == "enhance" and len(mixes) == 3 and len(pos_clips) == 3
    assert not pos_clips[0].any() and not pos_clips[2].any() and len(pos_clips[0]) >= 32240   # digital silence, >= 200 frames
    assert np.array_equal(pos_clips[1], synth.noise_clip(1, "pos"))
    # a file literally called Silent.wav is never read from disk
    fake.calls.clear()
    sn_apply.apply_snc_batch(mixed[:2], [str(d / "nowhere" / "Silent.wav"), poss[1]], negs[:2], outs[:2])
    assert not fake.calls[0][2][0].any()
    # all entries silent: the engine's cached Silent embedding (pos = None)
    fake.calls.clear()
    sn_apply.apply_snc_batch(mixed[:1], [None], negs[:1], outs[:1])
    assert fake.calls[0][2] is None
    for n in ("a", "b", "c"):
        assert os.path.exists(str(d / "out" / (n + "_denoised.wav")))


def test_stereo_takes_the_float_path_exactly(tmp_path, monkeypatch):
    """SN/apply.py:46-53: stereo files are averaged in float64; the mean is NOT rounded back to int16."""
    rng = np.random.default_rng(0)
    st = rng.integers(-20000, 20000, (4000, 2)).astype(np.int16)
    st[7] = (3, 4)                                                           # a half-integer mean
    write_wav(str(tmp_path / "st.wav"), st)
    x = read_wav(str(tmp_path / "st.wav"))
    assert not is_pcm16(x) and x.dtype == np.float64 and x[7] == 3.5
    ref = (st.mean(axis=1) / (max(abs(st.mean(axis=1))) + 0.000001)).astype(np.float32)       # the reference's arithmetic
    assert np.array_equal(normalise_host(x), ref)
    mono = synth.mixture(0.25, 2)
    mono[5] = -32768                                                         # int16 abs wrap, like numpy in the reference
    assert np.array_equal(normalise_host(mono), (mono / (float(max(abs(mono))) + 0.000001)).astype(np.float32))
    write_wav(str(tmp_path / "neg.wav"), synth.noise_clip(0))
    fake = _FakeEngine()
    monkeypatch.setattr(sn_apply, "get_engine", lambda v: fake)
    sn_apply.apply_snc_batch([str(tmp_path / "st.wav")], [None], [str(tmp_path / "neg.wav")], [str(tmp_path / "o_denoised.wav")])
    assert fake.calls[0][0] == "enhance_float" and fake.calls[0][1].dtype == np.float64


def _write_full_checkpoint(prefix, tensors):
    from test_weights import _write_bundle
    _write_bundle(prefix, tensors)


def test_checkpoint_policy(tmp_path, monkeypatch):
    """Missing checkpoint -> error unless random init is allowed; truncated shard -> always an error; a complete
    checkpoint in the reference's tensor-bundle format is restored bit for bit."""
    monkeypatch.delenv("NHANS_ALLOW_RANDOM_INIT", raising=False)
    with pytest.raises(W.CheckpointMissing):
        W.load_or_init(0, str(tmp_path / "nothing"))
    monkeypatch.setenv("NHANS_ALLOW_RANDOM_INIT", "1")
    assert W.load_or_init(0, str(tmp_path / "nothing"))[1] == "random-init"
    w = W.seeded_init(0, 5)
    d = tmp_path / "trained_model"
    os.makedirs(str(d))
    _write_full_checkpoint(str(d / "81448_0-1000000"), w)
    got, src = W.load_or_init(0, str(d), allow_random=False)
    assert src == "checkpoint" and set(got) == set(w)
    assert all(np.array_equal(got[k], w[k]) for k in w)
    shard = str(d / "81448_0-1000000.data-00000-of-00001")
    with open(shard, "r+b") as f:
        f.truncate(os.path.getsize(shard) // 2)
    with pytest.raises(ValueError):
        W.load_or_init(0, str(d), allow_random=True)                          # truncated: never silently random
    with open(shard, "wb") as f:
        f.write(b"version https://git-lfs.github.com/spec/v1\noid sha256:0\nsize 115999524\n")
    with pytest.raises(W.CheckpointMissing):
        W.load_or_init(0, str(d), allow_random=False)                         # LFS pointer = absent
