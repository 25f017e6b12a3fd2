"""CPU ORACLE — test infrastructure only (never imported by the product path).

A restatement, in numpy / torch-CPU, of the reference's inference path:

  * wav normalise/trim ........ N_HANS___Selective_Noise/apply.py:142-163 (SS/apply.py:111-136)
  * STFT, log-mag, phase ....... SN/apply.py:368-375 (tf.signal.stft, periodic Hann, no end pad)
  * window extraction .......... SN/apply.py:170-186, 378, 389
  * context slice .............. SN/apply.py:381-387
  * network (eval mode) ........ SN/main.py:98-242 + SN/blocks.py:23-48, 64-69, 104-108
  * reconstruction ............. SN/apply.py:189-204 (tf.signal.inverse_stft + inverse_stft_window_fn)
  * mini-batch loop ............ SN/apply.py:440-454 (mb = 100)
  * post-mix outputs ........... SN/apply.py:456-472
  * apply_demo / eval reader ... SN/apply.py:56-139, 212-337; SN/reader.py:131-223, 398-420; SN/main.py:243-246

PARITY UNPINNED for the network arithmetic: the reference cannot run here (TensorFlow is not
installed, the trained blobs are git-LFS pointers, the repo ships no tests or golden vectors —
SURVEY.md F4/F5, §8c).  What *is* pinned against reference-produced artefacts: the variable
inventory against the reference's own checkpoint ``.index`` files (tests/golden/ckpt_index_*.json);
the mixing arithmetic (domixing: the target / noise scaling quirk and the SNR convention) against the
13 wav sets the reference's evaluate() wrote under DEMO_N-HANS/ (tests/golden/demo_relations.npz);
the trim / output-length rule against audio_examples/exp{1,2}_{noisy,denoised}.wav.  Pinned against
definitions only: the DSP against numpy.fft and the STFT->iSTFT identity, the conv/padding rule
against torch's conv2d, the window/frame indexing identities of SURVEY.md §4.

Two modes of ``apply_snc``: ``faithful=True`` follows the reference structure exactly (windows
and tiled contexts materialised, both embedding towers evaluated for every window of every
mini-batch of 100) and is the timed "reference CPU path"; ``faithful=False`` evaluates each tower
once per context clip (identical result in eval mode).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from nhans_b200 import weights as W

FS = 16000
WIN = 400            # int(Fs * 0.025)
HOP = 160            # int(Fs * 0.010)
NBIN = 201
MIX_WIN = 35         # SN/apply.py:38
NOISE_WIN = 200      # SN/apply.py:37
BN_EPS = 0.001       # SN/blocks.py:108
LOG_EPS = 1e-5       # SN/apply.py:373


# ---------------------------------------------------------------------------------------------
# A.1 load / normalise / trim
# ---------------------------------------------------------------------------------------------
def normalise(pcm):
    """x / (max|x| + 1e-6) in float64 then float32 (SN/apply.py:150-155).  ``pcm`` is the int16
    (or float64 stereo-mean) array ``read_wav`` returns (SN/apply.py:46-53).  numpy's
    abs(int16(-32768)) wraps to -32768; the reference inherits that, so do we."""
    pcm = np.asarray(pcm)
    if pcm.size == 0:
        return pcm.astype(np.float32)
    peak = max(abs(pcm))
    return (pcm / (peak + 0.000001)).astype(np.float32)


def trim_len(n):
    """Whole number of frames (SN/apply.py:158-161)."""
    if n < WIN:
        return n
    r = (n - WIN) % HOP
    return n - r


def peak_of(pcm):
    pcm = np.asarray(pcm)
    return float(max(abs(pcm))) if pcm.size else 0.0


# ---------------------------------------------------------------------------------------------
# A.2 / A.3 STFT front end
# ---------------------------------------------------------------------------------------------
def hann_periodic(n=WIN):
    """tf.signal.hann_window(periodic=True): 0.5 - 0.5 cos(2 pi k / n), float32."""
    k = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)).astype(np.float32)


def frame_index(n_samples):
    """Frame t covers samples [160 t, 160 t + 400) (tf.signal.frame, pad_end=False)."""
    t = 1 + (n_samples - WIN) // HOP if n_samples >= WIN else 0
    return np.arange(t)[:, None] * HOP + np.arange(WIN)[None, :]


def stft(x):
    """tf.signal.stft(x, 400, 160, fft_length=400) -> complex [T, 201] (float64 arithmetic on the
    float32 samples and float32 window, so rounding noise is below the compared tolerance)."""
    x = np.asarray(x, dtype=np.float32)
    idx = frame_index(len(x))
    if idx.shape[0] == 0:
        return np.zeros((0, NBIN), dtype=np.complex128)
    frames = x[idx].astype(np.float64) * hann_periodic().astype(np.float64)[None, :]
    return np.fft.rfft(frames, n=WIN, axis=1)


def logmag_phase(x):
    """log(|X| + 1e-5), angle(X) as float32 (SN/apply.py:373-375)."""
    X = stft(x)
    return (np.log(np.abs(X) + LOG_EPS).astype(np.float32), np.angle(X).astype(np.float32))


# ---------------------------------------------------------------------------------------------
# A.4 windows, A.5 contexts
# ---------------------------------------------------------------------------------------------
def strided_crop(spec, length, stride=1):
    """pad_1D_for_windowing + extract_image_patches (SN/apply.py:170-186): zero rows
    (length+1)//2-1 before and length//2 after, then all stride-1 windows of ``length`` rows."""
    before = (length + 1) // 2 - 1
    after = length // 2
    padded = np.concatenate([np.zeros((before, spec.shape[1]), spec.dtype), spec,
                             np.zeros((after, spec.shape[1]), spec.dtype)], axis=0)
    n = (padded.shape[0] - length) // stride + 1
    idx = np.arange(n)[:, None] * stride + np.arange(length)[None, :]
    return padded[idx]


def context_of(logmag_ctx):
    """First 200 frames (SN/apply.py:381-382); shorter clips are an error in the reference
    (tf.reshape to [200, 201] fails)."""
    if logmag_ctx.shape[0] < NOISE_WIN:
        raise ValueError("context clip yields %d < 200 STFT frames" % logmag_ctx.shape[0])
    return logmag_ctx[:NOISE_WIN]


# ---------------------------------------------------------------------------------------------
# A.6 network
# ---------------------------------------------------------------------------------------------
def same_pads(n, k, s):
    """TF 'SAME': out = ceil(n/s); pad = max((out-1) s + k - n, 0); before = pad//2."""
    out = -(-n // s)
    pad = max((out - 1) * s + k - n, 0)
    return pad // 2, pad - pad // 2


def _t(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


class Net:
    """Eval-mode forward of ``model()`` on NCHW torch tensors (weights are HWIO as in TF)."""

    def __init__(self, weights, variant=W.SELECTIVE_NOISE, dtype=torch.float32):
        self.w = {k: _t(v, dtype) for k, v in weights.items()}
        self.variant = variant
        self.dtype = dtype
        self.sa, self.sb = W.cond_names(variant)

    # blocks.py:38-48
    def conv(self, x, scope, stride, bias, padding="SAME"):
        w = self.w[scope + "/w"]                                   # [kh, kw, cin, cout]
        kh, kw = w.shape[0], w.shape[1]
        if padding == "SAME":
            pt, pb = same_pads(x.shape[2], kh, stride[0])
            pl, pr = same_pads(x.shape[3], kw, stride[1])
            x = F.pad(x, (pl, pr, pt, pb))
        y = F.conv2d(x, w.permute(3, 2, 0, 1).contiguous(), stride=stride)
        if bias:
            y = y + self.w[scope + "/b"].reshape(1, -1, 1, 1)
        return y

    # blocks.py:104-108
    def bn(self, x, scope):
        g = self.w[scope + "/gamma"].reshape(-1)
        b = self.w[scope + "/beta"].reshape(-1)
        m = self.w[scope + "/pop_mean"].reshape(-1)
        v = self.w[scope + "/pop_variance"].reshape(-1)
        shape = (1, -1, 1, 1) if x.dim() == 4 else (1, -1)
        inv = torch.rsqrt(v + BN_EPS) * g
        return x * inv.reshape(shape) + (b - m * inv).reshape(shape)

    # main.py:102-124
    def noise_resnet_block(self, x, stride, c, scope):
        p1 = self.conv(x, scope + "_conv1", stride, False)
        p1 = torch.relu(self.bn(p1, scope + "_conv1"))
        p1 = self.conv(p1, scope + "_conv2", (1, 1), True)
        if x.shape[1] == c:
            p2 = x
        else:
            p2 = self.conv(x, scope + "_transform", stride, True)
        return torch.relu(self.bn(p1 + p2, scope + "_addition"))

    # main.py:189-202
    def tower(self, ctx):
        """ctx [R, 200, 201] -> embedding [R, 512]."""
        x = ctx[:, None, :, :]
        for name, _, stride, c in W.TOWER_BLOCKS:
            x = self.noise_resnet_block(x, stride, c, "embedding/" + name)
        assert x.shape[2:] == (23, 26), x.shape
        return x.mean(dim=(2, 3))

    # main.py:127-137
    def cont_embed(self, n, scope):
        v = torch.arange(n, dtype=self.dtype).reshape(n, 1)
        v = v @ self.w[scope + "_dense1/w"]
        v = torch.relu(self.bn(v, scope + scope + "_dense1"))
        v = v @ self.w[scope + "_dense2/w"]
        v = torch.relu(self.bn(v, scope + scope + "_dense2"))
        return v @ self.w[scope + "_dense3/w"]

    # main.py:139-159
    def cond(self, like, emb_a, emb_b, scope):
        pa = emb_a @ self.w[scope + self.sa + "/w"] + self.w[scope + self.sa + "/b"]
        pb = emb_b @ self.w[scope + self.sb + "/w"] + self.w[scope + self.sb + "/b"]
        ts, fs = like.shape[2], like.shape[3]
        tout = self.cont_embed(ts, scope + "_temb")                 # [ts, C]
        fout = self.cont_embed(fs, scope + "_femb")                 # [fs, C]
        return (pa[:, :, None, None] + pb[:, :, None, None]
                + tout.t()[None, :, :, None] + fout.t()[None, :, None, :])

    # main.py:126-187
    def resnet_block(self, x, emb_a, emb_b, k, s, c, scope):
        p1 = self.conv(x, scope + "_conv1", (s, s), False)
        p1 = p1 + self.cond(p1, emb_a, emb_b, scope + "_conv1")
        p1 = torch.relu(self.bn(p1, scope + "_conv1"))
        p1 = self.conv(p1, scope + "_conv2", (1, 1), True)
        p1 = p1 + self.cond(p1, emb_a, emb_b, scope + "_conv2")
        if x.shape[1] == c:
            p2 = x
        else:
            p2 = self.conv(x, scope + "_transform", (s, s), True)
        return torch.relu(self.bn(p1 + p2, scope + "_addition"))

    # main.py:218-242
    def mask_net(self, mixed, emb_a, emb_b, taps=None):
        """mixed [mb, 35, 201]; emb_* [mb, 512] -> denoised log-magnitude [mb, 201].
        ``taps`` (dict) optionally collects every block output in NHWC for layer-wise tests."""
        x = mixed[:, None, :, :]
        for name, k, s, c in W.MAIN_BLOCKS:
            x = self.resnet_block(x, emb_a, emb_b, k, s, c, name)
            if taps is not None:
                taps[name] = x.permute(0, 2, 3, 1).contiguous()
        x = self.conv(x, "last_conv", (1, 1), False, padding="VALID")
        x = torch.relu(self.bn(x, "last_conv"))
        if taps is not None:
            taps["last_conv"] = x.permute(0, 2, 3, 1).contiguous()
        x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)          # NHWC flatten: f*512 + c
        out = x @ self.w["last_dense/w"] + self.w["last_dense/b"]
        return mixed[:, MIX_WIN // 2, :] + out

    def forward(self, mixed, ctx_a, ctx_b):
        """The reference graph: both towers evaluated for every row of the mini-batch."""
        return self.mask_net(mixed, self.tower(ctx_a), self.tower(ctx_b))


# ---------------------------------------------------------------------------------------------
# A.7 reconstruction
# ---------------------------------------------------------------------------------------------
def inverse_window():
    """tf.signal.inverse_stft_window_fn(160, hann periodic)(400): w[n] / sum_j w^2[n mod 160 + 160 j]
    with w^2 zero-extended to 480."""
    w = hann_periodic().astype(np.float64)
    w2 = np.concatenate([w * w, np.zeros(480 - WIN)])
    denom = w2.reshape(3, HOP).sum(axis=0)                          # [160]
    return (w / np.tile(denom, 3)[:WIN]).astype(np.float32)


def istft(logmag, phase):
    """recover_samples_from_spectrum without the file write (SN/apply.py:189-201): float32 samples
    of length (T-1)*160 + 400."""
    T = logmag.shape[0]
    if T == 0:
        return np.zeros(0, np.float32)
    spec = np.exp(logmag.astype(np.float32)).astype(np.complex64) * np.exp(1j * phase.astype(np.float32)).astype(np.complex64)
    frames = np.fft.irfft(spec.astype(np.complex128), n=WIN, axis=1) * inverse_window().astype(np.float64)[None, :]
    out = np.zeros((T - 1) * HOP + WIN, dtype=np.float64)
    for t in range(T):
        out[t * HOP:t * HOP + WIN] += frames[t]
    return out.astype(np.float32)


def to_int16(y, peak):
    """New-repo convention (not reference behaviour): undo the peak normalisation of
    SN/apply.py:150 in float32, round half to even, saturate."""
    v = np.asarray(y, np.float32) * np.float32(peak + 0.000001)
    return np.clip(np.rint(v), -32768, 32767).astype(np.int16)


# ---------------------------------------------------------------------------------------------
# apply_snc / apply_separator restated on arrays
# ---------------------------------------------------------------------------------------------
def apply_arrays(net, mix_pcm, ctx_a_pcm, ctx_b_pcm, faithful=False, mb=100, return_all=False):
    """SN/apply.py:339-457 (SS/apply.py:288-397) on in-memory PCM.

    ctx_a / ctx_b follow ``weights.cond_names``: SN (pos, neg), SS (interference, target)."""
    mix = normalise(mix_pcm)
    mix = mix[:trim_len(len(mix))]
    lm, ph = logmag_phase(mix)
    ca = context_of(logmag_phase(normalise(ctx_a_pcm))[0])
    cb = context_of(logmag_phase(normalise(ctx_b_pcm))[0])
    windows = strided_crop(lm, MIX_WIN, 1)                          # [T, 35, 201]
    T = windows.shape[0]
    dt = net.dtype
    den = []
    with torch.no_grad():
        if faithful:
            ta = _t(ca, dt)[None].expand(mb, -1, -1)
            tb = _t(cb, dt)[None].expand(mb, -1, -1)
            for i in range(int(math.ceil(T / float(mb)))):
                b = _t(windows[i * mb:(i + 1) * mb], dt)
                den.append(net.forward(b, ta[:b.shape[0]].contiguous(), tb[:b.shape[0]].contiguous()))
        else:
            ea = net.tower(_t(ca, dt)[None])
            eb = net.tower(_t(cb, dt)[None])
            for i in range(int(math.ceil(T / float(mb)))):
                b = _t(windows[i * mb:(i + 1) * mb], dt)
                den.append(net.mask_net(b, ea.expand(b.shape[0], -1), eb.expand(b.shape[0], -1)))
    den = torch.cat(den, 0).to(torch.float32).numpy() if den else np.zeros((0, NBIN), np.float32)
    y = istft(den, ph)
    if not return_all:
        return y
    return dict(samples=y, denoised=den, logmag=lm, phase=ph, ctx_a=ca, ctx_b=cb,
                mixed_processed=istft(windows[:, MIX_WIN // 2, :], ph), peak=peak_of(mix_pcm))


def post_mix(denoised_samples, mixed_samples, compensate=0.0, ac=False):
    """SN/apply.py:459-470."""
    removed = mixed_samples - denoised_samples
    snr_est = float(np.mean(np.square(denoised_samples)) / np.mean(np.square(removed)))
    factor = snr_est / 20 if ac else compensate
    return removed, snr_est, denoised_samples + removed * factor


# ---------------------------------------------------------------------------------------------
# apply_demo (row n3): on-the-fly mixing + processing from frame 200 on
# ---------------------------------------------------------------------------------------------
def _pysum(x):
    """Python builtin sum() over a float32 array as the reference calls it (SN/apply.py:77): sequential, and -
    starting from int 0 under pre-NEP-50 numpy - carried in float64."""
    acc = 0.0
    for v in np.asarray(x, np.float64).tolist():
        acc += v
    return acc


def _repeat_or_cut(noise, n):
    """SN/apply.py:57-72."""
    nse = noise
    while n - len(nse) > 0:
        diff = n - len(nse)
        nse = np.concatenate([nse, noise[:diff]], axis=0)
    if n - len(noise) < 0:
        nse = noise[:n]
    return nse


def domixing_sn(clean, pos, neg, snr_pos=0, snr_neg=0, with_target=False):
    """SN/apply.py:56-104 == SN/reader.py:131-180 -> (mixed, noise_pos_signal, noise_neg_signal[, target]), float32."""
    sig = clean
    nse_pos, nse_neg = _repeat_or_cut(pos, len(sig)), _repeat_or_cut(neg, len(sig))
    ps = _pysum(np.abs(sig) * np.abs(sig)) / sig.shape[0]
    pp = _pysum(np.abs(nse_pos) * np.abs(nse_pos)) / nse_pos.shape[0]
    pn = _pysum(np.abs(nse_neg) * np.abs(nse_neg)) / nse_neg.shape[0]
    kp = 1.0 if pp == 0 else math.sqrt((ps / pp) * pow(10, -snr_pos / 10.0))
    kn = 1.0 if pn == 0 else math.sqrt((ps / pn) * pow(10, -snr_neg / 10.0))
    a = (np.float32(kp) * nse_pos).astype(np.float32)
    b = (np.float32(kn) * nse_neg).astype(np.float32)
    mixed = (sig + a + b).astype(np.float32)
    mixed = (mixed / np.float32(float(np.max(np.abs(mixed))) + 0.000001)).astype(np.float32)
    d = np.float32(float(np.max(np.abs(mixed))) + 0.000001)
    if with_target:
        return mixed, (a / d).astype(np.float32), (b / d).astype(np.float32), ((sig + a) / d).astype(np.float32)
    return mixed, (a / d).astype(np.float32), (b / d).astype(np.float32)


def domixing_ss(clean, noise, snr=0):
    """SS/apply.py:55-80 -> (mixed, K)."""
    nse = _repeat_or_cut(noise, len(clean))
    ps = _pysum(np.abs(clean) * np.abs(clean)) / clean.shape[0]
    pn = _pysum(np.abs(nse) * np.abs(nse)) / nse.shape[0]
    k = math.sqrt(1.0 if pn == 0 else (ps / pn) * pow(10, -snr / 10.0))
    mixed = (clean + np.float32(k) * nse).astype(np.float32)
    return (mixed / np.float32(float(np.max(np.abs(mixed))) + 0.000001)).astype(np.float32), k


def demo_signals(variant, clean_pcm, noise_a_pcm, noise_b_pcm=None):
    """combine_signals (SN/apply.py:107-139, SS/apply.py:83-108) on in-memory PCM.
    -> (mixed, ctx_a signal, ctx_b signal) in the engine's ctx order (SN: pos, neg; SS: interference, target)."""
    clean = normalise(clean_pcm)
    rem = (len(clean) - WIN) % HOP
    if rem:
        clean = clean[:-rem]
    if variant == W.SELECTIVE_NOISE:
        mixed, pos_sig, neg_sig = domixing_sn(clean, normalise(noise_a_pcm), normalise(noise_b_pcm))
        return mixed, pos_sig, neg_sig
    noise = normalise(noise_a_pcm)
    mixed, k = domixing_ss(clean, noise)
    return mixed, (noise * np.float32(k)).astype(np.float32), clean


def apply_demo_arrays(net, mixed, sig_a, sig_b, mb=100):
    """SN/apply.py:247-337 / SS/apply.py:198-285 after combine_signals -> (denoised samples, mixture-centre samples)."""
    lm, ph = logmag_phase(mixed)
    ca = context_of(logmag_phase(sig_a)[0])
    cb = context_of(logmag_phase(sig_b)[0])
    windows = strided_crop(lm[NOISE_WIN:], MIX_WIN, 1)
    phs = ph[NOISE_WIN:]
    dt = net.dtype
    den = []
    with torch.no_grad():
        ea = net.tower(_t(ca, dt)[None])
        eb = net.tower(_t(cb, dt)[None])
        for i in range(int(math.ceil(windows.shape[0] / float(mb)))):
            b = _t(windows[i * mb:(i + 1) * mb], dt)
            den.append(net.mask_net(b, ea.expand(b.shape[0], -1), eb.expand(b.shape[0], -1)))
    den = torch.cat(den, 0).to(torch.float32).numpy()
    return istft(den, phs), istft(windows[:, MIX_WIN // 2, :], phs)


# ---------------------------------------------------------------------------------------------
# eval-mode reader + model outputs (row n4)
# ---------------------------------------------------------------------------------------------
SNRS_SN = [-3, 0, 3, 5, 8]                 # SN/reader.py:205
SNRS_SS = [-5, -3, -1, 0, 1, 3, 5]         # SS/reader.py:138


def eval_snrs(variant, cleanpath):
    """SN/reader.py:215-219 / SS/reader.py:147-148: SNRs picked by the MD5 of the clean path (bytes)."""
    import hashlib
    h = hashlib.md5(cleanpath if isinstance(cleanpath, bytes) else cleanpath.encode("utf-8")).hexdigest()
    if variant == W.SELECTIVE_NOISE:
        return SNRS_SN[int(h[:8], 16) % len(SNRS_SN)], SNRS_SN[int(h[:6], 16) % len(SNRS_SN)]
    return (SNRS_SS[int(h[:8], 16) % len(SNRS_SS)],)


def eval_outputs_arrays(net, variant, cleanpath, clean_pcm, noise_a_pcm, noise_b_pcm=None, mb=100):
    """get_examples(istrain=False) (SN/reader.py:398-420, SS/reader.py:300-313) + model outputs and per-example
    loss (SN/main.py:231-252, SS/main.py:253-264) for one seed tuple -> dict of arrays."""
    clean = normalise(clean_pcm)
    rem = (len(clean) - WIN) % HOP
    if rem:
        clean = clean[:-rem]
    snrs = eval_snrs(variant, cleanpath)
    out = {}
    if variant == W.SELECTIVE_NOISE:
        mixed, pos_sig, neg_sig, target = domixing_sn(clean, normalise(noise_a_pcm), normalise(noise_b_pcm), snrs[0], snrs[1], with_target=True)
        sig_a, sig_b = pos_sig, neg_sig
        plm, pph = logmag_phase(pos_sig)
        nlm, nph = logmag_phase(neg_sig)
        out.update(pos=plm[NOISE_WIN:], posph=pph[NOISE_WIN:], neg=nlm[NOISE_WIN:], negph=nph[NOISE_WIN:],
                   snr_pos=snrs[0], snr_neg=snrs[1])
    else:
        noise = normalise(noise_a_pcm)
        mixed, k = domixing_ss(clean, noise, snrs[0])
        target = clean
        sig_a, sig_b = (noise * np.float32(k)).astype(np.float32), clean      # noisecontext, cleancontext
        out.update(snr=snrs[0])
    lm, ph = logmag_phase(mixed)
    tlm, tph = logmag_phase(target)
    ca, cb = context_of(logmag_phase(sig_a)[0]), context_of(logmag_phase(sig_b)[0])
    windows = strided_crop(lm[NOISE_WIN:], MIX_WIN, 1)
    dt = net.dtype
    den = []
    with torch.no_grad():
        ea, eb = net.tower(_t(ca, dt)[None]), net.tower(_t(cb, dt)[None])
        for i in range(int(math.ceil(windows.shape[0] / float(mb)))):
            b = _t(windows[i * mb:(i + 1) * mb], dt)
            den.append(net.mask_net(b, ea.expand(b.shape[0], -1), eb.expand(b.shape[0], -1)))
    den = torch.cat(den, 0).to(torch.float32).numpy()
    tgt = tlm[NOISE_WIN:]
    imp = np.linspace(2, 1, NBIN, dtype=np.float32).reshape(1, NBIN)
    loss = np.mean(np.square(den - tgt) * imp, axis=1, dtype=np.float32)
    out.update(loss=loss, mixed=lm[NOISE_WIN:], denoised=den, mixedph=ph[NOISE_WIN:],
               location=np.arange(den.shape[0], dtype=np.int32))
    if variant == W.SELECTIVE_NOISE:
        out.update(target=tgt, targetph=tph[NOISE_WIN:])
    else:
        out.update(clean=tgt)
    return out
