"""Eval-mode mirror of N_HANS___Selective_Noise/reader.py on the B200 engine.

``combine_signals(istrain=False, ...)`` (SN/reader.py:183-223), ``read_seeds`` with the 'valid' / 'test' seed
lists (:228-262) and the eval branch of ``get_examples`` (:398-420).  The reference materialises one
[35, 201] window per frame and tiles the two [200, 201] contexts per window; here an example stream is the
list of seed tuples and the engine consumes the spectrogram rows directly (``Engine.eval_outputs``).  Training
mode (random crops, shuffling queues) is out of scope and raises."""
from __future__ import annotations

import hashlib
import pickle

import numpy as np

from .. import weights as W
from ..wavio import FS, read_wav
from .apply import _norm64, domixing

VARIANT = W.SELECTIVE_NOISE
SNRs = [-3, 0, 3, 5, 8]                                   # SN/reader.py:205


class _Flags:                                              # SN/reader.py:34-41
    Fs = FS
    window_frames = 35
    context_frames = 200
    random_slices = 50
    eval_seeds = "valid"
    speech_wav_dir = "./speech_wav_dir/"
    noise_wav_dir = "./noise_wav_dir/"


FLAGS = _Flags()


def _bytes(path):
    return path if isinstance(path, bytes) else str(path).encode("utf-8")


def eval_snrs(cleanpath):
    """SN/reader.py:215-219: validation / test SNRs depend on the clean file's path only."""
    h = hashlib.md5(_bytes(cleanpath)).hexdigest()
    return SNRs[int(h[:8], 16) % len(SNRs)], SNRs[int(h[:6], 16) % len(SNRs)]


def combine_signals(istrain, cleanpath, noisepospath, noisenegpath):
    """SN/reader.py:183-223 -> (target, noise_pos_signal, noise_neg_signal, mixed, snr_pos, snr_neg)."""
    if istrain:
        raise NotImplementedError("training-mode example generation is outside the inference hot path")
    dec = lambda p: p.decode("utf-8") if isinstance(p, bytes) else p
    clean = _norm64(read_wav(dec(cleanpath)))
    pos = _norm64(read_wav(dec(noisepospath)))
    neg = _norm64(read_wav(dec(noisenegpath)))
    rem = (len(clean) - 400) % 160
    if rem != 0:
        clean = clean[:-rem]
    snr_pos, snr_neg = eval_snrs(cleanpath)
    mixed, target, _, _, pos_sig, neg_sig = domixing(clean, pos, neg, snr_pos, snr_neg)
    return target, pos_sig, neg_sig, mixed, np.array(snr_pos, np.int32), np.array(snr_neg, np.int32)


class read_seeds:
    """SN/reader.py:228-300 for name in ('valid', 'test'): the seed lists and their pairing.  One example stream
    element = (clean seed, positive-noise seed, negative-noise seed); the two noise seeds are consecutive
    entries of the noise list, one epoch, no shuffling (input_producer(shuffle=False, num_epochs=1))."""

    def __init__(self, name, queuesize=0, min_after_dequeue=0, nthreads=1):
        if name == "train":
            raise NotImplementedError("training-mode reader is outside the inference hot path")
        self.Fs = FLAGS.Fs
        self.frame_length = int(self.Fs * 0.025)
        self.frame_step = int(self.Fs * 0.010)
        self.window_frames = FLAGS.window_frames
        self.context_frames = FLAGS.context_frames
        self.eval_stride = 1
        self.istrain = False
        self.name = name
        seeds = FLAGS.eval_seeds
        self.seedspaths = [FLAGS.speech_wav_dir + seeds + ".pkl", FLAGS.noise_wav_dir + seeds + ".pkl"]
        self.seeds = None

    def preparations(self):
        self.seeds = []
        for sp in self.seedspaths:
            with open(sp, "rb") as f:
                self.seeds.append(list(pickle.load(f)))
        return self

    def seed_tuples(self):
        if self.seeds is None:
            self.preparations()
        speech, noise = self.seeds
        n = min(len(speech), len(noise) // 2)
        return [(speech[i], noise[2 * i], noise[2 * i + 1]) for i in range(n)]

    def get_examples(self, engine):
        """Generator over seed tuples -> the model's ``outputs`` dict for that utterance (SN/main.py:247-251)."""
        for clean, pos, neg in self.seed_tuples():
            yield model_outputs(engine, clean, pos, neg)


def model_outputs(engine, cleanpath, noisepospath, noisenegpath):
    target, pos_sig, neg_sig, mixed, snr_pos, snr_neg = combine_signals(False, cleanpath, noisepospath, noisenegpath)
    o = engine.eval_outputs(mixed, target, pos_sig, neg_sig, extra=(pos_sig, neg_sig))
    n = len(o["location"])
    (pos, posph), (neg, negph) = o.pop("extra")
    o.update(pos=pos, posph=posph, neg=neg, negph=negph,
             cleanpath=np.array([_bytes(cleanpath)] * n), noisepospath=np.array([_bytes(noisepospath)] * n),
             noisenegpath=np.array([_bytes(noisenegpath)] * n),
             snr_pos=np.full(n, snr_pos, np.int32), snr_neg=np.full(n, snr_neg, np.int32))
    return o
