"""Eval-mode mirror of N_HANS___Source_Separation/reader.py on the B200 engine:
``combine_signals(istrain=False, ...)`` (SS/reader.py:120-153), ``read_seeds`` ('valid' / 'test') and the eval
branch of ``get_examples`` (:300-313)."""
from __future__ import annotations

import hashlib
import pickle

import numpy as np

from .. import weights as W
from ..selective_noise.apply import _norm64
from ..wavio import FS, read_wav
from .apply import domixing

VARIANT = W.SEPARATOR
SNRs = [-5, -3, -1, 0, 1, 3, 5]                           # SS/reader.py:138


class _Flags:
    Fs = FS
    window_frames = 35
    context_frames = 200
    eval_seeds = "valid"
    speech_wav_dir = "./speech_wav_dir/"
    noise_wav_dir = "./noise_wav_dir/"


FLAGS = _Flags()


def _bytes(path):
    return path if isinstance(path, bytes) else str(path).encode("utf-8")


def eval_snr(cleanpath):
    """SS/reader.py:147-148."""
    return SNRs[int(hashlib.md5(_bytes(cleanpath)).hexdigest()[:8], 16) % len(SNRs)]


def combine_signals(istrain, cleanpath, noisepath):
    """SS/reader.py:120-153 -> (clean, noise * K, mixed, snr)."""
    if istrain:
        raise NotImplementedError("training-mode example generation is outside the inference hot path")
    dec = lambda p: p.decode("utf-8") if isinstance(p, bytes) else p
    clean = _norm64(read_wav(dec(cleanpath)))
    noise = _norm64(read_wav(dec(noisepath)))
    rem = (len(clean) - 400) % 160
    if rem != 0:                                           # the reference's clean[:-0] would empty the signal
        clean = clean[:-rem]
    snr = eval_snr(cleanpath)
    mixed, K = domixing(clean, noise, snr)
    return clean, noise * np.float32(K), mixed, np.array(snr, np.int32)


class read_seeds:
    """SS/reader.py:161-225 for 'valid' / 'test': (target-speaker seed, interfering-speaker seed) pairs in list
    order, one epoch, no shuffling."""

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
        return list(zip(self.seeds[0], self.seeds[1]))

    def get_examples(self, engine):
        for clean, noise in self.seed_tuples():
            yield model_outputs(engine, clean, noise)


def model_outputs(engine, cleanpath, noisepath):
    """SS/main.py:253-264 outputs for one seed pair."""
    clean, noise_k, mixed, snr = combine_signals(False, cleanpath, noisepath)
    o = engine.eval_outputs(mixed, clean, noise_k, clean)      # ctx_a = noisecontext, ctx_b = cleancontext
    o.pop("extra")
    n = len(o["location"])
    o["clean"] = o.pop("target")
    o.pop("targetph")
    o.update(cleanpath=np.array([_bytes(cleanpath)] * n), noisepath=np.array([_bytes(noisepath)] * n), snr=np.full(n, snr, np.int32))
    return o
