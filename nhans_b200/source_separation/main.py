"""Evaluation half of N_HANS___Source_Separation/main.py on the B200 engine:
the eval loop of ``save_and_eval`` and ``evaluate`` (SS/main.py:273-334).  Training is out of scope."""
from __future__ import annotations

import os

import numpy as np

from .. import weights as W
from ..session import get_engine
from ..wavio import FS, write_wav
from . import reader

VARIANT = W.SEPARATOR


class _Flags:
    eval_before_training = True
    dump_results = ""
    wav_dump_folder = "./wav_dump/"
    eval_mb = 100
    Fs = FS


FLAGS = _Flags()
modelname = "nhans_b200"


def run_eval(ereader, engine=None):
    engine = engine or get_engine(VARIANT)
    agg = {}
    for out in ereader.get_examples(engine):
        for k, v in out.items():
            agg.setdefault(k, []).append(v)
    return {k: np.concatenate(v) for k, v in agg.items()}


def evaluate(outputs, ereader, step, engine=None):
    """SS/main.py:273-334: mean loss + mixed / denoised waveforms per utterance.  -> mean loss."""
    engine = engine or get_engine(VARIANT)
    print(ereader.name)
    loss = float(outputs["loss"].mean())
    print("loss: {}".format(loss))
    starts = np.where(outputs["location"] == 0)[0]
    os.makedirs(FLAGS.wav_dump_folder, exist_ok=True)
    for i, s in enumerate(starts):
        e = len(outputs["mixed"]) if i == len(starts) - 1 else starts[i + 1]
        fo = np.array([0, e - s], np.int64)
        name = lambda p: p.decode("utf-8").split("/")[-1][:-4]
        tag = "{}_{}_{}_{}_{}".format(modelname, step, name(outputs["cleanpath"][s]), name(outputs["noisepath"][s]), outputs["snr"][s])
        for kind in ("mixed", "denoised"):
            y, _ = engine.istft(outputs[kind][s:e], outputs["mixedph"][s:e], fo)
            write_wav(os.path.join(FLAGS.wav_dump_folder, "{}_{}.wav".format(tag, kind)), y)
    return loss


def eval_before_training(names=("valid",), engine=None):
    losses = {}
    for n in names:
        er = reader.read_seeds(n)
        er.preparations()
        losses[n] = evaluate(run_eval(er, engine), er, 0, engine)
    return losses
