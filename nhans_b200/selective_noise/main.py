"""Evaluation half of N_HANS___Selective_Noise/main.py on the B200 engine:
``--eval_before_training`` scoring = ``save_and_eval`` without the save (SN/main.py:470-547) and ``evaluate``
(:266-353).  Training (optimizer, queues, summaries) is out of scope."""
from __future__ import annotations

import os

import numpy as np

from .. import weights as W
from ..session import get_engine
from ..wavio import FS, write_wav
from . import reader

VARIANT = W.SELECTIVE_NOISE


class _Flags:                                              # SN/main.py:42-75 (the ones evaluation reads)
    eval_before_training = True
    dump_results = ""
    wav_dump_folder = "./wav_dump/"
    eval_mb = 100
    Fs = FS


FLAGS = _Flags()
modelname = "nhans_b200"


def run_eval(ereader, engine=None):
    """The `while True: batch -> outputs -> aggregators` loop of save_and_eval (SN/main.py:511-537), one
    utterance at a time instead of eval_mb windows at a time; -> aggregated outputs dict."""
    engine = engine or get_engine(VARIANT)
    agg = {}
    for out in ereader.get_examples(engine):
        for k, v in out.items():
            agg.setdefault(k, []).append(v)
    return {k: np.concatenate(v) for k, v in agg.items()}


def evaluate(outputs, ereader, step, engine=None):
    """SN/main.py:266-353: mean loss, then per utterance (location == 0 starts one) the mixed / denoised /
    target / posNoise / negNoise waveforms written to FLAGS.wav_dump_folder.  -> mean loss."""
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
        tag = "{}_{}_{}_{}_{}_{}_{}".format(modelname, step, name(outputs["cleanpath"][s]), name(outputs["noisepospath"][s]),
                                            name(outputs["noisenegpath"][s]), outputs["snr_pos"][s], outputs["snr_neg"][s])
        for kind, mag, phs in (("mixed", "mixed", "mixedph"), ("denoised", "denoised", "mixedph"), ("target", "target", "targetph"),
                               ("posNoise", "pos", "posph"), ("negNoise", "neg", "negph")):
            y, _ = engine.istft(outputs[mag][s:e], outputs[phs][s:e], fo)
            write_wav(os.path.join(FLAGS.wav_dump_folder, "{}_{}.wav".format(tag, kind)), y)
    return loss


def eval_before_training(names=("valid",), engine=None):
    """The eval pass `main.py --eval_before_training` runs before (instead of) training."""
    losses = {}
    for n in names:
        er = reader.read_seeds(n)
        er.preparations()
        outputs = run_eval(er, engine)
        if FLAGS.dump_results:
            os.makedirs(FLAGS.dump_results, exist_ok=True)
            for k, v in outputs.items():
                np.save(os.path.join(FLAGS.dump_results, "{}_{}_{}_{}".format(modelname, er.name, 0, k)), v)
        losses[n] = evaluate(outputs, er, 0, engine)
    return losses
