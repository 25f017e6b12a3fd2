"""Process-wide engine cache used by the apply.py mirrors and the CLIs: one context per (device, variant),
weights restored from the reference checkpoint directory (the reference restores './trained_model/<ckpt>' on
every call, N_HANS___Selective_Noise/apply.py:430-432; here it happens once per process).  Like the reference,
a missing checkpoint is an error; NHANS_ALLOW_RANDOM_INIT=1 opts into the seeded random initialisation of the
identical architecture (tests, benchmarks, machines whose checkpoint files are git-LFS pointers)."""
from __future__ import annotations

import os
import sys

from . import weights as W
from .engine import Engine

_ENGINES = {}


def get_engine(variant, device=None, model_dir=None):
    device = int(os.environ.get("NHANS_DEVICE", "0")) if device is None else device
    key = (device, variant)
    if key not in _ENGINES:
        model_dir = model_dir or os.environ.get("NHANS_MODEL_DIR", "./trained_model")
        eng = Engine(device, variant,
                     win_capacity=int(os.environ.get("NHANS_WIN_CAPACITY", "0")),
                     row_capacity=int(os.environ.get("NHANS_ROW_CAPACITY", "0")))
        try:
            src = eng.load_default_weights(model_dir, seed=int(os.environ.get("NHANS_SEED", "0")))
        except Exception:
            eng.close()
            raise
        if src != "checkpoint":
            sys.stderr.write("nhans_b200: NHANS_ALLOW_RANDOM_INIT is set and there is no trained checkpoint under %r: "
                             "using seeded random-init weights of the identical architecture\n" % model_dir)
        _ENGINES[key] = eng
    return _ENGINES[key]


def drop_engine(variant, device=None):
    """Forget (and destroy) a cached engine, e.g. after it reported NHANS_ERR_KERNEL; the next get_engine
    builds a fresh context."""
    device = int(os.environ.get("NHANS_DEVICE", "0")) if device is None else device
    eng = _ENGINES.pop((device, variant), None)
    if eng is not None:
        eng.close()


def close_all():
    for e in _ENGINES.values():
        e.close()
    _ENGINES.clear()
