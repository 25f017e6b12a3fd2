"""Generates tests/golden/golden_{sn,ss}.npz with the CPU oracle (oracle/nhans_oracle.py, float32 torch):

    python tests/golden/make_golden.py

Inputs are the deterministic synthetic signals of nhans_b200/synth.py and the seeded random-init weights
(seed 0) - the reference itself cannot run here (no TensorFlow, weights are git-LFS pointers), so these
vectors pin the ORACLE against drift and give the GPU tests a fixture that does not need the oracle's
network code; they are not reference outputs."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from nhans_b200 import synth, weights as W  # noqa: E402
from oracle import nhans_oracle as O  # noqa: E402


def inputs(tag):
    if tag == "sn":
        return synth.mixture(0.15, 7), synth.silence(), synth.noise_clip(7, "neg")
    return synth.mixture(0.12, 9), synth.speaker_clip(9, "interference"), synth.speaker_clip(9, "target")


if __name__ == "__main__":
    for tag, variant in (("sn", 0), ("ss", 1)):
        net = O.Net(W.seeded_init(variant, 0), variant)
        mix, a, b = inputs(tag)
        r = O.apply_arrays(net, mix, a, b, faithful=False, return_all=True)
        np.savez_compressed(os.path.join(HERE, "golden_%s.npz" % tag), logmag=r["logmag"], phase=r["phase"],
                            denoised=r["denoised"], samples=r["samples"], mixed_processed=r["mixed_processed"])
        print(tag, r["denoised"].shape, float(np.sqrt(np.mean((r["denoised"] - r["logmag"]) ** 2))))
