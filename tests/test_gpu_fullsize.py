"""BASELINE.json configurations at their full sizes, through the default-capacity engine (2048-window passes, so
every layer - including the head - runs the CTA-pair kernel the benchmark runs).  The oracle needs seconds per
utterance, so a few utterances are compared with it directly (SNR >= 40 dB) and the rest of the batch is covered by
size-independent properties: copies of an utterance give bit-identical output wherever they sit in the batch /
whichever pass they fall into, outputs have the trimmed length, and a small-capacity engine (other tile and
kernel choices) agrees to fp16-accumulation-order noise."""
import numpy as np
import pytest

from nhans_b200 import synth, weights as W
from nhans_b200.engine import Engine
from oracle import nhans_oracle as O

pytestmark = pytest.mark.gpu


def _snr(ref, got):
    ref = ref.astype(np.float64)
    err = got.astype(np.float64) - ref
    return 10 * np.log10(np.sum(ref ** 2) / (np.sum(err ** 2) + 1e-30))


@pytest.fixture(scope="module")
def big_sn(weights_sn):
    e = Engine(0, W.SELECTIVE_NOISE)                   # default capacities: what bench.py and the CLIs use
    e.load_weights(weights_sn)
    yield e
    e.close()


def test_cfg2_denoise_256x4s(big_sn, engine_sn, oracle_sn):
    """configs[1]: speech denoising, 256 x 4 s, --neg conditioning (the benchmark workload)."""
    base_m = [synth.mixture(4.0, u) for u in range(16)]
    base_n = [synth.noise_clip(u) for u in range(16)]
    order = [(7 * u + 3) % 16 for u in range(256)]     # 16 distinct utterances, 16 copies each, interleaved
    res = big_sn.enhance([base_m[i] for i in order], None, [base_n[i] for i in order])
    first = {}
    for u, i in enumerate(order):
        assert len(res["f32"][u]) == O.trim_len(64000)
        if i in first:
            assert np.array_equal(res["i16"][u], res["i16"][first[i]])
            assert np.array_equal(res["f32"][u], res["f32"][first[i]])
        else:
            first[i] = u
    assert len(first) == 16
    for i in (0, 5, 9, 14):                            # four distinct utterances against the oracle
        ref = O.apply_arrays(oracle_sn, base_m[i], synth.silence(), base_n[i])
        assert _snr(ref, res["f32"][first[i]]) >= 40.0
    # the same utterance through 256-window passes (single-CTA head, other tile walk)
    small = engine_sn.enhance([base_m[3]], None, [base_n[3]])
    assert _snr(small["f32"][0], res["f32"][first[3]]) >= 60.0


def test_cfg3_selective_pos_neg_8s(big_sn, oracle_sn):
    """configs[2]: selective noise suppression with --pos and --neg, 8 s utterances (batch reduced to 32: the
    per-utterance arithmetic does not depend on the batch size, which the copies check)."""
    base = [(synth.mixture(8.0, 20 + u), synth.noise_clip(20 + u, "pos"), synth.noise_clip(20 + u, "neg")) for u in range(4)]
    order = [u % 4 for u in range(32)]
    res = big_sn.enhance([base[i][0] for i in order], [base[i][1] for i in order], [base[i][2] for i in order])
    for u, i in enumerate(order):
        assert len(res["f32"][u]) == O.trim_len(128000)
        assert np.array_equal(res["f32"][u], res["f32"][i])
    ref = O.apply_arrays(oracle_sn, *base[1])
    assert _snr(ref, res["f32"][1]) >= 40.0
    # swapping the conditioning clips changes the result (the two towers are different networks; with the seeded
    # random-init weights the conditioning path is weak, ~54 dB below the signal, but far above the kernel noise)
    sw = big_sn.enhance([base[1][0]], [base[1][2]], [base[1][1]])
    ref_sw = O.apply_arrays(oracle_sn, base[1][0], base[1][2], base[1][1])
    assert _snr(res["f32"][1], sw["f32"][0]) < 60.0
    assert _snr(ref_sw - ref, sw["f32"][0] - res["f32"][1]) >= 3.0       # and the *change* itself matches the oracle's


def test_cfg4_separator_10s(weights_ss, oracle_ss):
    """configs[3]: N_HANS___Source_Separation on 10 s mixtures (target / interference speaker conditioning)."""
    eng = Engine(0, W.SEPARATOR)
    eng.load_weights(weights_ss)
    base = [(synth.mixture(10.0, 40 + u), synth.speaker_clip(40 + u, "interference"), synth.speaker_clip(40 + u, "target")) for u in range(2)]
    order = [u % 2 for u in range(16)]
    res = eng.enhance([base[i][0] for i in order], [base[i][1] for i in order], [base[i][2] for i in order])
    for u, i in enumerate(order):
        assert len(res["f32"][u]) == O.trim_len(160000)
        assert np.array_equal(res["i16"][u], res["i16"][i])
    ref = O.apply_arrays(oracle_ss, *base[0])
    assert _snr(ref, res["f32"][0]) >= 40.0
    eng.close()


def test_cfg5_shard_512x10s(big_sn, oracle_sn):
    """configs[4]: the utterance-sharded sweep gives each of 8 GPUs 1024 x 10 s clips; half such a shard (512 x 10 s,
    0.5 M windows, 250 passes) goes through one nhans_enhance_batch call here.  Copies must come back bit-identical
    wherever they sit, outputs have the trimmed length, a clip processed alone gives the same samples, and one 10 s
    utterance of the batch is compared with the oracle (998 windows)."""
    base_m = [synth.mixture(10.0, 60 + u) for u in range(4)]
    base_n = [synth.noise_clip(60 + u) for u in range(4)]
    order = [(3 * u + 1) % 4 for u in range(512)]
    res = big_sn.enhance([base_m[i] for i in order], None, [base_n[i] for i in order], want_f32=False)
    first = {}
    for u, i in enumerate(order):
        assert len(res["i16"][u]) == O.trim_len(160000)
        if i in first:
            assert np.array_equal(res["i16"][u], res["i16"][first[i]])
        else:
            first[i] = u
    solo = big_sn.enhance([base_m[2]], None, [base_n[2]], want_f32=True)
    assert np.array_equal(solo["i16"][0], res["i16"][first[2]])
    ref = O.apply_arrays(oracle_sn, base_m[2], synth.silence(), base_n[2], return_all=True)
    assert _snr(ref["samples"], solo["f32"][0]) >= 40.0
    i16 = O.to_int16(ref["samples"], ref["peak"]).astype(np.int32)
    assert np.abs(res["i16"][first[2]].astype(np.int32) - i16).max() <= max(2, int(3e-3 * np.abs(i16).max()))
    assert np.abs(res["i16"][first[2]].astype(np.int32)).max() > 1000      # not silence
