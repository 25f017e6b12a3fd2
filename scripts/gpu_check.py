"""Diagnostic run on a GPU box: every stage of the CUDA path against the oracle / the CPU plan interpreter,
ordered from least to most risky, with per-layer buffer comparisons.  Writes gpurun_out/gpu_check.log."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "gpu_check.log"), "w")


def say(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


from nhans_b200 import synth, weights as W
from nhans_b200.engine import Engine
from oracle import nhans_oracle as O
from oracle.planexec import PlanExec, grid_gather


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b) + 1e-30))


def main():
    variant = 0
    w = W.seeded_init(variant, 0)
    t = time.time()
    eng = Engine(0, variant, win_capacity=8, row_capacity=2)
    say("device", eng.device_info())
    eng.load_weights(w)
    say("load_weights ok %.2fs" % (time.time() - t))

    # ---------------- DSP ----------------
    clips = [synth.mixture(1.0, 0), synth.mixture(0.537, 1), synth.mixture(0.03, 2)[:450]]
    try:
        norm = eng.normalise(clips, trim=True)
        for c, n in zip(clips, norm):
            ref = O.normalise(c)
            ref = ref[:O.trim_len(len(ref))]
            say("normalise bit-exact:", len(n) == len(ref) and bool(np.array_equal(n.view(np.uint32), ref.view(np.uint32))))
        lm, ph, fo, peak = eng.stft(clips)
        say("frame_offs", fo.tolist(), "peak", peak.tolist())
        for u, c in enumerate(clips):
            rl, rp = O.logmag_phase(O.normalise(c))
            g = lm[fo[u]:fo[u + 1]]
            gp = ph[fo[u]:fo[u + 1]]
            mag = np.exp(rl.astype(np.float64))
            pherr = np.abs(np.exp(1j * gp.astype(np.float64)) - np.exp(1j * rp.astype(np.float64))) * mag
            say("stft clip", u, "frames", g.shape[0], "max|dlogmag|", float(np.abs(g - rl).max()) if g.size else 0,
                "max phase err x mag", float(pherr.max()) if g.size else 0, "max mag", float(mag.max()) if g.size else 0)
        y, oo = eng.istft(lm, ph, fo)
        for u, c in enumerate(clips):
            rl, rp = O.logmag_phase(O.normalise(c))
            ry = O.istft(rl, rp)
            gy = y[oo[u]:oo[u + 1]]
            say("istft clip", u, "len", len(gy), len(ry), "max err", float(np.abs(gy - ry).max()) if len(ry) else 0)
        y2, i16, oo = eng.istft(lm, ph, fo, peak=peak, want_i16=True)
        c0 = clips[0][:O.trim_len(len(clips[0]))]
        say("roundtrip int16 clip0 max |diff| (interior)", int(np.abs(i16[oo[0]:oo[1]][400:-400].astype(int) - c0[400:-400].astype(int)).max()))
    except Exception:
        say("DSP FAILED\n" + traceback.format_exc())

    # ---------------- tower ----------------
    pe = PlanExec(w, variant, win_cap=8, row_cap=2)
    ctx = np.stack([O.context_of(O.logmag_phase(O.normalise(synth.noise_clip(u)))[0]) for u in range(3)])
    try:
        t = time.time()
        emb = eng.embed(ctx)
        say("embed ok %.3fs" % (time.time() - t))
        emb_pe = pe.embed(ctx)
        say("embed vs planexec rel", rel(emb, emb_pe), "max", float(np.abs(emb - emb_pe).max()))
        # per-buffer comparison for the last chunk (row index 2 -> unit 0 of chunk 2)
        plan = eng.plan(1)
        for g in plan["bufs"]:
            a = grid_gather(g, eng.read_buffer(1, g["buf"]).astype(np.float32), 1)
            b = grid_gather(g, pe.read_buffer(1, g["buf"]), 1)
            say("  tower buf", g["buf"], (g["H"], g["W"], g["C"]), "rel", rel(a, b), "max", float(np.abs(a - b).max()), "ref max", float(np.abs(b).max()))
        import torch
        with torch.no_grad():
            emb_o = O.Net(w, variant).tower(torch.from_numpy(ctx)).numpy()
        say("embed vs oracle rel", rel(emb, emb_o))
    except Exception:
        say("TOWER FAILED\n" + traceback.format_exc())
        return 1

    # ---------------- mask net ----------------
    try:
        mixes = [synth.mixture(0.2, 0)[:400 + 160 * 9], synth.mixture(0.2, 1)[:400 + 160 * 4]]
        lms = [O.logmag_phase(O.normalise(m))[0] for m in mixes]
        fo = np.cumsum([0] + [l.shape[0] for l in lms])
        lmc = np.concatenate(lms)
        rng = np.random.default_rng(0)
        ea = rng.normal(0, 2, (2, 512)).astype(np.float32)
        eb = rng.normal(0, 2, (2, 512)).astype(np.float32)
        t = time.time()
        den = eng.masknet(lmc, fo, ea, eb)
        say("masknet ok %.3fs" % (time.time() - t))
        den_pe = pe.masknet(lmc, fo, ea, eb)
        say("masknet vs planexec: rel(out)", rel(den - lmc, den_pe - lmc), "max", float(np.abs(den - den_pe).max()))
        plan = eng.plan(0)
        nlast = int(fo[-1]) - 8          # units in the last chunk (cap 8)
        for g in plan["bufs"]:
            a = grid_gather(g, eng.read_buffer(0, g["buf"]).astype(np.float32), nlast)
            b = grid_gather(g, pe.read_buffer(0, g["buf"]), nlast)
            say("  main buf", g["buf"], (g["H"], g["W"], g["C"]), "rel", rel(a, b), "max", float(np.abs(a - b).max()), "ref max", float(np.abs(b).max()))
    except Exception:
        say("MASKNET FAILED\n" + traceback.format_exc())
        return 1

    # ---------------- end to end ----------------
    try:
        eng2 = Engine(0, variant)           # default capacities
        eng2.load_weights(w)
        mixes = [synth.mixture(1.0, 0), synth.mixture(0.6, 1)]
        negs = [synth.noise_clip(0), synth.noise_clip(1)]
        t = time.time()
        res = eng2.enhance(mixes, None, negs, want_mixproc=True)
        say("enhance ok %.3fs" % (time.time() - t))
        net = O.Net(w, variant)
        for u in range(2):
            r = O.apply_arrays(net, mixes[u], synth.silence(), negs[u], return_all=True)
            y = res["f32"][u]
            err = y - r["samples"]
            snr = 10 * np.log10(np.sum(r["samples"].astype(np.float64) ** 2) / (np.sum(err.astype(np.float64) ** 2) + 1e-30))
            say("e2e utt", u, "len", len(y), len(r["samples"]), "SNR dB", float(snr),
                "mixproc max err", float(np.abs(res["mixed_processed"][u] - r["mixed_processed"]).max()))
        # timing of a bigger batch
        mixes = [synth.mixture(4.0, u) for u in range(16)]
        negs = [synth.noise_clip(u) for u in range(16)]
        eng2.enhance(mixes, None, negs)
        eng2.profile(True)
        t = time.time()
        eng2.enhance(mixes, None, negs)
        dt = time.time() - t
        st = eng2.profile_get(0)
        say("16 x 4 s: %.3f s wall -> %.1f audio-s/s; gemm %.1f ms, %.1f TFLOP/s" % (dt, 64 / dt, st["ms"], st["flops"] / st["ms"] / 1e9))
        for k in range(5):
            say("  kind", k, eng2.profile_get(k))
    except Exception:
        say("E2E FAILED\n" + traceback.format_exc())
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
