#!/usr/bin/env python
"""Benchmark of the N-HANS inference hot path on B200 (contract: see the task statement / DESIGN.md §6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--utts U] [--seconds S]

Metric (BASELINE.json): audio-seconds denoised per wall-second.  A step = one pass of the hot path
(STFT -> embedding towers -> conditioned residual network -> iSTFT/overlap-add) over one batch of synthetic
16 kHz utterances with --neg conditioning; N = 1 runs BASELINE config 2 (256 x 4 s on one B200), N > 1 runs
the same batch on every GPU (weak scaling, utterance-sharded, no collective on the data path).
`value` is measured with the inputs resident in HBM (CUDA events on the engine's stream); `e2e` goes through
nhans_enhance_batch with host buffers (pinned), H2D and D2H inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 16000
GFLOP_PER_WINDOW = 10.327145856       # 2 * 5 163 572 928 MAC
CLOCK_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
               "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + CLOCK_QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = []
        reasons = set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            busy = [v for v in sm if v > 0.5 * max(sm)] or sm
            out["sm_mhz"] = float(np.median(busy))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1400.0, 6650.0, "fallback"


def cpu_baseline(weights, variant, seconds, steps=1, warmup=0):
    """The reference's own structure on the host cores: oracle faithful mode (mb = 100, both towers per
    window, materialised windows), torch-CPU fp32, all threads."""
    import torch
    from nhans_b200 import synth
    from oracle import nhans_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    net = O.Net(weights, variant)
    mix, neg, pos = synth.mixture(seconds, 0), synth.noise_clip(0), synth.silence()
    times = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        O.apply_arrays(net, mix, pos, neg, faithful=True)
        if i >= warmup:
            times.append(time.perf_counter() - t)
    dt = float(np.mean(times))
    return dict(value=seconds / dt, unit="audio-s/s", cores=os.cpu_count() or 1, kind="port",
                sample="oracle faithful mode (torch-CPU fp32, mb=100, towers per window) on one %.1f s utterance + --neg clip, "
                       "%d step(s), %.1f s per step" % (seconds, steps, dt))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--win-capacity", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from nhans_b200 import synth, weights as W
    variant = W.SELECTIVE_NOISE
    weights, wsrc = W.load_or_init(variant, os.environ.get("NHANS_MODEL_DIR"), 0)
    workload = "nhans_denoiser, %d x %.0f s 16 kHz utterances + --neg clip per GPU (BASELINE config 2)" % (a.utts, a.seconds)
    base = {"metric": "audio-sec/sec", "unit": "audio-s/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
            "config": {"workload": workload, "utterances_per_gpu": a.utts, "seconds": a.seconds, "weights": wsrc,
                       "l2": "working set per step (GBs of fp16 activations, 164 MB of spectra) exceeds the 126 MB L2",
                       "parallelism": "utterance-sharded x%d, no collective" % a.gpus}}

    if a.impl == "reference":
        # The reference (TensorFlow graph code) cannot run on this image; its CPU path is the oracle's
        # faithful mode.  Rank 0 alone runs it; each step is a bounded sample (one 1 s utterance).
        if rank != 0:
            return 0
        secs = 1.0
        cb = cpu_baseline(weights, variant, secs, steps=a.steps, warmup=a.warmup)
        line = dict(base)
        line.update({"impl": "reference", "value": cb["value"], "ms_per_step": 1e3 * secs / cb["value"], "dtype": "f32",
                     "cpu_baseline": cb, "gpu_launches": 0,
                     "e2e": {"value": cb["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(line), flush=True)
        return 0

    from nhans_b200.engine import Engine, PinnedArray, pack, KIND_GEMM, KIND_STFT, KIND_ISTFT
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set; stdout carries exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()                    # creates the communicator (the only NCCL use: barrier + max over ranks)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    eng = Engine(local_rank, variant, win_capacity=a.win_capacity)
    eng.load_weights(weights, wsrc)

    # ---- synthetic batch (distinct utterances per rank), staged in pinned host memory ----
    n_distinct = min(a.utts, 16)
    mixes_d = [synth.mixture(a.seconds, rank * 1000 + u) for u in range(n_distinct)]
    negs_d = [synth.noise_clip(rank * 1000 + u) for u in range(n_distinct)]
    mixes = [mixes_d[u % n_distinct] for u in range(a.utts)]
    negs = [negs_d[u % n_distinct] for u in range(a.utts)]
    mix, mo = pack(mixes)
    neg, no = pack(negs)
    oo = eng.output_offsets(mo)
    p_mix, p_neg = PinnedArray(mix.shape, np.int16), PinnedArray(neg.shape, np.int16)
    p_out = PinnedArray((int(oo[-1]),), np.int16)
    p_mix.array[:] = mix
    p_neg.array[:] = neg
    audio_s = float(mo[-1]) / FS

    def barrier():
        eng.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident: inputs already in HBM when the timed region starts ----
    eng.upload(p_mix.array, mo, None, None, p_neg.array, no)
    for _ in range(a.warmup):
        eng.run()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.profile_reset()
    eng.profile(True)
    eng.event_record(0)
    for _ in range(a.steps):
        eng.run()
    eng.event_record(1)
    barrier()
    ms_res = max_over_ranks(eng.event_elapsed_ms(0, 1))
    st = {k: eng.profile_get(k) for k in range(6)}
    layers = eng.profile_layers(0)
    eng.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    value = world * a.steps * audio_s / (ms_res / 1e3)

    # ---- end to end: host buffers in, int16 PCM out, copies inside the timed region ----
    for _ in range(max(1, a.warmup - 1)):
        eng.enhance_packed(p_mix.array, mo, None, None, p_neg.array, no, out_i16=p_out.array)
    barrier()
    eng.event_record(2)
    for _ in range(a.steps):
        eng.enhance_packed(p_mix.array, mo, None, None, p_neg.array, no, out_i16=p_out.array, sync=False)
    eng.event_record(3)
    barrier()
    ms_e2e = max_over_ranks(eng.event_elapsed_ms(2, 3))
    e2e = world * a.steps * audio_s / (ms_e2e / 1e3)

    if rank == 0:
        tf_peak, hbm_peak, which = peaks()
        g = st[KIND_GEMM]
        achieved = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = dict(base)
        line.update({
            "value": value, "ms_per_step": ms_res / a.steps, "dtype": "f16 operands, f32 accumulate",
            # whole-job figures: every rank copies its own shard and launches its own kernels
            "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": int(mix.nbytes + neg.nbytes) * world,
                    "d2h_bytes_per_step": int(p_out.array.nbytes) * world},
            "gpu_launches": int(st[5]["launches"]) * world,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_shift_kernel (tcgen05 implicit-GEMM conv layers)",
                         "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                         "peak_source": which + " bf16_tflops_sustained", "traffic": traffic,
                         "launches": g["launches"], "avg_launch_ms": g["ms"] / max(1, g["launches"]),
                         "share_of_step": g["ms"] / ms_res,
                         "algorithmic_flop_per_launch": g["flops"] / max(1, g["launches"])},
            "kernels": {
                "stft": {"launches": st[KIND_STFT]["launches"], "ms": st[KIND_STFT]["ms"],
                         "GBps": st[KIND_STFT]["bytes"] / max(st[KIND_STFT]["ms"], 1e-9) / 1e6, "frac_hbm": st[KIND_STFT]["bytes"] / max(st[KIND_STFT]["ms"], 1e-9) / 1e6 / hbm_peak},
                "istft": {"launches": st[KIND_ISTFT]["launches"], "ms": st[KIND_ISTFT]["ms"],
                          "GBps": st[KIND_ISTFT]["bytes"] / max(st[KIND_ISTFT]["ms"], 1e-9) / 1e6, "frac_hbm": st[KIND_ISTFT]["bytes"] / max(st[KIND_ISTFT]["ms"], 1e-9) / 1e6 / hbm_peak},
                "direct_conv": {"launches": st[3]["launches"], "ms": st[3]["ms"]},
                "other": {"launches": st[4]["launches"], "ms": st[4]["ms"]}},
            "layers": [{"name": l["name"], "ms_per_step": l["ms"] / a.steps, "TFLOPs": round(l["tflops"], 1)} for l in layers],
            "tensor_ceiling_audio_s_per_s": tf_peak * 1e3 / (GFLOP_PER_WINDOW * 100) * world,
        })
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(weights, variant, 4.0)    # BASELINE configs[0] in full: one 4 s utterance
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
