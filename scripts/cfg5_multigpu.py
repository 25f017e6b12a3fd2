"""BASELINE config 5 through the repo's own multi-GPU runtime (nhans_b200.runtime.MultiGpu): U x S-second clips dealt
dynamically in chunks to one engine per GPU, no collective on the data path.  Prints one JSON record with the
whole-job throughput (host wall clock around MultiGpu.enhance: packing, H2D, compute, D2H, unpacking), the per-GPU
chunk counts / busy times, and a bit-exact comparison of a sample of utterances with a single-engine run.

  python scripts/cfg5_multigpu.py --gpus 8 --utts 8192 --seconds 10 --chunk 64 [--repeat 2]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nhans_b200 import synth, weights as W          # noqa: E402
from nhans_b200.engine import Engine                # noqa: E402
from nhans_b200.runtime import MultiGpu             # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=8)
ap.add_argument("--utts", type=int, default=8192)
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--chunk", type=int, default=32)
ap.add_argument("--repeat", type=int, default=2)
ap.add_argument("--check", type=int, default=6, help="utterances compared bit for bit with a single-engine run")
a = ap.parse_args()

w = W.seeded_init(W.SELECTIVE_NOISE, 0)
n_distinct = min(a.utts, 32)
base_m = [synth.mixture(a.seconds, 7000 + u) for u in range(n_distinct)]
base_n = [synth.noise_clip(7000 + u) for u in range(n_distinct)]
mixes = [base_m[u % n_distinct] for u in range(a.utts)]
negs = [base_n[u % n_distinct] for u in range(a.utts)]
audio_s = sum(len(m) for m in mixes) / 16000.0

mg = MultiGpu(list(range(a.gpus)), W.SELECTIVE_NOISE, w)
runs = []
try:
    warm = min(a.utts, 2 * a.chunk * a.gpus)
    mg.enhance(mixes[:warm], None, negs[:warm], chunk_utts=a.chunk, want_f32=False)      # warm-up: allocations, first launches
    for _ in range(a.repeat):
        t = time.perf_counter()
        out = mg.enhance(mixes, None, negs, chunk_utts=a.chunk, want_f32=False)
        dt = time.perf_counter() - t
        st = mg.last_stats
        runs.append(dict(wall_s=dt, audio_s_per_s=audio_s / dt, per_gpu=st["per_gpu"], n_chunks=st["n_chunks"]))
finally:
    mg.close()

# bit-exact against one engine processing the same utterances alone (batch invariance of the whole path)
eng = Engine(0, W.SELECTIVE_NOISE)
eng.load_weights(w)
pick = sorted(set(int(x) for x in np.linspace(0, a.utts - 1, a.check)))
ref = eng.enhance([mixes[i] for i in pick], None, [negs[i] for i in pick], want_f32=False)
eng.close()
same = all(np.array_equal(ref["i16"][j], out["i16"][i]) for j, i in enumerate(pick))

best = max(runs, key=lambda r: r["audio_s_per_s"])
busy = [g["busy_s"] for g in best["per_gpu"]]
print(json.dumps(dict(workload="BASELINE config 5 through runtime.MultiGpu: %d x %.0f s clips + --neg, chunks of %d utterances dealt "
                               "dynamically to %d GPUs, int16 PCM in / out through pinned double-buffered staging"
                               % (a.utts, a.seconds, a.chunk, a.gpus),
                      n_gpus=a.gpus, utterances=a.utts, seconds=a.seconds, audio_seconds=audio_s, chunk_utts=a.chunk,
                      audio_s_per_s=best["audio_s_per_s"], wall_s=best["wall_s"], runs=[r["audio_s_per_s"] for r in runs],
                      per_gpu=best["per_gpu"], busy_spread=(max(busy) - min(busy)) / max(busy) if busy else None,
                      bit_identical_to_single_engine=bool(same), checked_utterances=pick,
                      timing="host wall clock around MultiGpu.enhance (everything inside: packing, copies, kernels, unpacking)")))
