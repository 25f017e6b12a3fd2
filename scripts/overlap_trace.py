"""Device timeline of pipelined batches (NHANS_DEBUG_TIMELINE=1): shows the H2D copies of batch i + 1 and the D2H copies of
batch i - 1 running while batch i computes.   python scripts/overlap_trace.py [batches] [utts] [seconds] > profiles/r02_overlap_trace.txt"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["NHANS_DEBUG_TIMELINE"] = "1"
from nhans_b200 import synth, weights as W      # noqa: E402
from nhans_b200.engine import Engine            # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 6
utts = int(sys.argv[2]) if len(sys.argv) > 2 else 64
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
eng = Engine(0, 0)
eng.load_weights(W.seeded_init(0, 0))
mixes = [synth.mixture(secs, u % 8) for u in range(utts)]
negs = [synth.noise_clip(u % 8) for u in range(utts)]
eng.enhance(mixes, None, negs, want_f32=False)            # warm-up: allocations
prev = None
for k in range(nb):                                       # two batches in flight, like runtime.MultiGpu
    cur = eng.submit(mixes, None, negs, want_f32=False)
    if prev is not None:
        eng.collect(prev, newer_in_flight=True)
    prev = cur
eng.collect(prev)
out = np.zeros(6 * nb, np.float64)
n = eng.lib.nhans_debug_timeline(eng.h, out.ctypes.data_as(ctypes.c_void_p), nb)
t = out[:6 * n].reshape(n, 6)
t = t - t[0, 0]
print("device timeline of %d pipelined nhans_enhance_batch calls, %d x %.0f s utterances each (%.1f MB in, %.1f MB out per batch); ms since the first copy"
      % (n, utts, secs, (sum(len(m) for m in mixes) + sum(len(c) for c in negs)) * 2 / 1e6, sum(len(m) for m in mixes) * 2 / 1e6))
print("%5s  %-21s %-21s %-21s  %s" % ("batch", "H2D (copy-in stream)", "compute stream", "D2H (copy-out stream)", "copies hidden behind compute of"))
for i in range(n):
    h0, h1, c0, c1, d0, d1 = t[i]
    hid = []
    for j in range(n):
        if j != i and t[j][2] <= h0 and h1 <= t[j][3]:
            hid.append("H2D under batch %d" % j)
        if j != i and t[j][2] <= d0 and d1 <= t[j][3]:
            hid.append("D2H under batch %d" % j)
    print("%5d  %9.2f - %9.2f  %9.2f - %9.2f  %9.2f - %9.2f  %s" % (i, h0, h1, c0, c1, d0, d1, ", ".join(hid) or "-"))
gaps = [t[i + 1][2] - t[i][3] for i in range(n - 1)]
print("idle gaps of the compute stream between consecutive batches (ms):", " ".join("%.3f" % g for g in gaps))
print("compute per batch (ms):", " ".join("%.1f" % (t[i][3] - t[i][2]) for i in range(n)), "| copies per batch (ms): H2D",
      " ".join("%.2f" % (t[i][1] - t[i][0]) for i in range(n)), "D2H", " ".join("%.2f" % (t[i][5] - t[i][4]) for i in range(n)))
eng.close()
