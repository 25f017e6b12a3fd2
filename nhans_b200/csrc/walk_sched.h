// Issue schedule of the row-walk convolution kernel (conv_walk.cu), shared with the host-side simulation in
// tests/host/walk_sched_check.cc (plain C++, no CUDA types).
//
// A CTA owns a strip of 128 GEMM rows (pixels (unit, x) of one image-row plane) and walks down the H image
// rows of the strip.  Input row r is loaded ONCE (one A slab) and multiplied by the weights of all KH kernel
// rows at once: B block bi holds kernel row kh = KH - 1 - bi, so the N = 64 * KH accumulator columns of one MMA
// are the output rows o = r - (KH - 1 - pt) + bi, bi = 0 .. KH - 1 (ascending o).  Output rows live in a ring of
// kWalkSlots TMEM slots of 64 columns: job J = tile_seq * H + o uses slot J % kWalkSlots.  A step therefore
// issues MMAs over a sliding window of up to KH consecutive jobs; the window is cut into segments where the
// ring wraps, at the image edges (rows outside [0, H) get no MMA at all) and - for the first K step only -
// where freshly claimed slots (accumulate = 0) meet slots that already hold partial sums.
//
// Everything is closed-form arithmetic on a handful of integers (no arrays): the issuer thread keeps the whole
// step in registers.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define NHANS_WALK_HD __host__ __device__ __forceinline__
#else
#define NHANS_WALK_HD inline
#endif

namespace nhans {

constexpr int kWalkSlots = 8;          // 8 x 64 fp32 columns = all 512 TMEM columns
constexpr int kWalkKH = 4;             // kernel rows (resblock1_x: 4 x 4, main.py:221-222)
constexpr int kWalkKW = 4;
constexpr int kWalkC = 64;             // channels in = channels out

struct WalkWin {
  int o_lo;                            // first output row of the window
  int n;                               // rows in the window (1 .. KH)
  int slot0;                           // ring slot of o_lo
  int bi0;                             // weight block of o_lo
  int n1;                              // rows before the ring wraps (== n when it does not)
  int n_fresh;                         // the LAST n_fresh rows of the window are claimed by this step
  int done_lo, n_done;                 // output rows complete after this step: [done_lo, done_lo + n_done)
};

// Window of input row r of the tile with per-CTA sequence number seq; pt = rows of top padding.
NHANS_WALK_HD WalkWin walk_window(long long seq, int r, int H, int pt) {
  WalkWin w;
  const int span = kWalkKH - 1 - pt;   // output rows above r that still receive input row r
  const int lo = r - span;
  w.o_lo = lo > 0 ? lo : 0;
  w.bi0 = w.o_lo - lo;
  const int o_hi = r + pt < H - 1 ? r + pt : H - 1;
  w.n = o_hi - w.o_lo + 1;
  w.slot0 = (int)((seq * H + w.o_lo) & (kWalkSlots - 1));
  const int room = kWalkSlots - w.slot0;
  w.n1 = w.n < room ? w.n : room;
  // output row o is first touched by input row max(o - pt, 0)
  int fresh_lo = (r == 0) ? 0 : r + pt;
  if (fresh_lo < w.o_lo) fresh_lo = w.o_lo;
  w.n_fresh = o_hi >= fresh_lo ? o_hi - fresh_lo + 1 : 0;
  // output row o is complete once input row min(o + span, H - 1) has been issued
  if (r < H - 1) {
    w.done_lo = lo;
    w.n_done = lo >= 0 ? 1 : 0;
  } else {
    w.done_lo = w.o_lo;
    w.n_done = H - w.o_lo;
  }
  return w;
}

// Calls f(slot, bi, nb, fresh) for every MMA of one K step: `first` = the step's first K step (fresh slots
// must not accumulate and are therefore issued apart from the ones that hold partial sums).
template <class F>
NHANS_WALK_HD void walk_segments(const WalkWin& w, bool first, F&& f) {
  auto part = [&](int a, int b, int fresh) {          // window rows [a, b), cut where the ring wraps
    const int m = b < w.n1 ? b : w.n1;
    if (m > a) f(w.slot0 + a, w.bi0 + a, m - a, fresh);
    const int s = a > w.n1 ? a : w.n1;
    if (b > s) f(w.slot0 + s - kWalkSlots, w.bi0 + s, b - s, fresh);
  };
  if (first) {
    const int n_old = w.n - w.n_fresh;
    part(0, n_old, 0);
    part(n_old, w.n, 1);
  } else {
    part(0, w.n, 0);
  }
}

}  // namespace nhans
