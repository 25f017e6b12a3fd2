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
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define NHANS_HD __host__ __device__ __forceinline__
#else
#define NHANS_HD inline
#endif

namespace nhans {

constexpr int kWalkSlots = 8;          // 8 x 64 fp32 columns = all 512 TMEM columns
constexpr int kWalkKH = 4;             // kernel rows (resblock1_x: 4 x 4, main.py:221-222)
constexpr int kWalkKW = 4;
constexpr int kWalkC = 64;             // channels in = channels out

struct WalkSeg {
  int slot;                            // first ring slot (accumulator column = 64 * slot)
  int bi;                              // first B block
  int nb;                              // blocks: N = 64 * nb
  int fresh;                           // 1: the slots are claimed by this step (first MMA must not accumulate)
};

struct WalkStep {
  int n_first, n_rest;                 // segment counts of the first K step / of every other K step
  WalkSeg first[4], rest[2];
  int n_claim;                         // jobs whose slot is claimed by this step (wait until the epilogue freed it)
  int claim_job[kWalkKH];
  int n_done;                          // jobs complete after this step (publish to the epilogue)
  int done_job[kWalkKH];
};

// Step for input row r of the tile with per-CTA sequence number seq; pt = rows of top padding.
NHANS_HD void walk_step(long long seq, int r, int H, int pt, WalkStep* s) {
  const int span = kWalkKH - 1 - pt;   // output rows above r that still receive input row r
  int o_lo = r - span, o_hi = r + pt;
  int bi0 = 0;
  if (o_lo < 0) { bi0 = -o_lo; o_lo = 0; }
  if (o_hi > H - 1) o_hi = H - 1;
  // output row o is first touched by input row max(o - pt, 0)
  const int fresh_lo = (r == 0) ? 0 : r + pt;          // rows >= fresh_lo are claimed now
  s->n_first = s->n_rest = s->n_claim = s->n_done = 0;
  const long long j0 = seq * H;
  for (int o = o_lo; o <= o_hi; ++o) {
    const long long J = j0 + o;
    const int slot = (int)(J % kWalkSlots);
    const int fresh = o >= fresh_lo ? 1 : 0;
    const int bi = bi0 + (o - o_lo);
    if (fresh) s->claim_job[s->n_claim++] = o;
    const bool wrap = (o != o_lo) && slot == 0;
    if (o == o_lo || wrap) {
      s->rest[s->n_rest++] = WalkSeg{slot, bi, 1, 0};
    } else {
      s->rest[s->n_rest - 1].nb++;
    }
    if (o == o_lo || wrap || s->first[s->n_first - 1].fresh != fresh) {
      s->first[s->n_first++] = WalkSeg{slot, bi, 1, fresh};
    } else {
      s->first[s->n_first - 1].nb++;
    }
  }
  // output row o is complete once input row min(o + span, H - 1) has been issued
  if (r < H - 1) {
    if (r - span >= 0) s->done_job[s->n_done++] = r - span;
  } else {
    for (int o = (r - span > 0 ? r - span : 0); o <= H - 1; ++o) s->done_job[s->n_done++] = o;
  }
}

}  // namespace nhans
