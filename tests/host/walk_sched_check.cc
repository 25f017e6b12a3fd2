// Host simulation of the row-walk issue schedule (nhans_b200/csrc/walk_sched.h): replays the MMA issuer's
// steps against a software model of the TMEM slot ring and checks that every output row receives exactly
// the (input row, kernel row) products of a 'SAME'-style convolution with `pt` rows of top padding, that a
// freshly claimed slot never accumulates, that the ring never wraps inside one MMA, and that the claim /
// publish order cannot deadlock against an in-order epilogue - also when, as in the kernel's issuer, the claims of
// step s + 1 are waited for BEFORE the jobs that finish in step s are published.
#include <cstdio>
#include <cstdlib>
#include <set>
#include <utility>
#include <vector>

#include "../../nhans_b200/csrc/walk_sched.h"

using namespace nhans;

#define CHECK(c, ...)                                   \
  do {                                                  \
    if (!(c)) {                                         \
      fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
      fprintf(stderr, __VA_ARGS__);                     \
      fprintf(stderr, "\n");                            \
      exit(1);                                          \
    }                                                   \
  } while (0)

struct Slot {
  long long job = -1;
  bool published = true;      // nothing to drain yet
  long long pub_step = -100;  // global step whose commit published the slot's last job
  std::set<std::pair<int, int>> terms;     // (input row, kernel row)
};

struct Model {
  std::vector<Slot> ring = std::vector<Slot>(kWalkSlots);
  long long seq;
  int r, H, pt;
  std::set<std::pair<int, int>> covered;

  void mma(int slot, int bi, int nb, int fresh, bool first) {
    CHECK(nb >= 1 && slot >= 0 && slot + nb <= kWalkSlots, "segment leaves the ring: slot %d nb %d", slot, nb);
    CHECK(bi >= 0 && bi + nb <= kWalkKH, "segment leaves the weight blocks");
    for (int b = 0; b < nb; ++b) {
      const int kh = kWalkKH - 1 - (bi + b);
      const int o = r - kh + pt;
      CHECK(o >= 0 && o < H, "product for output row %d outside the image (r %d kh %d)", o, r, kh);
      Slot& s = ring[slot + b];
      CHECK(s.job == seq * H + o, "slot %d holds job %lld, expected %lld", slot + b, s.job, seq * H + o);
      CHECK(!s.published, "MMA into a published slot");
      if (first) {
        if (fresh) CHECK(s.terms.empty(), "fresh slot already has terms");
        else CHECK(!s.terms.empty(), "accumulating into a slot that holds nothing");
        CHECK(covered.insert({slot + b, bi + b}).second, "block issued twice in one K step");
        s.terms.insert({r, kh});
      } else {
        CHECK(!fresh, "only the first K step may skip the accumulate");
        CHECK(s.terms.count({r, kh}) == 1, "rest list covers a block the first list did not");
        CHECK(covered.erase({slot + b, bi + b}) == 1, "rest list / first list mismatch");
      }
    }
  }
};

int main() {
  long long steps = 0, mmas_first = 0, mmas_rest = 0;
  for (int pt = 0; pt <= 2; ++pt)
    for (int H : {1, 2, 3, 4, 5, 7, 8, 9, 16, 35, 37}) {
      Model M;
      std::vector<Slot>& ring = M.ring;
      M.H = H; M.pt = pt;
      long long published_upto = -1;          // the epilogue drains jobs in order
      long long gstep = 0;                    // steps since the start of the launch
      for (long long seq = 0; seq < 9; ++seq) {
        for (int r = 0; r < H; ++r) {
          M.seq = seq; M.r = r;
          const WalkWin w = walk_window(seq, r, H, pt);
          CHECK(w.n >= 1 && w.n <= kWalkKH && w.n1 >= 1 && w.n1 <= w.n && w.n_fresh >= 0 && w.n_fresh <= w.n, "window");
          for (int c = 0; c < w.n_fresh; ++c) {
            const long long J = seq * H + w.o_lo + w.n - w.n_fresh + c;
            Slot& s = ring[J % kWalkSlots];
            // the issuer blocks until the epilogue has freed the slot: that needs the previous owner to be
            // published by an EARLIER step, otherwise the kernel deadlocks
            CHECK(s.published, "claim of job %lld: slot still owned by unpublished job %lld", J, s.job);
            CHECK(s.job < 0 || s.job == J - kWalkSlots, "slot reuse out of order");
            // conv_walk.cu takes this wait in the middle of the PREVIOUS step, ahead of that step's tmem_full commits: the
            // slot's last job must have been published by a step before that one
            CHECK(s.pub_step <= gstep - 2, "early claim of job %lld: slot published only in step %lld (now %lld)", J, s.pub_step, gstep);
            s.job = J;
            s.published = false;
            s.terms.clear();
          }
          int n_first = 0, n_rest = 0;
          walk_segments(w, true, [&](int slot, int bi, int nb, int fresh) { M.mma(slot, bi, nb, fresh, true); ++n_first; });
          walk_segments(w, false, [&](int slot, int bi, int nb, int fresh) { M.mma(slot, bi, nb, fresh, false); ++n_rest; });
          CHECK(M.covered.empty(), "first list covers blocks the rest list does not");
          CHECK(n_first >= 1 && n_first <= 4 && n_rest >= 1 && n_rest <= 2, "segment counts %d %d", n_first, n_rest);
          // every kernel row that sees input row r inside the image was issued
          for (int kh = 0; kh < kWalkKH; ++kh) {
            const int o = r - kh + pt;
            if (o < 0 || o >= H) continue;
            CHECK(ring[(seq * H + o) % kWalkSlots].terms.count({r, kh}) == 1, "missing product r %d kh %d", r, kh);
          }
          for (int d = 0; d < w.n_done; ++d) {
            const int o = w.done_lo + d;
            const long long J = seq * H + o;
            Slot& s = ring[J % kWalkSlots];
            CHECK(s.job == J && !s.published, "publish of a job that is not live");
            std::set<std::pair<int, int>> want;
            for (int kh = 0; kh < kWalkKH; ++kh) {
              const int rr = o + kh - pt;
              if (rr >= 0 && rr < H) want.insert({rr, kh});
            }
            CHECK(s.terms == want, "pt %d H %d: job %lld published with %zu of %zu products", pt, H, J, s.terms.size(), want.size());
            CHECK(J == published_upto + 1, "jobs published out of order (%lld after %lld)", J, published_upto);
            published_upto = J;
            s.published = true;
            s.pub_step = gstep;
          }
          ++steps;
          ++gstep;
          mmas_first += n_first;
          mmas_rest += n_rest;
        }
        CHECK(published_upto == (seq + 1) * H - 1, "tile %lld left unpublished rows", seq);
      }
    }
  printf("walk schedule ok: %lld steps, %.2f / %.2f MMAs per first / other K step\n", steps, (double)mmas_first / steps,
         (double)mmas_rest / steps);
  return 0;
}
