// Host simulation of the row-walk issue schedule (nhans_b200/csrc/walk_sched.h): replays the MMA issuer's
// steps against a software model of the TMEM slot ring and checks that every output row receives exactly
// the (input row, kernel row) products of a 'SAME'-style convolution with `pt` rows of top padding, that a
// freshly claimed slot never accumulates, that the ring never wraps inside one MMA, and that the claim /
// publish order cannot deadlock against an in-order epilogue.
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
  std::set<std::pair<int, int>> terms;     // (input row, kernel row)
};

static void apply(const WalkSeg* segs, int n, bool first, long long seq, int r, int H, int pt, std::vector<Slot>& ring,
                  std::set<std::pair<int, int>>* covered) {
  for (int i = 0; i < n; ++i) {
    const WalkSeg& sg = segs[i];
    CHECK(sg.nb >= 1 && sg.slot >= 0 && sg.slot + sg.nb <= kWalkSlots, "segment leaves the ring: slot %d nb %d", sg.slot, sg.nb);
    CHECK(sg.bi >= 0 && sg.bi + sg.nb <= kWalkKH, "segment leaves the weight blocks");
    for (int b = 0; b < sg.nb; ++b) {
      const int kh = kWalkKH - 1 - (sg.bi + b);
      const int o = r - kh + pt;
      CHECK(o >= 0 && o < H, "product for output row %d outside the image (r %d kh %d)", o, r, kh);
      Slot& s = ring[sg.slot + b];
      CHECK(s.job == seq * H + o, "slot %d holds job %lld, expected %lld", sg.slot + b, s.job, seq * H + o);
      CHECK(!s.published, "MMA into a published slot");
      if (first) {
        if (sg.fresh) CHECK(s.terms.empty(), "fresh slot already has terms");
        else CHECK(!s.terms.empty(), "accumulating into a slot that holds nothing");
        CHECK(covered->insert({sg.slot + b, sg.bi + b}).second, "block issued twice in one K step");
        s.terms.insert({r, kh});
      } else {
        CHECK(s.terms.count({r, kh}) == 1, "rest list covers a block the first list did not");
        CHECK(covered->erase({sg.slot + b, sg.bi + b}) == 1, "rest list / first list mismatch");
      }
    }
  }
}

int main() {
  long long steps = 0, mmas_first = 0, mmas_rest = 0;
  for (int pt = 0; pt <= 2; ++pt)
    for (int H : {1, 2, 3, 4, 5, 7, 8, 9, 16, 35, 37}) {
      std::vector<Slot> ring(kWalkSlots);
      long long published_upto = -1;          // the epilogue drains jobs in order
      for (long long seq = 0; seq < 9; ++seq) {
        for (int r = 0; r < H; ++r) {
          WalkStep st;
          walk_step(seq, r, H, pt, &st);
          CHECK(st.n_first >= 1 && st.n_first <= 4 && st.n_rest >= 1 && st.n_rest <= 2, "segment counts %d %d", st.n_first, st.n_rest);
          for (int c = 0; c < st.n_claim; ++c) {
            const long long J = seq * H + st.claim_job[c];
            Slot& s = ring[J % kWalkSlots];
            // the issuer blocks until the epilogue has freed the slot: that needs the previous owner to be
            // published by an EARLIER step, otherwise the kernel deadlocks
            CHECK(s.published, "claim of job %lld: slot still owned by unpublished job %lld", J, s.job);
            CHECK(s.job < 0 || s.job == J - kWalkSlots, "slot reuse out of order");
            s.job = J;
            s.published = false;
            s.terms.clear();
          }
          std::set<std::pair<int, int>> covered;
          apply(st.first, st.n_first, true, seq, r, H, pt, ring, &covered);
          apply(st.rest, st.n_rest, false, seq, r, H, pt, ring, &covered);
          CHECK(covered.empty(), "first list covers blocks the rest list does not");
          // every kernel row that sees input row r inside the image was issued
          for (int kh = 0; kh < kWalkKH; ++kh) {
            const int o = r - kh + pt;
            if (o < 0 || o >= H) continue;
            CHECK(ring[(seq * H + o) % kWalkSlots].terms.count({r, kh}) == 1, "missing product r %d kh %d", r, kh);
          }
          for (int d = 0; d < st.n_done; ++d) {
            const int o = st.done_job[d];
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
          }
          ++steps;
          mmas_first += st.n_first;
          mmas_rest += st.n_rest;
        }
        CHECK(published_upto == (seq + 1) * H - 1, "tile %lld left unpublished rows", seq);
      }
    }
  printf("walk schedule ok: %lld steps, %.2f / %.2f MMAs per first / other K step\n", steps, (double)mmas_first / steps,
         (double)mmas_rest / steps);
  return 0;
}
