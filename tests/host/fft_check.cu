// Host-side self check of nhans_b200/csrc/fft400.cuh: runs the exact __host__ __device__ butterflies the
// STFT / iSTFT kernels use against a naive O(N^2) double-precision DFT.  Built and run by
// tests/test_host_fft.py (no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../nhans_b200/csrc/fft400.cuh"

using namespace nhans::fft;

int main() {
  const double kPi = 3.14159265358979323846;
  std::vector<float2> t200(200), t400(201), t25(25);
  for (int i = 0; i < 200; ++i) t200[i] = make_float2((float)cos(2 * kPi * i / 200), (float)-sin(2 * kPi * i / 200));
  for (int i = 0; i < 201; ++i) t400[i] = make_float2((float)cos(2 * kPi * i / 400), (float)-sin(2 * kPi * i / 400));
  for (int i = 0; i < 25; ++i) t25[i] = make_float2((float)cos(2 * kPi * i / 25), (float)-sin(2 * kPi * i / 25));
  srand(1);
  const int NF = 3;
  std::vector<float> x(NF * 400);
  for (auto& v : x) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  std::vector<float2> a(NF * 200), b(NF * 200), c(NF * 200);
  for (int f = 0; f < NF; ++f)
    for (int m = 0; m < 200; ++m) a[f * 200 + m] = make_float2(x[f * 400 + 2 * m], x[f * 400 + 2 * m + 1]);
  for (int f = 0; f < NF; ++f) for (int m2 = 0; m2 < 25; ++m2) fft200_step_a<false>(a.data(), b.data(), f, m2, t200.data());
  for (int f = 0; f < NF; ++f) for (int k1 = 0; k1 < 8; ++k1) fft200_step_b<false>(b.data(), c.data(), f, k1, t25.data());
  {  // the split radix-5 passes give the same DFT-25 as the monolithic step B
    std::vector<float2> g(NF * 200), c2(NF * 200);
    for (int f = 0; f < NF; ++f) for (int k1 = 0; k1 < 8; ++k1) for (int bb = 0; bb < 5; ++bb) fft200_step_b1<false>(b.data(), g.data(), f, k1, bb, t25.data());
    for (int f = 0; f < NF; ++f) for (int k1 = 0; k1 < 8; ++k1) for (int cc = 0; cc < 5; ++cc) fft200_step_b2<false>(g.data(), c2.data(), f, k1, cc);
    double d = 0;
    for (int i = 0; i < NF * 200; ++i) d = fmax(d, fmax(fabs(c2[i].x - c[i].x), fabs(c2[i].y - c[i].y)));
    printf("split step B max diff %.3e\n", d);
    if (d > 1e-4) return 2;
    c = c2;
  }
  {  // the in-register DFT-25 of the warp-per-frame kernels (dft25, output for k2 = c + 5 d left at y[5 c + d])
    double d = 0;
    for (int inv = 0; inv < 2; ++inv)
      for (int f = 0; f < NF; ++f)
        for (int k1 = 0; k1 < 8; ++k1) {
          float2 y[25], ref[200] = {};
          for (int i = 0; i < 25; ++i) y[i] = b[f * 200 + k1 * 25 + i];
          if (inv) { dft25<true>(y); fft200_step_b<true>(b.data() + f * 200, ref, 0, k1, t25.data()); }
          else { dft25<false>(y); fft200_step_b<false>(b.data() + f * 200, ref, 0, k1, t25.data()); }
          for (int c5 = 0; c5 < 5; ++c5)
            for (int d5 = 0; d5 < 5; ++d5) {
              const float2 want = ref[k1 + 8 * (c5 + 5 * d5)];
              d = fmax(d, fmax(fabs(y[5 * c5 + d5].x - want.x), fabs(y[5 * c5 + d5].y - want.y)));
            }
        }
    printf("in-register dft25 max diff %.3e\n", d);
    if (d > 1e-4) return 3;
  }
  double max_err = 0, max_mag = 0;
  std::vector<float2> S(NF * 201);
  for (int f = 0; f < NF; ++f)
    for (int k = 0; k <= 200; ++k) {
      double re = 0, im = 0;
      for (int n = 0; n < 400; ++n) { re += x[f * 400 + n] * cos(2 * kPi * k * n / 400); im -= x[f * 400 + n] * sin(2 * kPi * k * n / 400); }
      float2 X = rfft_post(c.data() + f * 200, k, t400.data());
      S[f * 201 + k] = X;
      max_err = fmax(max_err, fmax(fabs(X.x - re), fabs(X.y - im)));
      max_mag = fmax(max_mag, sqrt(re * re + im * im));
    }
  printf("rfft400 max_abs_err %.3e max_mag %.3e\n", max_err, max_mag);
  // inverse round trip
  for (int f = 0; f < NF; ++f) { S[f * 201].y = 0; S[f * 201 + 200].y = 0; }
  for (int f = 0; f < NF; ++f) for (int k = 0; k < 200; ++k) a[f * 200 + k] = irfft_pre(S.data() + f * 201, k, t400.data());
  for (int f = 0; f < NF; ++f) for (int m2 = 0; m2 < 25; ++m2) fft200_step_a<true>(a.data(), b.data(), f, m2, t200.data());
  for (int f = 0; f < NF; ++f) for (int k1 = 0; k1 < 8; ++k1) fft200_step_b<true>(b.data(), c.data(), f, k1, t25.data());
  double rt = 0;
  for (int f = 0; f < NF; ++f)
    for (int m = 0; m < 200; ++m) {
      rt = fmax(rt, fabs(c[f * 200 + m].x / 400.0 - x[f * 400 + 2 * m]));
      rt = fmax(rt, fabs(c[f * 200 + m].y / 400.0 - x[f * 400 + 2 * m + 1]));
    }
  printf("irfft400 roundtrip max_abs_err %.3e\n", rt);
  return (max_err < 2e-4 * fmax(1.0, max_mag / 10) && rt < 1e-5) ? 0 : 1;
}
