// FFT-400 building blocks shared by the STFT / iSTFT kernels (dsp.cu) and the host-side self check
// (tests/host/fft_check.cu): complex DFT-200 = 8 x 25 (radix-8, then 5 x 5) and the real <-> packed-complex
// pre/post passes.  All functions are __host__ __device__ so the arithmetic can be verified without a GPU.
#pragma once

#include <cuda_runtime.h>

#define NHANS_HD __host__ __device__ __forceinline__

namespace nhans {
namespace fft {

// Complex helpers.  On the device they are Blackwell's packed fp32x2 instructions (add / sub / mul / fma.rn.f32x2 ->
// FADD2 / FMUL2 / FFMA2 in SASS: one issue slot for both components; ptxas folds component swaps, negations and
// scalar broadcasts into the operand selectors, so multiplying by +-i or by a real constant costs nothing extra).
// The STFT kernels are issue bound, and half of their instructions were scalar fp32 adds / multiplies on (re, im)
// pairs.  The host versions (tests/host/fft_check.cu) are the plain scalar formulas.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned long long f2_bits(float2 a) { return *reinterpret_cast<unsigned long long*>(&a); }
__device__ __forceinline__ float2 bits_f2(unsigned long long b) { return *reinterpret_cast<float2*>(&b); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(r);
}
__device__ __forceinline__ float2 cmul2(float2 a, float2 b) {          // component-wise product
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)));
  return bits_f2(r);
}
__device__ __forceinline__ float2 cfma2(float2 a, float2 b, float2 c) { // component-wise a * b + c
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
  return bits_f2(r);
}
#else
inline float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
inline float2 cmul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 cfma2(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
#endif
NHANS_HD float2 cscale(float2 a, float s) { return cmul2(a, make_float2(s, s)); }                 // a * s
NHANS_HD float2 caxpy(float2 a, float s, float2 c) { return cfma2(a, make_float2(s, s), c); }     // a * s + c
// complex product a * b = a.x * b + a.y * (i b)
NHANS_HD float2 cmul(float2 a, float2 b) {
  return cfma2(make_float2(a.x, a.x), b, cmul2(make_float2(a.y, a.y), make_float2(-b.y, b.x)));
}
NHANS_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
NHANS_HD float2 rot(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }
template <bool INV>
NHANS_HD float2 tw(float2 w) { return INV ? cconj(w) : w; }

template <bool INV>
NHANS_HD void dft4(float2 c0, float2 c1, float2 c2, float2 c3, float2& x0, float2& x1, float2& x2, float2& x3) {
  float2 e0 = cadd(c0, c2), e1 = csub(c0, c2), o0 = cadd(c1, c3), o1 = rot<INV>(csub(c1, c3));
  x0 = cadd(e0, o0); x2 = csub(e0, o0); x1 = cadd(e1, o1); x3 = csub(e1, o1);
}

template <bool INV>
NHANS_HD void dft8(float2 (&v)[8]) {
  const float r = 0.70710678118654752f;
  float2 a[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { a[j] = cadd(v[j], v[j + 4]); b[j] = csub(v[j], v[j + 4]); }
  // b[j] *= W8^j  (forward W8 = e^{-i pi/4})
  // W8 b = (b + rot(b)) / sqrt 2,  W8^3 b = (rot(b) - b) / sqrt 2   (rot = multiplication by -+i)
  b[1] = cscale(cadd(b[1], rot<INV>(b[1])), r);
  b[2] = rot<INV>(b[2]);
  b[3] = cscale(csub(rot<INV>(b[3]), b[3]), r);
  dft4<INV>(a[0], a[1], a[2], a[3], v[0], v[2], v[4], v[6]);
  dft4<INV>(b[0], b[1], b[2], b[3], v[1], v[3], v[5], v[7]);
}

template <bool INV>
NHANS_HD void dft5(float2 x0, float2 x1, float2 x2, float2 x3, float2 x4, float2& y0, float2& y1, float2& y2, float2& y3, float2& y4) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
  const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
  float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
  y0 = cadd(cadd(x0, t1), t2);
  float2 m1 = caxpy(t2, c2, caxpy(t1, c1, x0));
  float2 m2 = caxpy(t2, c1, caxpy(t1, c2, x0));
  float2 u1 = rot<INV>(caxpy(t4, s2, cscale(t3, s1)));
  float2 u2 = rot<INV>(caxpy(t4, -s1, cscale(t3, s2)));
  y1 = cadd(m1, u1); y4 = csub(m1, u1); y2 = cadd(m2, u2); y3 = csub(m2, u2);
}

// Complex DFT-200 (unnormalised) of sequence f: in[f][200] -> out[f][200] through tmp[f][200].
// Index split: input m = 25 m1 + m2, output k = k1 + 8 k2.
// step A (one call per (f, m2)): radix-8 butterfly over m1, then twiddle W200^{m2 k1}
template <bool INV>
NHANS_HD void fft200_step_a(const float2* in, float2* tmp, int f, int m2, const float2* tw200) {
  float2 v[8];
#pragma unroll
  for (int m1 = 0; m1 < 8; ++m1) v[m1] = in[f * 200 + 25 * m1 + m2];
  dft8<INV>(v);
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) tmp[f * 200 + k1 * 25 + m2] = cmul(v[k1], tw<INV>(tw200[m2 * k1]));
}
// step B (one call per (f, k1)): DFT-25 = 5 x 5 over m2
template <bool INV>
NHANS_HD void fft200_step_b(const float2* tmp, float2* out, int f, int k1, const float2* tw25) {
  const float2* y = tmp + f * 200 + k1 * 25;
  float2 g[5][5];                       // g[b][c] = sum_a y[5a + b] W5^{a c}
#pragma unroll
  for (int b = 0; b < 5; ++b) {
    dft5<INV>(y[b], y[5 + b], y[10 + b], y[15 + b], y[20 + b], g[b][0], g[b][1], g[b][2], g[b][3], g[b][4]);
#pragma unroll
    for (int c = 1; c < 5; ++c)
      if (b > 0) g[b][c] = cmul(g[b][c], tw<INV>(tw25[b * c]));
  }
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    float2 z0, z1, z2, z3, z4;          // k2 = c + 5 d
    dft5<INV>(g[0][c], g[1][c], g[2][c], g[3][c], g[4][c], z0, z1, z2, z3, z4);
    float2* o = out + f * 200 + k1;
    o[8 * (c)] = z0; o[8 * (c + 5)] = z1; o[8 * (c + 10)] = z2; o[8 * (c + 15)] = z3; o[8 * (c + 20)] = z4;
  }
}

// The same DFT-25 as two radix-5 passes with 40 independent tasks per frame each (the kernels use these: one
// task per thread keeps all lanes busy, the monolithic step B has only 8 tasks per frame).
// pass B1 (one call per (f, k1, b)): g[b][c] = W25^{bc} sum_a y[5a + b] W5^{ac}
template <bool INV>
NHANS_HD void fft200_step_b1(const float2* tmp, float2* g, int f, int k1, int b, const float2* tw25) {
  const float2* y = tmp + f * 200 + k1 * 25;
  float2 o0, o1, o2, o3, o4;
  dft5<INV>(y[b], y[5 + b], y[10 + b], y[15 + b], y[20 + b], o0, o1, o2, o3, o4);
  float2* gg = g + f * 200 + k1 * 25 + b * 5;
  gg[0] = o0;
  gg[1] = cmul(o1, tw<INV>(tw25[b]));
  gg[2] = cmul(o2, tw<INV>(tw25[2 * b]));
  gg[3] = cmul(o3, tw<INV>(tw25[3 * b]));
  gg[4] = cmul(o4, tw<INV>(tw25[4 * b]));
}
// pass B2 (one call per (f, k1, c)): Z[k1 + 8 (c + 5 d)] = sum_b g[b][c] W5^{bd}
template <bool INV>
NHANS_HD void fft200_step_b2(const float2* g, float2* out, int f, int k1, int c) {
  const float2* gg = g + f * 200 + k1 * 25 + c;
  float2 z0, z1, z2, z3, z4;
  dft5<INV>(gg[0], gg[5], gg[10], gg[15], gg[20], z0, z1, z2, z3, z4);
  float2* o = out + f * 200 + k1;
  o[8 * c] = z0; o[8 * (c + 5)] = z1; o[8 * (c + 10)] = z2; o[8 * (c + 15)] = z3; o[8 * (c + 20)] = z4;
}

// DFT-25 of 25 values held in registers, in place (two radix-5 stages, compile-time twiddles): y[5 a + b] in, the
// output for k2 = c + 5 d is left at y[5 c + d].
template <bool INV>
NHANS_HD void dft25(float2 (&y)[25]) {
  const float2 W[17] = {  // e^{-2 pi i j / 25}, j = 0 .. 16
      {1.000000000e+00f, -0.000000000e+00f}, {9.685831611e-01f, -2.486898872e-01f}, {8.763066800e-01f, -4.817536741e-01f},
      {7.289686274e-01f, -6.845471059e-01f}, {5.358267950e-01f, -8.443279255e-01f}, {3.090169944e-01f, -9.510565163e-01f},
      {6.279051953e-02f, -9.980267284e-01f}, {-1.873813146e-01f, -9.822872507e-01f}, {-4.257792916e-01f, -9.048270525e-01f},
      {-6.374239897e-01f, -7.705132428e-01f}, {-8.090169944e-01f, -5.877852523e-01f}, {-9.297764859e-01f, -3.681245527e-01f},
      {-9.921147013e-01f, -1.253332336e-01f}, {-9.921147013e-01f, 1.253332336e-01f}, {-9.297764859e-01f, 3.681245527e-01f},
      {-8.090169944e-01f, 5.877852523e-01f}, {-6.374239897e-01f, 7.705132428e-01f}};
#pragma unroll
  for (int b = 0; b < 5; ++b) {       // g[b][c] = W25^{bc} sum_a y[5a + b] W5^{ac}, stored at y[5 c + b]
    float2 o0, o1, o2, o3, o4;
    dft5<INV>(y[b], y[5 + b], y[10 + b], y[15 + b], y[20 + b], o0, o1, o2, o3, o4);
    y[b] = o0;
    y[5 + b] = b ? cmul(o1, tw<INV>(W[b])) : o1;
    y[10 + b] = b ? cmul(o2, tw<INV>(W[2 * b])) : o2;
    y[15 + b] = b ? cmul(o3, tw<INV>(W[3 * b])) : o3;
    y[20 + b] = b ? cmul(o4, tw<INV>(W[4 * b])) : o4;
  }
#pragma unroll
  for (int c = 0; c < 5; ++c)         // Z[c + 5 d] = sum_b g[b][c] W5^{bd}, stored at y[5 c + d]
    dft5<INV>(y[5 * c], y[5 * c + 1], y[5 * c + 2], y[5 * c + 3], y[5 * c + 4], y[5 * c], y[5 * c + 1], y[5 * c + 2], y[5 * c + 3], y[5 * c + 4]);
}

// rfft-400 bin k (0..200) from the DFT-200 Z of the packed sequence z[m] = x[2m] + i x[2m+1]:
// X[k] = (Z[k] + conj Z[200-k]) / 2 - (i/2) e^{-2 pi i k / 400} (Z[k] - conj Z[200-k])
NHANS_HD float2 rfft_post(const float2* Z, int k, const float2* tw400) {
  const float2 zk = Z[k == 200 ? 0 : k];
  const float2 zc = cconj(Z[(200 - k) % 200]);
  const float2 e = cadd(zk, zc), d = csub(zk, zc);
  const float2 wd = cmul(tw400[k], d);
  return make_float2(0.5f * (e.x + wd.y), 0.5f * (e.y - wd.x));
}
// inverse: Z[k] = (S[k] + conj S[200-k]) + i e^{+2 pi i k / 400} (S[k] - conj S[200-k]), k = 0..199;
// the unnormalised inverse DFT-200 of Z is 400 * (x[2m] + i x[2m+1])
NHANS_HD float2 irfft_pre(const float2* S, int k, const float2* tw400) {
  const float2 sk = S[k];
  const float2 sc = cconj(S[200 - k]);
  const float2 e = cadd(sk, sc), d = csub(sk, sc);
  const float2 wd = cmul(cconj(tw400[k]), d);
  return make_float2(e.x - wd.y, e.y + wd.x);
}

}  // namespace fft
}  // namespace nhans
