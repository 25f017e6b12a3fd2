// STFT front end and inverse-STFT overlap-add back end (HBM-bound CUDA-core kernels).
//
// Forward (N_HANS___Selective_Noise/apply.py:142-163, 368-375; reader.py:334-350):
//   peak normalise in float64 -> float32, frames x[160 t : 160 t + 400] * periodic Hann, rfft-400,
//   log(|X| + 1e-5) and angle(X), fused in one kernel (stft_kernel).
// Inverse (SN/apply.py:189-204 = tf.signal.inverse_stft with inverse_stft_window_fn(160, hann)):
//   exp(logmag) e^{j phase} -> irfft-400 -> synthesis window -> 3-frame gather overlap-add ->
//   float32 and/or int16 PCM, fused in one kernel (istft_kernel), no atomics.
//
// FFT-400 is not a power of two: the real transform is done as one complex FFT-200 on
// packed even/odd samples, 200 = 8 x 25 (radix-8 butterflies, then 5 x 5), all in shared memory and
// registers; each CTA stages the contiguous sample span of its frames once.
#include "kernels.h"
#include "fft400.cuh"

#include <math.h>
#include <vector>

namespace nhans {

namespace {

constexpr int kWin = 400;
constexpr int kHop = 160;
constexpr int kBinsD = 201;
constexpr int kFB = 10;                 // frames per CTA (forward)
constexpr int kOH = 8;                  // output hops per CTA (inverse) -> kOH + 2 frames
constexpr int kThreads = 256;

__device__ float2 g_tw200[200];         // e^{-2 pi i t / 200}
__device__ float2 g_tw400[201];         // e^{-2 pi i k / 400}
__device__ float g_hann[400];           // 0.5 - 0.5 cos(2 pi n / 400)
__device__ float g_winv[400];           // hann[n] / sum_j hann^2[n mod 160 + 160 j]
__constant__ float2 c_tw25[25];         // e^{-2 pi i t / 25}

using namespace fft;

template <bool INV>
__device__ void fft200_block(const float2* __restrict__ in, float2* __restrict__ tmp, float2* __restrict__ out, int nf,
                             const float2* __restrict__ s_tw200) {
  for (int t = threadIdx.x; t < nf * 25; t += blockDim.x) fft200_step_a<INV>(in, tmp, t / 25, t % 25, s_tw200);
  __syncthreads();
  for (int t = threadIdx.x; t < nf * 8; t += blockDim.x) fft200_step_b<INV>(tmp, out, t >> 3, t & 7, c_tw25);
  __syncthreads();
}

// numpy: abs(int16(-32768)) == -32768, so that sample never wins the max (SN/apply.py:150)
__device__ __forceinline__ int abs16(int v) { return v == -32768 ? -32768 : (v < 0 ? -v : v); }

__global__ void peak_kernel(const int16_t* __restrict__ pcm, const long long* __restrict__ offs, int* __restrict__ peak) {
  const int u = blockIdx.x;
  const long long b = offs[u], e = offs[u + 1];
  int m = -32768;
  for (long long i = b + threadIdx.x; i < e; i += blockDim.x) m = max(m, abs16((int)pcm[i]));
  __shared__ int red[32];
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -32768;
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) peak[u] = (e > b) ? m : 0;
  }
}

// grid: (ceil(max_frames / kFB), U)
__global__ void __launch_bounds__(kThreads)
stft_kernel(const int16_t* __restrict__ pcm, const long long* __restrict__ offs, const long long* __restrict__ frame_offs,
            const int* __restrict__ peak, float* __restrict__ logmag, float* __restrict__ phase) {
  __shared__ float s_x[(kFB - 1) * kHop + kWin];
  __shared__ float2 s_a[kFB * 200];
  __shared__ float2 s_b[kFB * 200];
  __shared__ float2 s_tw200[200];
  __shared__ float2 s_tw400[201];
  const int u = blockIdx.y;
  const int T = (int)(frame_offs[u + 1] - frame_offs[u]);
  const int t0 = blockIdx.x * kFB;
  if (t0 >= T) return;
  const int nf = min(kFB, T - t0);
  const long long base = offs[u] + (long long)t0 * kHop;
  const int span = (nf - 1) * kHop + kWin;
  const double denom = (double)peak[u] + 0.000001;
  for (int i = threadIdx.x; i < span; i += blockDim.x) s_x[i] = (float)((double)pcm[base + i] / denom);
  for (int i = threadIdx.x; i < 200; i += blockDim.x) s_tw200[i] = g_tw200[i];
  for (int i = threadIdx.x; i < 201; i += blockDim.x) s_tw400[i] = g_tw400[i];
  __syncthreads();
  // window and pack even/odd samples into a complex-200 sequence
  for (int i = threadIdx.x; i < nf * 200; i += blockDim.x) {
    const int f = i / 200, m = i - f * 200;
    const float2 xx = *reinterpret_cast<const float2*>(&s_x[f * kHop + 2 * m]);
    s_a[i] = make_float2(xx.x * g_hann[2 * m], xx.y * g_hann[2 * m + 1]);
  }
  __syncthreads();
  fft200_block<false>(s_a, s_b, s_a, nf, s_tw200);
  // X[k] = (Z[k] + conj Z[200-k]) / 2 - (i/2) e^{-2 pi i k / 400} (Z[k] - conj Z[200-k]),  k = 0..200
  const long long row0 = frame_offs[u] + t0;
  for (int i = threadIdx.x; i < nf * kBinsD; i += blockDim.x) {
    const int f = i / kBinsD, k = i - f * kBinsD;
    const float2 X = rfft_post(s_a + f * 200, k, s_tw400);
    const float re = X.x, im = X.y;
    const float mag = sqrtf(re * re + im * im);
    const size_t o = (size_t)(row0 + f) * kBinsD + k;
    logmag[o] = logf(mag + 1e-5f);
    if (phase) phase[o] = atan2f(im, re);
  }
}

// Normalised (and, for the mixture, trimmed) float32 samples: stage a2 on its own, for the bit-exact test.
__global__ void normalise_kernel(const int16_t* __restrict__ pcm, const long long* __restrict__ offs,
                                 const long long* __restrict__ out_offs, const int* __restrict__ peak, float* __restrict__ out) {
  const int u = blockIdx.y;
  const long long n = out_offs[u + 1] - out_offs[u];
  const double denom = (double)peak[u] + 0.000001;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[out_offs[u] + i] = (float)((double)pcm[offs[u] + i] / denom);
}

// grid: (ceil((max_frames + 2) / kOH), U).  Output hop h (160 samples) sums frames h-2, h-1, h.
__global__ void __launch_bounds__(kThreads)
istft_kernel(const float* __restrict__ logmag, const float* __restrict__ phase, const long long* __restrict__ frame_offs,
             const long long* __restrict__ out_offs, const int* __restrict__ peak, float* __restrict__ out_f32,
             int16_t* __restrict__ out_i16) {
  constexpr int NF = kOH + 2;
  __shared__ float2 s_s[NF * kBinsD];     // spectrum, then FFT scratch, then the time-domain frames
  __shared__ float2 s_a[NF * 200];
  __shared__ float2 s_tw200[200];
  __shared__ float2 s_tw400[201];
  __shared__ float s_winv[kWin];
  const int u = blockIdx.y;
  const int T = (int)(frame_offs[u + 1] - frame_offs[u]);
  if (T <= 0) return;
  const int h0 = blockIdx.x * kOH;        // first output hop
  if (h0 >= T + 2) return;
  const int fa = h0 - 2;                  // first frame needed (may be negative)
  for (int i = threadIdx.x; i < 200; i += blockDim.x) s_tw200[i] = g_tw200[i];
  for (int i = threadIdx.x; i < 201; i += blockDim.x) s_tw400[i] = g_tw400[i];
  for (int i = threadIdx.x; i < kWin; i += blockDim.x) s_winv[i] = g_winv[i];
  const long long row0 = frame_offs[u];
  for (int i = threadIdx.x; i < NF * kBinsD; i += blockDim.x) {
    const int f = i / kBinsD, k = i - f * kBinsD;
    const int t = fa + f;
    float2 s = make_float2(0.f, 0.f);
    if (t >= 0 && t < T) {
      const size_t o = (size_t)(row0 + t) * kBinsD + k;
      const float a = expf(logmag[o]);
      float sn, cs;
      sincosf(phase[o], &sn, &cs);
      s = make_float2(a * cs, a * sn);
      if (k == 0 || k == 200) s.y = 0.f;  // irfft ignores the imaginary part of DC / Nyquist
    }
    s_s[i] = s;
  }
  __syncthreads();
  // Z[k] = (S[k] + conj S[200-k]) + i e^{+2 pi i k / 400} (S[k] - conj S[200-k]),  k = 0..199
  for (int i = threadIdx.x; i < NF * 200; i += blockDim.x) {
    const int f = i / 200, k = i - f * 200;
    s_a[i] = irfft_pre(s_s + f * kBinsD, k, s_tw400);
  }
  __syncthreads();
  fft200_block<true>(s_a, s_s, s_a, NF, s_tw200);
  // frames: y_f[2m] = Re z[m] / 400, y_f[2m+1] = Im z[m] / 400, times the synthesis window
  float* s_y = reinterpret_cast<float*>(s_s);          // [NF][400]
  for (int i = threadIdx.x; i < NF * 200; i += blockDim.x) {
    const int f = i / 200, m = i - f * 200;
    const float2 z = s_a[i];
    s_y[f * kWin + 2 * m] = z.x * (1.0f / 400.0f) * s_winv[2 * m];
    s_y[f * kWin + 2 * m + 1] = z.y * (1.0f / 400.0f) * s_winv[2 * m + 1];
  }
  __syncthreads();
  const long long n_out = out_offs[u + 1] - out_offs[u];   // (T - 1) * 160 + 400
  const float scale = (float)((double)peak[u] + 0.000001);
  for (int i = threadIdx.x; i < kOH * kHop; i += blockDim.x) {
    const int hl = i / kHop, r = i - hl * kHop;             // local hop, offset in hop
    const long long n = (long long)(h0 + hl) * kHop + r;
    if (n >= n_out) continue;
    // frame local index f = hl + 2 - j covers sample offset r + 160 j, j = 0..2 (zero frames outside [0, T))
    float acc = s_y[(hl + 2) * kWin + r] + s_y[(hl + 1) * kWin + r + kHop];
    if (r + 2 * kHop < kWin) acc += s_y[hl * kWin + r + 2 * kHop];
    const long long o = out_offs[u] + n;
    if (out_f32) out_f32[o] = acc;
    if (out_i16) {
      float v = acc * scale;
      v = fminf(fmaxf(v, -32768.f), 32767.f);
      out_i16[o] = (int16_t)__float2int_rn(v);
    }
  }
}

}  // namespace

cudaError_t dsp_init_tables() {
  const double kPi = 3.14159265358979323846;
  std::vector<float2> t200(200), t400(201), t25(25);
  std::vector<float> hann(400), winv(400);
  for (int i = 0; i < 200; ++i) t200[i] = make_float2((float)cos(2 * kPi * i / 200), (float)-sin(2 * kPi * i / 200));
  for (int i = 0; i < 201; ++i) t400[i] = make_float2((float)cos(2 * kPi * i / 400), (float)-sin(2 * kPi * i / 400));
  for (int i = 0; i < 25; ++i) t25[i] = make_float2((float)cos(2 * kPi * i / 25), (float)-sin(2 * kPi * i / 25));
  for (int i = 0; i < 400; ++i) hann[i] = (float)(0.5 - 0.5 * cos(2 * kPi * i / 400));
  for (int i = 0; i < 400; ++i) {
    double d = 0;
    for (int j = 0; j < 3; ++j) {
      int n = (i % 160) + 160 * j;
      if (n < 400) d += (double)hann[n] * (double)hann[n];
    }
    winv[i] = (float)((double)hann[i] / d);
  }
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(g_tw200, t200.data(), sizeof(float2) * 200)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(g_tw400, t400.data(), sizeof(float2) * 201)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_tw25, t25.data(), sizeof(float2) * 25)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(g_hann, hann.data(), sizeof(float) * 400)) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(g_winv, winv.data(), sizeof(float) * 400)) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t launch_peaks(cudaStream_t s, const int16_t* pcm, const long long* offs, int U, int* peak) {
  if (U <= 0) return cudaSuccess;
  peak_kernel<<<U, 1024, 0, s>>>(pcm, offs, peak);
  return cudaGetLastError();
}

cudaError_t launch_stft(cudaStream_t s, const int16_t* pcm, const long long* offs, const long long* frame_offs, int U,
                        const int* peak, int max_frames_per_clip, long long total_frames, float* logmag, float* phase) {
  (void)total_frames;
  if (U <= 0 || max_frames_per_clip <= 0) return cudaSuccess;
  dim3 grid((max_frames_per_clip + kFB - 1) / kFB, U);
  stft_kernel<<<grid, kThreads, 0, s>>>(pcm, offs, frame_offs, peak, logmag, phase);
  return cudaGetLastError();
}

cudaError_t launch_normalise(cudaStream_t s, const int16_t* pcm, const long long* offs, const long long* out_offs, int U,
                             const int* peak, float* out) {
  if (U <= 0) return cudaSuccess;
  dim3 grid(64, U);
  normalise_kernel<<<grid, 256, 0, s>>>(pcm, offs, out_offs, peak, out);
  return cudaGetLastError();
}

cudaError_t launch_istft(cudaStream_t s, const float* logmag, const float* phase, const long long* frame_offs,
                         const long long* out_offs, int U, const int* peak, long long total_blocks_hint,
                         int max_frames_per_clip, float* out_f32, int16_t* out_i16) {
  (void)total_blocks_hint;
  if (U <= 0 || max_frames_per_clip <= 0) return cudaSuccess;
  dim3 grid((max_frames_per_clip + 2 + kOH - 1) / kOH, U);
  istft_kernel<<<grid, kThreads, 0, s>>>(logmag, phase, frame_offs, out_offs, peak, out_f32, out_i16);
  return cudaGetLastError();
}

}  // namespace nhans
