// Shifted-row implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a).
//
//   D[m, n] = sum_j A_{map_j}[m + row_off_j, col_j : col_j + 64] . B[n, 64 j : 64 j + 64]      (fp16 x fp16 -> fp32)
//
// Every convolution and dense layer of N_HANS___Selective_Noise/main.py:98-242 except the two Cin = 1
// convolutions is lowered to this form by plan.cc.  One persistent CTA per SM, warp-specialised:
//
//   warp 0   TMA producer   one 128 x 64 A box (rows m0 + row_off_j) and one BN x 64 B box per k-block,
//                           128-byte swizzle, mbarrier complete_tx, multi-stage ring
//   warp 1   MMA issuer     tcgen05.mma.cta_group::1.kind::f16, M = 128, N = BN, 4 x K = 16 per k-block,
//                           accumulators in TMEM (2 x 256 columns, double buffered against the epilogue)
//   warps 2-5 epilogue      tcgen05.ld -> + per-utterance conditioning bias + time / frequency embedding
//                           tables + scaled identity residual / rank-1 transform -> ReLU -> fp16 store
//                           into the consumer's padded grid (or fp32 + centre frame for the head)
//
// The epilogue is the fusion of blocks.py:104-108 (batch-norm), main.py:166,172 (conditioning adds),
// main.py:184-186 (residual add, ReLU) folded as in SURVEY.md App. A.6.
#include "kernels.h"
#include "ptx.cuh"

namespace nhans {

namespace {

constexpr int kABytes = 128 * 128;        // 128 rows x 64 fp16
constexpr int kCtrlBytes = 4096;
constexpr int kMaxKb = 384;
constexpr int kSmemLimit = 227 * 1024;

struct __align__(8) Ctrl {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  KBlockDev kb[kMaxKb];
};
static_assert(sizeof(Ctrl) <= kCtrlBytes, "control block too large");

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_shift_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapB, const GemmDev p, const int stages) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int b_bytes = p.BN * 128;
  const int stage_bytes = kABytes + b_bytes;
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem + (size_t)stages * stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = p.N / p.BN;
  const int m_tiles = (p.M + 127) / 128;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.num_kb;

  for (int i = threadIdx.x; i < num_kb; i += blockDim.x) ctrl->kb[i] = p.kb[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&ctrl->full[s], 1);
      ptx::mbar_init(&ctrl->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&ctrl->tmem_full[a], 1);
      ptx::mbar_init(&ctrl->tmem_empty[a], 128);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&mapA0);
    ptx::tma_prefetch_desc(&mapA1);
    ptx::tma_prefetch_desc(&mapB);
  }
  if (warp == 1) ptx::tmem_alloc(&ctrl->tmem_base, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * 128;
        const int n0 = (tile % n_tiles) * p.BN;
        for (int j = 0; j < num_kb; ++j) {
          ptx::mbar_wait(&ctrl->empty[stage], phase ^ 1, p.err_flag, 1);
          const KBlockDev kb = ctrl->kb[j];
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          ptx::mbar_expect_tx(&ctrl->full[stage], (uint32_t)stage_bytes);
          ptx::tma_load_2d(sa, kb.map ? &mapA1 : &mapA0, &ctrl->full[stage], kb.col, m0 + kb.row_off);
          ptx::tma_load_2d(sa + kABytes, &mapB, &ctrl->full[stage], j * 64, n0);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_f16((uint32_t)p.BN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(&ctrl->tmem_empty[acc], acc_phase ^ 1, p.err_flag, 2);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int j = 0; j < num_kb; ++j) {
          ptx::mbar_wait(&ctrl->full[stage], phase, p.err_flag, 3);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t da = ptx::umma_desc_sw128(sa);
          const uint64_t db = ptx::umma_desc_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)         // 4 x (K = 16) = 64; +32 B inside the swizzle atom
            ptx::umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, (uint32_t)((j | k) != 0));
          ptx::umma_commit(&ctrl->empty[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(&ctrl->tmem_full[acc]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const EpiDev& e = p.epi;
    const int q = warp & 3;                   // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int hw = p.Hq * p.Wq;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * 128;
      const int n0 = (tile % n_tiles) * p.BN;
      const int m = m0 + row;
      bool valid = m < p.M;
      int unit = 0, ho = 0, wo = 0;
      if (valid) {
        unit = m / hw;
        const int rem = m - unit * hw;
        ho = rem / p.Wq;
        wo = rem - ho * p.Wq;
        valid = (ho < p.Ho) && (wo < p.Wo);
      }
      const float* bias_row = e.bias;
      const float* t_row = nullptr;
      const float* f_row = nullptr;
      const __half* res_row = nullptr;
      const float* raw_row = nullptr;
      float rawv = 0.f;
      __half* out_row = nullptr;
      float* outf_row = nullptr;
      if (valid) {
        const int utt = p.units.utt ? p.units.utt[unit] : 0;
        bias_row = e.bias + (size_t)utt * e.bias_stride;
        if (e.ttab) t_row = e.ttab + (size_t)ho * p.N;
        if (e.ftab) f_row = e.ftab + (size_t)wo * p.N;
        if (e.res) res_row = e.res + (size_t)m * e.res_C;
        if (e.r1_vec) {
          const int frame = p.units.frame[unit] + ho * e.r1_sh + e.raw_oh;
          if (frame >= p.units.lo[unit] && frame < p.units.hi[unit])
            rawv = e.raw[(size_t)frame * 201 + wo * e.r1_sw];
        }
        if (e.head) {
          raw_row = e.raw + (size_t)p.units.frame[unit] * 201;
          outf_row = e.out_f32 + (size_t)unit * 201;
        } else {
          long long pix;
          if (e.o_mode == 1) {
            pix = ((long long)unit * e.o_W + wo) * e.o_H + ho;
          } else {
            const int y = ho + e.o_oy, x = wo + e.o_ox;
            const int plane = (y % e.o_sh) * e.o_sw + (x % e.o_sw);
            pix = plane * e.o_plane + (long long)unit * e.o_Hq * e.o_Wq + (long long)(y / e.o_sh) * e.o_Wq + (x / e.o_sw);
          }
          out_row = e.out + pix * e.out_C;
        }
      }

      ptx::mbar_wait(&ctrl->tmem_full[acc], acc_phase, p.err_flag, 4);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(t_addr + c0, v);
        ptx::tmem_ld_wait();
        if (valid) {
          const int col = n0 + c0;
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b = *reinterpret_cast<const float4*>(bias_row + col + i);
            f[i] = __uint_as_float(v[i]) + b.x;
            f[i + 1] = __uint_as_float(v[i + 1]) + b.y;
            f[i + 2] = __uint_as_float(v[i + 2]) + b.z;
            f[i + 3] = __uint_as_float(v[i + 3]) + b.w;
          }
          if (t_row) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 t = *reinterpret_cast<const float4*>(t_row + col + i);
              f[i] += t.x; f[i + 1] += t.y; f[i + 2] += t.z; f[i + 3] += t.w;
            }
          }
          if (f_row) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 t = *reinterpret_cast<const float4*>(f_row + col + i);
              f[i] += t.x; f[i + 1] += t.y; f[i + 2] += t.z; f[i + 3] += t.w;
            }
          }
          if (res_row) {
            const uint4 r0 = *reinterpret_cast<const uint4*>(res_row + col);
            const uint4 r1 = *reinterpret_cast<const uint4*>(res_row + col + 8);
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 x = __half22float2(*reinterpret_cast<const __half2*>(&rr[i]));
              f[2 * i] = fmaf(e.res_scale[col + 2 * i], x.x, f[2 * i]);
              f[2 * i + 1] = fmaf(e.res_scale[col + 2 * i + 1], x.y, f[2 * i + 1]);
            }
          }
          if (e.r1_vec) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaf(e.r1_vec[col + i], rawv, f[i]);
          }
          if (e.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          if (e.head) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (col + i < 201) outf_row[col + i] = f[i] + raw_row[col + i];
          } else {
            uint4 o0, o1;
            o0.x = pack_half2(f[0], f[1]);   o0.y = pack_half2(f[2], f[3]);
            o0.z = pack_half2(f[4], f[5]);   o0.w = pack_half2(f[6], f[7]);
            o1.x = pack_half2(f[8], f[9]);   o1.y = pack_half2(f[10], f[11]);
            o1.z = pack_half2(f[12], f[13]); o1.w = pack_half2(f[14], f[15]);
            *reinterpret_cast<uint4*>(out_row + col) = o0;
            *reinterpret_cast<uint4*>(out_row + col + 8) = o1;
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ctrl->tmem_empty[acc]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int gemm_smem_bytes(int BN, int* stages_out) {
  const int stage_bytes = kABytes + BN * 128;
  int stages = (kSmemLimit - kCtrlBytes - 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages_out) *stages_out = stages;
  return stages * stage_bytes + kCtrlBytes + 1024;
}

cudaError_t gemm_configure() {
  return cudaFuncSetAttribute(gemm_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
}

cudaError_t launch_gemm(cudaStream_t s, int n_sm, const CUtensorMap& mapA0, const CUtensorMap& mapA1,
                        const CUtensorMap& mapB, const GemmDev& p) {
  if (p.M <= 0) return cudaSuccess;
  if (p.num_kb > kMaxKb || p.BN % 16 != 0 || p.BN > 256 || p.N % p.BN != 0) return cudaErrorInvalidValue;
  int stages = 0;
  const int smem = gemm_smem_bytes(p.BN, &stages);
  const int tiles = ((p.M + 127) / 128) * (p.N / p.BN);
  const int grid = tiles < n_sm ? tiles : n_sm;
  gemm_shift_kernel<<<grid, kGemmThreads, smem, s>>>(mapA0, mapA1, mapB, p, stages);
  return cudaGetLastError();
}

}  // namespace nhans
