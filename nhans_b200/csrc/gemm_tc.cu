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
//   warps 2-9 epilogue      two warps per TMEM lane quarter, alternating 32-column chunks:
//                           tcgen05.ld (thread = row) -> shared-memory transpose (warp-private, 32 rows x 32
//                           columns) -> 8 lanes per row x 4 channels each, so that every table load, the
//                           residual load and the fp16 store are coalesced; + per-utterance conditioning
//                           bias + time / frequency embedding tables + scaled identity residual / rank-1
//                           transform -> ReLU -> fp16 into the consumer's padded grid (or fp32 + centre
//                           frame for the head)
//
// The epilogue is the fusion of blocks.py:104-108 (batch-norm), main.py:166,172 (conditioning adds),
// main.py:184-186 (residual add, ReLU) folded as in SURVEY.md App. A.6.
#include "kernels.h"
#include "ptx.cuh"

namespace nhans {

namespace {

constexpr int kABytes = 128 * 128;        // 128 rows x 64 fp16
constexpr int kCtrlBytes = 4096;
constexpr int kStagePitch = 36;           // floats per staged row (32 + 4: conflict-free 16-byte accesses)
constexpr int kEpiWarps = 8;
constexpr int kEpiWarpBytes = 32 * kStagePitch * 4 + 32 * 5 * 4;   // staging + per-row metadata of one warp
constexpr int kEpiBytes = kEpiWarps * kEpiWarpBytes;
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
      ptx::mbar_init(&ctrl->tmem_empty[a], kEpiWarps * 32);
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
    // ===================== epilogue (warps 2..9) =====================
    const EpiDev& e = p.epi;
    const int q = warp & 3;                   // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;         // which of the two warps of the quarter: even / odd 32-column chunks
    const int hw = p.Hq * p.Wq;
    uint8_t* epi_base = reinterpret_cast<uint8_t*>(ctrl) + kCtrlBytes + (warp - 2) * kEpiWarpBytes;
    float* stage = reinterpret_cast<float*>(epi_base);
    int* m_pix = reinterpret_cast<int*>(epi_base + 32 * kStagePitch * 4);        // [32] output pixel, -1 = skip
    int* m_ho = m_pix + 32;
    int* m_wo = m_ho + 32;
    int* m_utt = m_wo + 32;
    float* m_raw = reinterpret_cast<float*>(m_utt + 32);
    const int sub = lane >> 3;                // row within a group of 4
    const int c4 = (lane & 7) * 4;            // 4 channels per lane
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * 128;
      const int n0 = (tile % n_tiles) * p.BN;
      if (p.debug_skip_epilogue) {
        ptx::mbar_wait(&ctrl->tmem_full[acc], acc_phase, p.err_flag, 4);
        ptx::tc_fence_before();
        ptx::mbar_arrive(&ctrl->tmem_empty[acc]);
        continue;
      }
      {
        // per-row metadata, computed by the thread that owns the row in TMEM
        const int m = m0 + q * 32 + lane;
        bool valid = m < p.M;
        int unit = 0, ho = 0, wo = 0;
        if (valid) {
          unit = m / hw;
          const int rem = m - unit * hw;
          ho = rem / p.Wq;
          wo = rem - ho * p.Wq;
          valid = (ho < p.Ho) && (wo < p.Wo);
        }
        int pix = -1, utt = 0;
        float rawv = 0.f;
        if (valid) {
          utt = p.units.utt ? p.units.utt[unit] : 0;
          if (e.r1_vec) {
            const int frame = p.units.frame[unit] + ho * e.r1_sh + e.raw_oh;
            if (frame >= p.units.lo[unit] && frame < p.units.hi[unit]) rawv = e.raw[(size_t)frame * 201 + wo * e.r1_sw];
          }
          if (e.head) {
            pix = unit;
            utt = p.units.frame[unit];       // head: the centre frame row replaces the utterance index
          } else if (e.o_mode == 1) {
            pix = (unit * e.o_W + wo) * e.o_H + ho;
          } else {
            const int y = ho + e.o_oy, x = wo + e.o_ox;
            const int plane = (y % e.o_sh) * e.o_sw + (x % e.o_sw);
            pix = (int)(plane * e.o_plane + (long long)unit * e.o_Hq * e.o_Wq + (long long)(y / e.o_sh) * e.o_Wq + (x / e.o_sw));
          }
        }
        m_pix[lane] = pix; m_ho[lane] = ho; m_wo[lane] = wo; m_utt[lane] = utt; m_raw[lane] = rawv;
      }
      __syncwarp();
      int pixs[8];
      float4 fb[8], ft[8], ff[8];
      uint2 fx[8];
      // Issues every global load of one 32-column chunk (read-only path); called before the accumulator
      // is awaited so that the memory latency hides behind the MMA main loop.
      auto issue_loads = [&](int c0) {
        const int col = n0 + c0 + c4;
        const bool lane_ok = c4 < p.BN - c0 && !e.head;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int r = g * 4 + sub;
          pixs[g] = m_pix[r];
          const bool ok = lane_ok && pixs[g] >= 0;
          fb[g] = ok ? __ldg(reinterpret_cast<const float4*>(e.bias + (size_t)m_utt[r] * e.bias_stride + col)) : zero4;
          ft[g] = (ok && e.ttab) ? __ldg(reinterpret_cast<const float4*>(e.ttab + (size_t)m_ho[r] * p.N + col)) : zero4;
          ff[g] = (ok && e.ftab) ? __ldg(reinterpret_cast<const float4*>(e.ftab + (size_t)m_wo[r] * p.N + col)) : zero4;
          fx[g] = (ok && e.res) ? __ldg(reinterpret_cast<const uint2*>(e.res + (size_t)(m0 + q * 32 + r) * e.res_C + col)) : make_uint2(0u, 0u);
        }
      };
      int c0 = half * 32;
      if (c0 < p.BN) issue_loads(c0);
      ptx::mbar_wait(&ctrl->tmem_full[acc], acc_phase, p.err_flag, 4);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      for (; c0 < p.BN; c0 += 64) {
        const int width = min(32, p.BN - c0);
        // phase 1: TMEM (thread = row) -> staging
        if (width == 32) {
          uint32_t v[32];
          ptx::tmem_ld32(t_addr + c0, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint4*>(stage + lane * kStagePitch + 4 * i) = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
          uint32_t v[16];
          ptx::tmem_ld16(t_addr + c0, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(stage + lane * kStagePitch + 4 * i) = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        __syncwarp();
        // phase 2: 8 lanes per row, 4 channels per lane
        if (c4 < width) {
          const int col = n0 + c0 + c4;
          if (e.head) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col));
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const int r = g * 4 + sub;
              const int pix = m_pix[r];
              if (pix < 0) continue;
              const float4 f = *reinterpret_cast<const float4*>(stage + r * kStagePitch + c4);
              const float* raw_row = e.raw + (size_t)m_utt[r] * 201;
              float* o = e.out_f32 + (size_t)pix * 201;
              if (col < 201) o[col] = f.x + b.x + raw_row[col];
              if (col + 1 < 201) o[col + 1] = f.y + b.y + raw_row[col + 1];
              if (col + 2 < 201) o[col + 2] = f.z + b.z + raw_row[col + 2];
              if (col + 3 < 201) o[col + 3] = f.w + b.w + raw_row[col + 3];
            }
          } else {
            const float4 rs = e.res ? __ldg(reinterpret_cast<const float4*>(e.res_scale + col)) : zero4;
            const float4 r1 = e.r1_vec ? __ldg(reinterpret_cast<const float4*>(e.r1_vec + col)) : zero4;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (pixs[g] < 0) continue;
              const int r = g * 4 + sub;
              float4 f = *reinterpret_cast<const float4*>(stage + r * kStagePitch + c4);
              f.x += fb[g].x + ft[g].x + ff[g].x; f.y += fb[g].y + ft[g].y + ff[g].y;
              f.z += fb[g].z + ft[g].z + ff[g].z; f.w += fb[g].w + ft[g].w + ff[g].w;
              const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&fx[g].x));
              const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&fx[g].y));
              f.x = fmaf(rs.x, x0.x, f.x); f.y = fmaf(rs.y, x0.y, f.y);
              f.z = fmaf(rs.z, x1.x, f.z); f.w = fmaf(rs.w, x1.y, f.w);
              if (e.r1_vec) {
                const float rawv = m_raw[r];
                f.x = fmaf(r1.x, rawv, f.x); f.y = fmaf(r1.y, rawv, f.y);
                f.z = fmaf(r1.z, rawv, f.z); f.w = fmaf(r1.w, rawv, f.w);
              }
              if (e.relu) {
                f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f);
              }
              uint2 o;
              o.x = pack_half2(f.x, f.y);
              o.y = pack_half2(f.z, f.w);
              *reinterpret_cast<uint2*>(e.out + (size_t)pixs[g] * e.out_C + col) = o;
            }
          }
        }
        __syncwarp();                         // staging is overwritten by the next chunk
        if (c0 + 64 < p.BN) issue_loads(c0 + 64);
      }
      __syncwarp();                           // metadata is overwritten by the next tile
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
  int stages = (kSmemLimit - kCtrlBytes - kEpiBytes - 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages_out) *stages_out = stages;
  return stages * stage_bytes + kCtrlBytes + kEpiBytes + 1024;
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
