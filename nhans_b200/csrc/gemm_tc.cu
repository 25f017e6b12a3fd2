// Shifted-row implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a).
//
//   D[m, n] = sum_g sum_{t < ntaps_g}  A_{map_g}[m + row_off_g + shift_{g,t}, col_g : col_g + 64] . B[n, 64 bk_{g,t} : +64]
//
// Every convolution and dense layer of N_HANS___Selective_Noise/main.py:98-242 except the two Cin = 1
// convolutions is lowered to this form by plan.cc (a group g = the kw taps of one kernel row and one
// 64-channel chunk).  One persistent CTA per SM; a CTA tile is (256 / BN) sub-tiles of 128 rows x BN
// columns, i.e. always 256 fp32 accumulator columns in TMEM, double buffered (2 x 256 = all 512 columns).
//
//   warp 8   A producer     one TMA box of 136 rows x 64 channels ("slab") per (group, sub-tile): the kw taps
//                           of a kernel row read the SAME slab through UMMA descriptors whose start address
//                           is shifted by `shift` rows, so A comes from L2 once per kernel row, not per tap
//   warp 9   B producer     one TMA box BN x 64 per k-block into its own ring; when all k-blocks of the layer
//                           fit (K * BN * 2 B <= ring) the weights are loaded once and stay resident
//   warp 10  MMA issuer     tcgen05.mma.cta_group::1.kind::f16, M = 128, N = BN, 4 x (K = 16) per tap,
//                           fp32 accumulators in TMEM
//   warps 0-7 epilogue      tcgen05.ld (thread = row) -> shared-memory transpose (warp-private 32 x 32) ->
//                           8 lanes per row x 4 channels, so every table load, the residual load and the
//                           fp16 store are coalesced; all global loads of a chunk are issued before the
//                           accumulator is awaited.  + per-utterance conditioning bias + time / frequency
//                           embedding tables + scaled identity residual / rank-1 transform -> ReLU -> fp16
//                           into the consumer's padded grid (or fp32 + centre frame for the head)
//
// The epilogue is the fusion of blocks.py:104-108 (batch-norm), main.py:166,172 (conditioning adds),
// main.py:184-186 (residual add, ReLU) folded as in SURVEY.md App. A.6.
#include "kernels.h"
#include "ptx.cuh"

namespace nhans {

namespace {

constexpr int kSlabRows = 136;                    // 128 + up to 7 rows of tap shift, multiple of 8
constexpr int kSlabBytes = kSlabRows * 128;       // 17 KB, a multiple of the 1024-byte swizzle atom
constexpr int kCtrlBytes = 8192;
constexpr int kEpiWarps = 8;
constexpr int kEpiWarpBytes = 32 * 64 + 32 * 16;                   // 32 x 16 fp32 transpose buffer + row metadata
constexpr int kEpiBytes = kEpiWarps * kEpiWarpBytes;
constexpr int kMaxGroups = 212;
constexpr int kMaxA = 8, kMaxB = 16;
constexpr int kSmemLimit = 227 * 1024;

struct __align__(8) Ctrl {
  uint64_t a_full[kMaxA], a_empty[kMaxA];
  uint64_t b_full[kMaxB], b_empty[kMaxB];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  KGroupDev groups[kMaxGroups];
};
static_assert(sizeof(Ctrl) <= kCtrlBytes, "control block too large");

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// MMA issuer (one warp): walks tiles / groups / sub-tiles in the producers' order; one elected lane issues
// tcgen05.mma.  IL sub-tiles are processed together so that consecutive MMAs target different TMEM
// accumulators (back-to-back MMAs into the same accumulator serialise on the accumulate dependency).
template <int IL>
__device__ __forceinline__ void mma_issuer(Ctrl* ctrl, const GemmDev& p, const GemmCfg& cfg, uint8_t* smem_a, uint8_t* smem_b,
                                           uint32_t tmem_base, int num_tiles, int b_bytes) {
  const int MT = cfg.mt;
  const int num_groups = p.num_groups;
  const uint32_t idesc = ptx::umma_idesc_f16((uint32_t)p.BN);
  const uint64_t desc_hi = ptx::umma_desc_sw128(0, 0);             // everything but the address field
  const uint32_t a_base = ptx::smem_u32(smem_a) >> 4, b_base = ptx::smem_u32(smem_b) >> 4;
  const uint32_t b_step = (uint32_t)b_bytes >> 4;
  uint32_t aslot = 0, aphase = 0, bslot0 = 0, bphase0 = 0, it = 0;
  bool b_ready = false;                                             // resident weights have landed
  long long w_tmem = 0, w_a = 0, w_b = 0;
  const long long t_start = clock64();
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
    ptx::mbar_wait_timed(&ctrl->tmem_empty[acc], acc_phase ^ 1, p.err_flag, 2, &w_tmem);
    ptx::tc_fence_after();
    for (int g = 0; g < num_groups; ++g) {
      const int ntaps = ctrl->groups[g].ntaps;
      for (int i0 = 0; i0 < MT; i0 += IL) {
        uint32_t a_lo[IL], d_tm[IL];
#pragma unroll
        for (int ii = 0; ii < IL; ++ii) {
          uint32_t sl = aslot + ii, ph = aphase;
          if (sl >= (uint32_t)cfg.na) { sl -= cfg.na; ph ^= 1; }
          ptx::mbar_wait_timed(&ctrl->a_full[sl], ph, p.err_flag, 3, &w_a);
          a_lo[ii] = a_base + sl * (kSlabBytes >> 4);
          d_tm[ii] = tmem_base + acc * 256 + (i0 + ii) * p.BN;
        }
        uint32_t bslot = bslot0, bphase = bphase0;
        for (int t = 0; t < ntaps; ++t) {
          if (cfg.resident) bslot = (uint32_t)ctrl->groups[g].bk[t];
          if ((i0 == 0 && !cfg.resident) || (cfg.resident && !b_ready))
            ptx::mbar_wait_timed(&ctrl->b_full[bslot], cfg.resident ? 0u : bphase, p.err_flag, 6, &w_b);
          ptx::tc_fence_after();
          const uint32_t sh = (uint32_t)ctrl->groups[g].shift[t] * 8;
          const uint32_t b_lo = b_base + bslot * b_step;
          const uint32_t first = (uint32_t)((g | t) != 0);
          if (ptx::elect_one()) {
            const uint64_t db = desc_hi | b_lo;
#pragma unroll
            for (int k = 0; k < 4; ++k) {                         // +32 B inside the swizzle atom per K = 16
#pragma unroll
              for (int ii = 0; ii < IL; ++ii)
                ptx::umma_f16(d_tm[ii], (desc_hi | (a_lo[ii] + sh)) + 2 * k, db + 2 * k, idesc, k == 0 ? first : 1u);
            }
            if (i0 + IL >= MT && !cfg.resident) ptx::umma_commit(&ctrl->b_empty[bslot]);
          }
          __syncwarp();
          if (!cfg.resident && ++bslot == (uint32_t)cfg.nb) { bslot = 0; bphase ^= 1; }
        }
#pragma unroll
        for (int ii = 0; ii < IL; ++ii) {
          if (ptx::elect_one()) ptx::umma_commit(&ctrl->a_empty[aslot]);
          __syncwarp();
          if (++aslot == (uint32_t)cfg.na) { aslot = 0; aphase ^= 1; }
        }
        if (i0 + IL >= MT) { bslot0 = bslot; bphase0 = bphase; }
      }
    }
    b_ready = true;
    if (ptx::elect_one()) ptx::umma_commit(&ctrl->tmem_full[acc]);
    __syncwarp();
  }
  if (p.debug_stats && (threadIdx.x & 31) == 0) {
    atomicAdd(p.debug_stats + 0, (unsigned long long)w_tmem);
    atomicAdd(p.debug_stats + 1, (unsigned long long)w_a);
    atomicAdd(p.debug_stats + 2, (unsigned long long)w_b);
    atomicAdd(p.debug_stats + 4, (unsigned long long)(clock64() - t_start));
  }
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_shift_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapB, const GemmDev p, const GemmCfg cfg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int b_bytes = p.BN * 128;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + cfg.na * kSlabBytes;
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(smem_b + (size_t)cfg.nb * b_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int MT = cfg.mt;                          // sub-tiles of 128 rows per CTA tile
  const int n_tiles = p.N / p.BN;
  const int m_tiles = (p.M + MT * 128 - 1) / (MT * 128);
  const int num_tiles = m_tiles * n_tiles;
  const int num_groups = p.num_groups;

  for (int i = threadIdx.x; i < num_groups; i += blockDim.x) ctrl->groups[i] = p.groups[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < cfg.na; ++s) { ptx::mbar_init(&ctrl->a_full[s], 1); ptx::mbar_init(&ctrl->a_empty[s], 1); }
    for (int s = 0; s < cfg.nb; ++s) { ptx::mbar_init(&ctrl->b_full[s], 1); ptx::mbar_init(&ctrl->b_empty[s], 1); }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&ctrl->tmem_full[a], 1);
      ptx::mbar_init(&ctrl->tmem_empty[a], kEpiWarps * 32);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async();
  }
  if (warp == 8 && lane == 0) {
    ptx::tma_prefetch_desc(&mapA0);
    ptx::tma_prefetch_desc(&mapA1);
    ptx::tma_prefetch_desc(&mapB);
  }
  if (warp == 10) ptx::tmem_alloc(&ctrl->tmem_base, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;

  // Warp roles: the hardware arbiter favours the highest warp id of a sub-partition, so the single-lane
  // issuers (warps 8-10) out-prioritise the ALU-heavy epilogue warps (0-7) they share a scheduler with.
  if (warp == 8) {
    // ===================== A producer: one slab per (group, sub-tile) =====================
    // (whole warp runs the loop; one elected lane issues the TMA - keeps the control flow warp-uniform)
    uint32_t slot = 0, phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * (MT * 128);
      for (int g = 0; g < num_groups; ++g) {
        const int row_off = ctrl->groups[g].row_off, col = ctrl->groups[g].col, map = ctrl->groups[g].map;
        for (int i = 0; i < MT; ++i) {
          ptx::mbar_wait(&ctrl->a_empty[slot], phase ^ 1, p.err_flag, 1);
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&ctrl->a_full[slot], kSlabBytes);
            ptx::tma_load_2d(smem_a + (size_t)slot * kSlabBytes, map ? &mapA1 : &mapA0, &ctrl->a_full[slot], col,
                             m0 + i * 128 + row_off);
          }
          __syncwarp();
          if (++slot == (uint32_t)cfg.na) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== B producer =====================
    if (cfg.resident) {
      // the whole packed weight matrix fits: load every k-block once, never release
      for (int t = 0; t < p.num_kb; ++t) {
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&ctrl->b_full[t], (uint32_t)b_bytes);
          ptx::tma_load_2d(smem_b + (size_t)t * b_bytes, &mapB, &ctrl->b_full[t], t * 64, 0);
        }
        __syncwarp();
      }
    } else {
      uint32_t slot = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = (tile % n_tiles) * p.BN;
        for (int g = 0; g < num_groups; ++g) {
          const int ntaps = ctrl->groups[g].ntaps;
          for (int t = 0; t < ntaps; ++t) {
            const int bk = ctrl->groups[g].bk[t];
            ptx::mbar_wait(&ctrl->b_empty[slot], phase ^ 1, p.err_flag, 5);
            if (ptx::elect_one()) {
              ptx::mbar_expect_tx(&ctrl->b_full[slot], (uint32_t)b_bytes);
              ptx::tma_load_2d(smem_b + (size_t)slot * b_bytes, &mapB, &ctrl->b_full[slot], bk * 64, n0);
            }
            __syncwarp();
            if (++slot == (uint32_t)cfg.nb) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 10) {
    // ===================== MMA issuer =====================
    if (cfg.il == 4) mma_issuer<4>(ctrl, p, cfg, smem_a, smem_b, tmem_base, num_tiles, b_bytes);
    else if (cfg.il == 2) mma_issuer<2>(ctrl, p, cfg, smem_a, smem_b, tmem_base, num_tiles, b_bytes);
    else mma_issuer<1>(ctrl, p, cfg, smem_a, smem_b, tmem_base, num_tiles, b_bytes);
  } else {
    // ===================== epilogue (warps 0..7) =====================
    // With ~225 KB of shared memory in use there is no L1 left: every table / residual read is an L2 round
    // trip of a few thousand cycles under load, so the epilogue is latency bound.  Hence: two warps per TMEM
    // lane quarter (alternating 16-column chunks), small chunks whose loads fit in registers twice, and the
    // loads of chunk k+1 in flight while chunk k is transposed and stored.
    const EpiDev& e = p.epi;
    const int ew = warp;
    const int q = warp & 3;                   // TMEM lane quarter this warp may read
    const int half = ew >> 2;                 // which of the two warps of the quarter
    const int hw = p.Hq * p.Wq;
    uint8_t* epi_base = reinterpret_cast<uint8_t*>(ctrl) + kCtrlBytes + ew * kEpiWarpBytes;
    uint4* stage = reinterpret_cast<uint4*>(epi_base);                 // [32 rows][4 x 16 B], XOR swizzled
    int4* meta = reinterpret_cast<int4*>(epi_base + 32 * 64);          // [32] {pixel, tf row, utt, raw bits}
    const int sub = lane >> 2;                // row within a group of 8
    const int jc = lane & 3;                  // 16-byte column slot: channels 4 jc .. 4 jc + 3 of the chunk
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    struct LoadSet {
      float4 b[4], t[4];
      uint2 x[4];
    };
    uint32_t it = 0;
    long long w_full = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int n0 = (tile % n_tiles) * p.BN;
      if (p.debug_skip_epilogue == 1) {
        ptx::mbar_wait(&ctrl->tmem_full[acc], acc_phase, p.err_flag, 4);
        ptx::tc_fence_before();
        ptx::mbar_arrive(&ctrl->tmem_empty[acc]);
        continue;
      }
      const int tile_m0 = (tile / n_tiles) * (MT * 128);
      // ---- per-row metadata of every sub-tile, computed by the thread that owns the row (loads overlap) ----
      int4 md[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        md[i] = make_int4(-1, 0, 0, 0);
        if (i >= MT) continue;
        const int m = tile_m0 + i * 128 + q * 32 + lane;
        if (m >= p.M) continue;
        const int unit = m / hw;
        const int rem = m - unit * hw;
        const int ho = rem / p.Wq;
        const int wo = rem - ho * p.Wq;
        if (ho >= p.Ho || wo >= p.Wo) continue;
        int pix, utt = p.units.utt[unit];
        float rawv = 0.f;
        if (e.r1_vec) {
          const int frame = p.units.frame[unit] + ho * e.r1_sh + e.raw_oh;
          if (frame >= p.units.lo[unit] && frame < p.units.hi[unit]) rawv = __ldg(e.raw + (size_t)frame * 201 + wo * e.r1_sw);
        }
        if (e.head) {
          pix = unit;
          utt = p.units.frame[unit];         // head: the centre frame row replaces the utterance index
        } else if (e.o_mode == 1) {
          pix = (unit * e.o_W + wo) * e.o_H + ho;
        } else {
          const int y = ho + e.o_oy, x = wo + e.o_ox;
          const int plane = (y % e.o_sh) * e.o_sw + (x % e.o_sw);
          pix = (int)(plane * e.o_plane + (long long)unit * e.o_Hq * e.o_Wq + (long long)(y / e.o_sh) * e.o_Wq + (x / e.o_sw));
        }
        md[i] = make_int4(pix, ho * p.Wo + wo, utt, __float_as_int(rawv));
      }
      bool waited = false;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i >= MT) break;
        const int m0 = tile_m0 + i * 128;
        if (m0 >= p.M) break;
        __syncwarp();
        meta[lane] = md[i];
        __syncwarp();
        int pixs[4], tfr[4], utts[4];
        float raws[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int4 mm = meta[g * 8 + sub];
          pixs[g] = mm.x; tfr[g] = mm.y; utts[g] = mm.z; raws[g] = __int_as_float(mm.w);
        }
        // every global load of one 16-column chunk (read-only path)
        auto issue_loads = [&](LoadSet& L, int c0) {
          const int col = n0 + c0 + 4 * jc;
          const bool lane_ok = c0 < p.BN && !e.head && p.debug_skip_epilogue != 2;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const bool ok = lane_ok && pixs[g] >= 0;
            L.b[g] = ok ? __ldg(reinterpret_cast<const float4*>(e.bias + (size_t)utts[g] * e.bias_stride + col)) : zero4;
            L.t[g] = (ok && e.tftab) ? __ldg(reinterpret_cast<const float4*>(e.tftab + (size_t)tfr[g] * p.N + col)) : zero4;
            L.x[g] = (ok && e.res) ? __ldg(reinterpret_cast<const uint2*>(e.res + (size_t)(m0 + q * 32 + g * 8 + sub) * e.res_C + col)) : make_uint2(0u, 0u);
          }
        };
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + i * p.BN;
        // one 16-column chunk: TMEM -> swizzled staging -> 4 lanes per row
        auto process = [&](const LoadSet& L, int c0) {
          {
            uint32_t v[16];
            if (p.debug_skip_epilogue != 5) {
              ptx::tmem_ld16(t_addr + c0, v);
              ptx::tmem_ld_wait();
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = lane + j;
            }
            if (p.debug_skip_epilogue == 4) {
              uint32_t x = 0;
#pragma unroll
              for (int j = 0; j < 16; ++j) x ^= v[j];
              if (x == 0x12345678u) stage[lane] = make_uint4(x, x, x, x);   // keeps the load alive
              return;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              stage[lane * 4 + ((j ^ (lane >> 1)) & 3)] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          __syncwarp();
          const int col = n0 + c0 + 4 * jc;
          if (e.head) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + col));
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (pixs[g] < 0) continue;
              const int r = g * 8 + sub;
              const uint4 u = stage[r * 4 + ((jc ^ (r >> 1)) & 3)];
              const float* raw_row = e.raw + (size_t)utts[g] * 201;
              float* o = e.out_f32 + (size_t)pixs[g] * 201;
              if (col < 201) o[col] = __uint_as_float(u.x) + b.x + raw_row[col];
              if (col + 1 < 201) o[col + 1] = __uint_as_float(u.y) + b.y + raw_row[col + 1];
              if (col + 2 < 201) o[col + 2] = __uint_as_float(u.z) + b.z + raw_row[col + 2];
              if (col + 3 < 201) o[col + 3] = __uint_as_float(u.w) + b.w + raw_row[col + 3];
            }
          } else {
            const float4 rs = e.res ? __ldg(reinterpret_cast<const float4*>(e.res_scale + col)) : zero4;
            const float4 r1 = e.r1_vec ? __ldg(reinterpret_cast<const float4*>(e.r1_vec + col)) : zero4;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (pixs[g] < 0) continue;
              const int r = g * 8 + sub;
              const uint4 u = stage[r * 4 + ((jc ^ (r >> 1)) & 3)];
              float4 f = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
              f.x += L.b[g].x + L.t[g].x; f.y += L.b[g].y + L.t[g].y;
              f.z += L.b[g].z + L.t[g].z; f.w += L.b[g].w + L.t[g].w;
              const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&L.x[g].x));
              const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&L.x[g].y));
              f.x = fmaf(rs.x, x0.x, f.x); f.y = fmaf(rs.y, x0.y, f.y);
              f.z = fmaf(rs.z, x1.x, f.z); f.w = fmaf(rs.w, x1.y, f.w);
              f.x = fmaf(r1.x, raws[g], f.x); f.y = fmaf(r1.y, raws[g], f.y);
              f.z = fmaf(r1.z, raws[g], f.z); f.w = fmaf(r1.w, raws[g], f.w);
              if (e.relu) {
                f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f);
              }
              uint2 o;
              o.x = pack_half2(f.x, f.y);
              o.y = pack_half2(f.z, f.w);
              if (p.debug_skip_epilogue != 3) *reinterpret_cast<uint2*>(e.out + (size_t)pixs[g] * e.out_C + col) = o;
            }
          }
          __syncwarp();                       // staging is overwritten by the next chunk
        };
        // chunks of this warp: c = (half + i) & 1, +2, ... (the two warps of a quarter alternate)
        LoadSet A, B;
        int c0 = ((half + i) & 1) * 16;
        issue_loads(A, c0);
        if (!waited) {
          ptx::mbar_wait_timed(&ctrl->tmem_full[acc], acc_phase, p.err_flag, 4, &w_full);
          ptx::tc_fence_after();
          waited = true;
        }
        for (; c0 < p.BN; c0 += 64) {
          issue_loads(B, c0 + 32);            // in flight while chunk c0 is processed (no-op past the end)
          process(A, c0);
          if (c0 + 32 < p.BN) {
            issue_loads(A, c0 + 64);
            process(B, c0 + 32);
          }
        }
      }
      if (!waited) ptx::mbar_wait(&ctrl->tmem_full[acc], acc_phase, p.err_flag, 4);
      ptx::tc_fence_before();
      ptx::mbar_arrive(&ctrl->tmem_empty[acc]);
    }
    if (p.debug_stats && threadIdx.x == 0) atomicAdd(p.debug_stats + 3, (unsigned long long)w_full);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    __syncwarp();
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int gemm_smem_bytes(int BN, int num_kb, GemmCfg* cfg) {
  const int b_bytes = BN * 128;
  GemmCfg c;
  c.mt = 256 / BN < 1 ? 1 : 256 / BN;
  c.na = 4;
  const int budget = kSmemLimit - 1024 - kCtrlBytes - kEpiBytes - c.na * kSlabBytes;
  c.nb = budget / b_bytes;
  if (c.nb > kMaxB) c.nb = kMaxB;
  c.resident = (num_kb <= c.nb) ? 1 : 0;
  c.desc_mode = 0;
  c.il = c.mt >= 2 ? 2 : 1;
  if (cfg) *cfg = c;
  return 1024 + c.na * kSlabBytes + c.nb * b_bytes + kCtrlBytes + kEpiBytes;
}

cudaError_t gemm_configure() {
  return cudaFuncSetAttribute(gemm_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
}

cudaError_t launch_gemm(cudaStream_t s, int n_sm, const CUtensorMap& mapA0, const CUtensorMap& mapA1,
                        const CUtensorMap& mapB, const GemmDev& p, int desc_mode) {
  if (p.M <= 0) return cudaSuccess;
  if (p.num_groups > kMaxGroups || p.BN % 16 != 0 || p.BN > 256 || p.N % p.BN != 0) return cudaErrorInvalidValue;
  GemmCfg cfg;
  const int smem = gemm_smem_bytes(p.BN, p.num_kb, &cfg);
  if (p.N != p.BN) cfg.resident = 0;
  cfg.desc_mode = desc_mode & 1;
  if (desc_mode >> 1) { cfg.il = desc_mode >> 1; if (cfg.il > cfg.mt) cfg.il = cfg.mt; }   // debug override
  if (cfg.nb < 4) return cudaErrorInvalidValue;   // a group has up to 4 taps in flight
  const int tiles = ((p.M + cfg.mt * 128 - 1) / (cfg.mt * 128)) * (p.N / p.BN);
  const int grid = tiles < n_sm ? tiles : n_sm;
  gemm_shift_kernel<<<grid, kGemmThreads, smem, s>>>(mapA0, mapA1, mapB, p, cfg);
  return cudaGetLastError();
}

}  // namespace nhans
