// GEMM (N = 512) with the residual add and post-LayerNorm(s) fused into the epilogue (sm_100a, CTA pair).
//
//   CHAIN = false:  h <- LN(h + A.W^T + b; g1, b1)                              (linear2 + norm3)
//   CHAIN = true :  h <- LN( LN(h + A.W^T + b; g1, b1) + c[row % Beff]; g2, b2 ) (self-attn out_proj + norm1 +
//                        1-token cross-attention constant + norm2, model/cmdm.py:224-227 via nn.TransformerDecoderLayer)
//
// A CTA pair owns a 256 x 512 tile: each CTA holds 128 complete rows of the result in all 512 TMEM columns, so the
// LayerNorm statistics are local to a thread (one thread = one row, tcgen05.ld 32x32b) plus one exchange between the
// two warps that share a row (one per column half).  Compared with GEMM -> tmp -> LayerNorm kernel this removes the
// fp32 tmp round trip (2 x 31 MB per LayerNorm at M = 15 360) and the LayerNorm launches, and the 512-wide tile cuts
// the operand traffic per flop by 25 % (A is loaded once for both N halves).
//
// Main loop: TMA -> 2-stage smem ring (A hi/lo 128x64, W hi/lo 2 x 128x64 per CTA) -> tcgen05.mma.cta_group::2,
// 2 (N halves) x 3 (bf16x3) MMAs of 256x256x16 per k-step.  Epilogue (8 warps per CTA, after the main loop, the
// operand ring is reused as staging): the residual and c tiles arrive by TMA (4-deep ring per warp), v / y are kept in
// TMEM between the passes (tcgen05.st), outputs (fp32 h + bf16 hi/lo) leave by TMA stores.
#pragma once
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "ptx.cuh"

namespace regen {
namespace gemmln {

constexpr int BM = 128, BK = 64, UMMA_K = 16, ND = 512;
constexpr int kThreads = 320;
constexpr int STAGE_BYTES = 6 * 16384;  // A_hi | A_lo | W_hi[0] | W_hi[1] | W_lo[0] | W_lo[1]
constexpr int STAGES = 2;
constexpr int EPI_WARP_BYTES = 24576;   // res ring 4 x 2 KB | c ring 4 x 2 KB | 2 x (fp32 2 KB + hi 1 KB + lo 1 KB)
constexpr int PARAM_BYTES = 5 * ND * 4; // bias, g1, b1, g2, b2
constexpr int STATS_BYTES = 2 * 128 * 2 * 8;
constexpr int BAR_BYTES = 1024;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + PARAM_BYTES + STATS_BYTES + BAR_BYTES + 1024;
static_assert(8 * EPI_WARP_BYTES <= STAGES * STAGE_BYTES, "epilogue staging lives in the operand ring");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct Params {
  int M, K, Beff;
  const float *bias, *g1, *b1, *g2, *b2;  // [512] each (g2/b2 unused without CHAIN)
  float ln_eps;
};

__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// tm_res: fp32 [M, 512] residual stream h (box 32 x 16, SWIZZLE_64B) -- used for the residual LOAD and the h STORE
// tm_c  : fp32 [Beff + 32, 512] cyclic per-sample constant (row r = c[r % Beff]); only read with CHAIN
// tm_ohi / tm_olo: bf16 [M, 512] split of h (store)
template <bool SPLIT, bool CHAIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
               const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_c,
               const __grid_constant__ CUtensorMap tm_ohi, const __grid_constant__ CUtensorMap tm_olo, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* s_par = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);             // bias | g1 | b1 | g2 | b2
  float2* s_stats = reinterpret_cast<float2*>(smem + STAGES * STAGE_BYTES + PARAM_BYTES);  // [2 exchanges][128 rows][2 halves]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + PARAM_BYTES + STATS_BYTES);
  uint64_t* full_bar = bars;          // [2]
  uint64_t* empty_bar = bars + 2;     // [2]
  uint64_t* tmem_full_bar = bars + 4;
  uint64_t* tmem_empty_bar = bars + 5;   // leader's copy: 16 arrivals (epilogue warps of both CTAs)
  uint64_t* epi_done_bar = bars + 6;     // local: 8 arrivals, operand ring free again for the producer
  uint64_t* ring_bar = bars + 8;         // [8 warps][8]: res slots 0..3, c slots 4..7
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 8 + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_kb = p.K / BK;
  const int num_tiles = (p.M + 2 * BM - 1) / (2 * BM);

  // LayerNorm / bias vectors -> shared memory (global loads are L2 round trips here: there is no L1 left)
  for (int i = threadIdx.x; i < 5 * ND / 4; i += kThreads) {
    const int which = i / (ND / 4), j = i % (ND / 4);
    const float* src = which == 0 ? p.bias : which == 1 ? p.g1 : which == 2 ? p.b1 : which == 3 ? p.g2 : p.b2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src) v = __ldg(reinterpret_cast<const float4*>(src) + j);
    reinterpret_cast<float4*>(s_par)[i] = v;
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a_hi);
    ptx::prefetch_tmap(&tm_w_hi);
    ptx::prefetch_tmap(&tm_res);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::mbar_init(tmem_empty_bar, 16);
    ptx::mbar_init(epi_done_bar, 8);
    for (int i = 0; i < 64; ++i) ptx::mbar_init(&ring_bar[i], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(tmem_base_smem, 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int m0 = tile * (2 * BM) + (int)rank * BM;
        if (it > 0) ptx::mbar_wait(epi_done_bar, (it - 1) & 1);  // the epilogue staged in the operand ring
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * STAGE_BYTES;
          if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * (SPLIT ? STAGE_BYTES : STAGE_BYTES / 2));
          ptx::tma_load_2d_2sm(st, &tm_a_hi, &full_bar[stage], kb * BK, m0);
          ptx::tma_load_2d_2sm(st + 2 * 16384, &tm_w_hi, &full_bar[stage], kb * BK, (int)rank * 128);
          ptx::tma_load_2d_2sm(st + 3 * 16384, &tm_w_hi, &full_bar[stage], kb * BK, 256 + (int)rank * 128);
          if (SPLIT) {
            ptx::tma_load_2d_2sm(st + 16384, &tm_a_lo, &full_bar[stage], kb * BK, m0);
            ptx::tma_load_2d_2sm(st + 4 * 16384, &tm_w_lo, &full_bar[stage], kb * BK, (int)rank * 128);
            ptx::tma_load_2d_2sm(st + 5 * 16384, &tm_w_lo, &full_bar[stage], kb * BK, 256 + (int)rank * 128);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(2 * BM, 256);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        ptx::mbar_wait(tmem_empty_bar, (it & 1) ^ 1);  // both CTAs' epilogues are done with the accumulator
        ptx::tcgen05_fence_after();
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tcgen05_fence_after();
          const uint32_t st = ptx::smem_u32(smem + stage * STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t adv = (uint32_t)k * 32;
            const uint64_t a_hi = ptx::umma_desc_k_sw128(st + adv);
            const uint64_t a_lo = ptx::umma_desc_k_sw128(st + 16384 + adv);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t w_hi = ptx::umma_desc_k_sw128(st + (2 + j) * 16384 + adv);
              const uint64_t w_lo = ptx::umma_desc_k_sw128(st + (4 + j) * 16384 + adv);
              const uint32_t acc = tmem_base + (uint32_t)(j * 256);
              if (SPLIT) {
                ptx::mma_f16_ss_2sm(acc, a_lo, w_hi, idesc, (kb | k) != 0);
                ptx::mma_f16_ss_2sm(acc, a_hi, w_lo, idesc, 1);
                ptx::mma_f16_ss_2sm(acc, a_hi, w_hi, idesc, 1);
              } else {
                ptx::mma_f16_ss_2sm(acc, a_hi, w_hi, idesc, (kb | k) != 0);
              }
            }
          }
          ptx::tcgen05_commit_2sm(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::tcgen05_commit_2sm(tmem_full_bar);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warps 2..9, one thread = one row
    const int ew = warp - 2;
    const int q = warp & 3, hf = ew >> 2;
    const int r_local = q * 32 + lane;                 // row inside this CTA's 128 rows == TMEM lane
    uint8_t* my = smem + ew * EPI_WARP_BYTES;
    uint8_t* res_ring = my;
    uint8_t* c_ring = my + 8192;
    uint8_t* out_buf = my + 16384;
    uint64_t* rbar = ring_bar + ew * 8;
    const float* s_bias = s_par + hf * 256;
    const float* s_g1 = s_par + ND + hf * 256;
    const float* s_b1 = s_par + 2 * ND + hf * 256;
    const float* s_g2 = s_par + 3 * ND + hf * 256;
    const float* s_b2 = s_par + 4 * ND + hf * 256;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 256);
    const float inv_n = 1.0f / ND;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int row0 = tile * (2 * BM) + (int)rank * BM + q * 32;  // global row of lane 0
      const int n_base = hf * 256;
      const int crow0 = row0 % p.Beff;                              // first row in the cyclic c table
      const uint32_t ring_phase0 = (uint32_t)(it * 4);              // each slot is filled 4 times per tile per ring
      ptx::mbar_wait(tmem_full_bar, it & 1);                        // accumulator complete => operand ring is idle
      ptx::tcgen05_fence_after();
      // prime the residual (and c) rings: 4 sub-chunks each
      if (lane == 0) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          ptx::mbar_expect_tx(&rbar[s], 2048);
          ptx::tma_load_2d(res_ring + s * 2048, &tm_res, &rbar[s], n_base + 16 * s, row0);
          if (CHAIN) {
            ptx::mbar_expect_tx(&rbar[4 + s], 2048);
            ptx::tma_load_2d(c_ring + s * 2048, &tm_c, &rbar[4 + s], n_base + 16 * s, crow0);
          }
        }
      }
      // ---- pass 1: v = acc + bias + residual, row statistics, v -> TMEM
      float sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int sc = 0; sc < 16; ++sc) {
        const int slot = sc & 3;
        uint32_t r[16];
        __syncwarp();
        ptx::tmem_ld_32x32b_x16(lane_addr + (uint32_t)(sc * 16), r);
        ptx::mbar_wait(&rbar[slot], (ring_phase0 + (uint32_t)(sc >> 2)) & 1);
        float4 rr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = *reinterpret_cast<const float4*>(res_ring + slot * 2048 + gemm::stg_off_f32(lane, j));
        ptx::tmem_ld_wait();
        __syncwarp();  // every lane has read the slot
        if (lane == 0 && sc + 4 < 16) {
          ptx::mbar_expect_tx(&rbar[slot], 2048);
          ptx::tma_load_2d(res_ring + slot * 2048, &tm_res, &rbar[slot], n_base + 16 * (sc + 4), row0);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + sc * 16 + 4 * j);
          float v0 = __uint_as_float(r[4 * j]) + b4.x + rr[j].x, v1 = __uint_as_float(r[4 * j + 1]) + b4.y + rr[j].y;
          float v2 = __uint_as_float(r[4 * j + 2]) + b4.z + rr[j].z, v3 = __uint_as_float(r[4 * j + 3]) + b4.w + rr[j].w;
          sum += (v0 + v1) + (v2 + v3);
          sq = fmaf(v0, v0, sq); sq = fmaf(v1, v1, sq); sq = fmaf(v2, v2, sq); sq = fmaf(v3, v3, sq);
          r[4 * j] = __float_as_uint(v0); r[4 * j + 1] = __float_as_uint(v1);
          r[4 * j + 2] = __float_as_uint(v2); r[4 * j + 3] = __float_as_uint(v3);
        }
        tmem_st_32x32b_x16(lane_addr + (uint32_t)(sc * 16), r);
      }
      tmem_st_wait();
      // exchange the half-row statistics with the warp that owns the other 256 columns of the same rows
      s_stats[r_local * 2 + hf] = make_float2(sum, sq);
      named_bar_sync(1 + q, 64);
      {
        const float2 o = s_stats[r_local * 2 + (hf ^ 1)];
        sum += o.x;
        sq += o.y;
      }
      float mean = sum * inv_n;
      float rstd = 1.0f / sqrtf(fmaxf(sq * inv_n - mean * mean, 0.f) + p.ln_eps);

      if (CHAIN) {
        // ---- pass 2: y = LN1(v) + c, statistics of y, y -> TMEM
        float sum2 = 0.f, sq2 = 0.f;
#pragma unroll 1
        for (int sc = 0; sc < 16; ++sc) {
          const int slot = sc & 3;
          uint32_t r[16];
          __syncwarp();
          ptx::tmem_ld_32x32b_x16(lane_addr + (uint32_t)(sc * 16), r);
          ptx::mbar_wait(&rbar[4 + slot], (ring_phase0 + (uint32_t)(sc >> 2)) & 1);
          float4 cc[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) cc[j] = *reinterpret_cast<const float4*>(c_ring + slot * 2048 + gemm::stg_off_f32(lane, j));
          ptx::tmem_ld_wait();
          __syncwarp();
          if (lane == 0 && sc + 4 < 16) {
            ptx::mbar_expect_tx(&rbar[4 + slot], 2048);
            ptx::tma_load_2d(c_ring + slot * 2048, &tm_c, &rbar[4 + slot], n_base + 16 * (sc + 4), crow0);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 g4 = *reinterpret_cast<const float4*>(s_g1 + sc * 16 + 4 * j);
            const float4 b4 = *reinterpret_cast<const float4*>(s_b1 + sc * 16 + 4 * j);
            float y0 = (__uint_as_float(r[4 * j]) - mean) * rstd * g4.x + b4.x + cc[j].x;
            float y1 = (__uint_as_float(r[4 * j + 1]) - mean) * rstd * g4.y + b4.y + cc[j].y;
            float y2 = (__uint_as_float(r[4 * j + 2]) - mean) * rstd * g4.z + b4.z + cc[j].z;
            float y3 = (__uint_as_float(r[4 * j + 3]) - mean) * rstd * g4.w + b4.w + cc[j].w;
            sum2 += (y0 + y1) + (y2 + y3);
            sq2 = fmaf(y0, y0, sq2); sq2 = fmaf(y1, y1, sq2); sq2 = fmaf(y2, y2, sq2); sq2 = fmaf(y3, y3, sq2);
            r[4 * j] = __float_as_uint(y0); r[4 * j + 1] = __float_as_uint(y1);
            r[4 * j + 2] = __float_as_uint(y2); r[4 * j + 3] = __float_as_uint(y3);
          }
          tmem_st_32x32b_x16(lane_addr + (uint32_t)(sc * 16), r);
        }
        tmem_st_wait();
        s_stats[256 + r_local * 2 + hf] = make_float2(sum2, sq2);
        named_bar_sync(1 + q, 64);
        {
          const float2 o = s_stats[256 + r_local * 2 + (hf ^ 1)];
          sum2 += o.x;
          sq2 += o.y;
        }
        mean = sum2 * inv_n;
        rstd = 1.0f / sqrtf(fmaxf(sq2 * inv_n - mean * mean, 0.f) + p.ln_eps);
      }

      // ---- final pass: z = LN(.), fp32 + bf16 (hi, lo) out through TMA stores
      const float* gg = CHAIN ? s_g2 : s_g1;
      const float* bb = CHAIN ? s_b2 : s_b1;
      if (lane == 0) ptx::bulk_wait_read<0>();  // output buffers of the previous tile
#pragma unroll 1
      for (int sc = 0; sc < 16; ++sc) {
        uint8_t* ob = out_buf + (sc & 1) * 4096;
        if (sc >= 2 && lane == 0) ptx::bulk_wait_read<1>();
        uint32_t r[16];
        __syncwarp();
        ptx::tmem_ld_32x32b_x16(lane_addr + (uint32_t)(sc * 16), r);
        ptx::tmem_ld_wait();
        float z[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 g4 = *reinterpret_cast<const float4*>(gg + sc * 16 + 4 * j);
          const float4 b4 = *reinterpret_cast<const float4*>(bb + sc * 16 + 4 * j);
          z[4 * j] = (__uint_as_float(r[4 * j]) - mean) * rstd * g4.x + b4.x;
          z[4 * j + 1] = (__uint_as_float(r[4 * j + 1]) - mean) * rstd * g4.y + b4.y;
          z[4 * j + 2] = (__uint_as_float(r[4 * j + 2]) - mean) * rstd * g4.z + b4.z;
          z[4 * j + 3] = (__uint_as_float(r[4 * j + 3]) - mean) * rstd * g4.w + b4.w;
          *reinterpret_cast<float4*>(ob + gemm::stg_off_f32(lane, j)) = make_float4(z[4 * j], z[4 * j + 1], z[4 * j + 2], z[4 * j + 3]);
        }
        uint32_t hw[8], lw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          hw[j] = gemm::pack_bf16x2(z[2 * j], z[2 * j + 1]);
          lw[j] = gemm::pack_bf16x2(z[2 * j] - __uint_as_float(hw[j] << 16), z[2 * j + 1] - __uint_as_float(hw[j] & 0xffff0000u));
        }
        *reinterpret_cast<uint4*>(ob + 2048 + gemm::stg_off_bf16(lane, 0)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(ob + 2048 + gemm::stg_off_bf16(lane, 1)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
        *reinterpret_cast<uint4*>(ob + 3072 + gemm::stg_off_bf16(lane, 0)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        *reinterpret_cast<uint4*>(ob + 3072 + gemm::stg_off_bf16(lane, 1)) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_2d(&tm_res, ob, n_base + 16 * sc, row0);
          ptx::tma_store_2d(&tm_ohi, ob + 2048, n_base + 16 * sc, row0);
          ptx::tma_store_2d(&tm_olo, ob + 3072, n_base + 16 * sc, row0);
          ptx::bulk_commit();
        }
      }
      // accumulator and operand ring are free again
      if (lane == 0) ptx::bulk_wait_read<0>();
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive_remote(tmem_empty_bar, 0);
        ptx::mbar_arrive(epi_done_bar);
      }
      named_bar_sync(1 + q, 64);  // the statistics slots are rewritten by the next tile
    }
    if (lane == 0) ptx::bulk_wait<0>();
  }

  ptx::tcgen05_fence_before();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <bool SPLIT, bool CHAIN>
inline cudaError_t launch(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                          const CUtensorMap& w_lo, const CUtensorMap& res, const CUtensorMap& c, const CUtensorMap& ohi,
                          const CUtensorMap& olo, const Params& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ln_kernel<SPLIT, CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int64_t tiles = ceil_div(p.M, 2 * BM);
  const int clusters = (int)(tiles < kNumSMs / 2 ? tiles : kNumSMs / 2);
  gemm_ln_kernel<SPLIT, CHAIN><<<2 * clusters, kThreads, SMEM_BYTES, stream>>>(a_hi, a_lo, w_hi, w_lo, res, c, ohi, olo, p);
  return cudaGetLastError();
}

}  // namespace gemmln
}  // namespace regen
