// Causal self-attention on tcgen05 tensor cores (sm_100a), bf16x3 operands, fp32 softmax.
//
// One CTA per (sample b, head h, 128-query block).  q | k | v arrive as bf16 (hi, lo) pairs
// [T*Beff, 3*512] (rows seq-first, written by the QKV GEMM epilogue) and are fetched with 3-D TMA boxes
// {64 d, 1 sample, TB frames} straight into 128-byte-swizzled UMMA operand tiles:
//
//   S[128 x TB]  = Q[128 x 128] . K_chunk[TB x 128]^T      (A, B K-major; 8 k-steps x 3 split MMAs)
//   P            = exp2((S - rowmax) * scale*log2e), causal-masked, written to shared memory as bf16 (hi, lo)
//                  in the K-major swizzled A-operand layout (one thread = one query row: max / sum are
//                  thread-local, no shuffles)
//   O[128 x 128] += P[128 x TB] . V_chunk[TB x 128]        (B operand MN-major: V tiles are used as loaded)
//   out          = O / rowsum  -> bf16 (hi, lo) [T*Beff, 512], the A operand of the output projection.
//
// TB = 64 (T <= 64: ~96 KB of shared memory, two CTAs per SM) or 128 (longer sequences, key chunks of 128 with the
// scores of all chunks resident in TMEM so the row max is exact before any exponential is taken).
// Reference semantics: nn.MultiheadAttention with the additive causal mask of model/cmdm.py:168-171, 220-227.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace regen {
namespace attn {

constexpr int HD = 128;
constexpr int DM = 512;
constexpr int kThreads = 160;  // warp 0: TMA + MMA control, warps 1..4: softmax / epilogue

template <int TB>
struct Cfg {
  static constexpr int TILE = TB * 128;            // TB rows x 64 bf16
  static constexpr int OPERAND = 4 * TILE;         // hi d0-63 | hi d64-127 | lo d0-63 | lo d64-127
  static constexpr int P_TILE = 128 * 128;         // 128 query rows x 64 keys
  static constexpr int P_BYTES = 2 * (TB / 64) * P_TILE;
  static_assert(P_BYTES == OPERAND, "P aliases the Q operand region");
  static constexpr int SMEM_BYTES = 3 * OPERAND + 256 + 1024;
};

// UMMA smem descriptor, MN-major operand, 128-byte swizzle: 64 contiguous MN elements per 128-byte row, rows
// run along K; 8-row groups every SBO = 1024 bytes, next 64-element MN block at LBO bytes.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                      uint32_t sbo_bytes = 1024) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct Params {
  __nv_bfloat16* out_hi;  // [T*Beff, 512]
  __nv_bfloat16* out_lo;
  int T, Beff;
  int dbg;  // test-hook only: bit 0 swaps the LBO / SBO fields of the V descriptor (bring-up A/B switch)
};

template <int TB>
__global__ void __launch_bounds__(kThreads, TB == 64 ? 2 : 1)
attention_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const Params p) {
  using C = Cfg<TB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                  // later: P
  uint8_t* sK = smem + C::OPERAND;
  uint8_t* sV = smem + 2 * C::OPERAND;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * C::OPERAND);
  uint64_t* barQ = bars + 0;
  uint64_t* barK = bars + 1;
  uint64_t* barV = bars + 2;
  uint64_t* barS = bars + 3;   // [2] S chunk kc complete (tcgen05.commit); one barrier per chunk because the
                               //     softmax threads may arrive after several chunks have completed
  uint64_t* barP = bars + 5;   // P chunk written by the 128 softmax threads
  uint64_t* barO = bars + 6;   // P.V chunk complete (tcgen05.commit)
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblocks = (p.T + 127) / 128;
  const int qb = blockIdx.x % qblocks;
  const int bh = blockIdx.x / qblocks;
  const int h = bh & 3, b = bh >> 2;
  const int q0 = qb * 128;                                   // first query frame of this CTA
  const int kv_len = min(p.T, q0 + 128);                     // causal: keys [0, kv_len)
  const int nkc = (kv_len + TB - 1) / TB;                    // key chunks
  constexpr uint32_t TMEM_COLS = TB == 64 ? 256 : 512;       // O: 128 cols, S: up to 256 cols
  constexpr uint32_t O_COL = 0, S_COL = 128;

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tm_hi);
      ptx::prefetch_tmap(&tm_lo);
      ptx::mbar_init(barQ, 1);
      ptx::mbar_init(barK, 1);
      ptx::mbar_init(barV, 1);
      ptx::mbar_init(&barS[0], 1);
      ptx::mbar_init(&barS[1], 1);
      ptx::mbar_init(barP, 128);
      ptx::mbar_init(barO, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_base_smem, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ control thread: TMA + MMA issue
      auto load_operand = [&](uint8_t* dst, uint64_t* bar, int col0, int t0) {
        ptx::mbar_expect_tx(bar, C::OPERAND);
        ptx::tma_load_3d(dst, &tm_hi, bar, col0, b, t0);
        ptx::tma_load_3d(dst + C::TILE, &tm_hi, bar, col0 + 64, b, t0);
        ptx::tma_load_3d(dst + 2 * C::TILE, &tm_lo, bar, col0, b, t0);
        ptx::tma_load_3d(dst + 3 * C::TILE, &tm_lo, bar, col0 + 64, b, t0);
      };
      load_operand(sQ, barQ, h * HD, q0);
      load_operand(sK, barK, DM + h * HD, 0);
      load_operand(sV, barV, 2 * DM + h * HD, 0);

      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16_f32(128, TB);
      // P.V: B operand (V) is MN-major -> b_major bit 16
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16_f32(128, HD) | (1u << 16);
      const uint32_t aQ = ptx::smem_u32(sQ), aK = ptx::smem_u32(sK), aV = ptx::smem_u32(sV);

      // ---- scores: S_kc = Q . K_kc^T for every key chunk (all chunks stay resident in TMEM)
      ptx::mbar_wait(barQ, 0);
      for (int kc = 0; kc < nkc; ++kc) {
        ptx::mbar_wait(barK, kc & 1);
        ptx::tcgen05_fence_after();
        const uint32_t accS = tmem_base + S_COL + (uint32_t)(kc * TB);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t tile = (uint32_t)(k >> 2) * C::TILE;        // d 0-63 | d 64-127
          const uint32_t adv = (uint32_t)(k & 3) * 32;               // 16 bf16 inside the swizzle row
          const uint64_t q_hi = ptx::umma_desc_k_sw128(aQ + tile + adv);
          const uint64_t q_lo = ptx::umma_desc_k_sw128(aQ + 2 * C::TILE + tile + adv);
          const uint64_t k_hi = ptx::umma_desc_k_sw128(aK + tile + adv);
          const uint64_t k_lo = ptx::umma_desc_k_sw128(aK + 2 * C::TILE + tile + adv);
          ptx::mma_f16_ss(accS, q_lo, k_hi, idesc_s, k != 0);
          ptx::mma_f16_ss(accS, q_hi, k_lo, idesc_s, 1);
          ptx::mma_f16_ss(accS, q_hi, k_hi, idesc_s, 1);
        }
        ptx::tcgen05_commit(&barS[kc]);
        ptx::mbar_wait(&barS[kc], 0);  // K buffer (and finally Q) free again
        if (kc + 1 < nkc) load_operand(sK, barK, DM + h * HD, (kc + 1) * TB);
      }
      // ---- O += P_kc . V_kc
      for (int kc = 0; kc < nkc; ++kc) {
        ptx::mbar_wait(barP, kc & 1);  // P chunk is in shared memory (written through the generic proxy + fence)
        ptx::mbar_wait(barV, kc & 1);
        ptx::tcgen05_fence_after();
        const uint32_t accO = tmem_base + O_COL;
#pragma unroll
        for (int k = 0; k < TB / 16; ++k) {
          // A = P: [128 rows x 64 keys] tiles, K-major; k-step = 16 keys
          const uint32_t ptile = (uint32_t)(k >> 2) * C::P_TILE + (uint32_t)(k & 3) * 32;
          const uint64_t p_hi = ptx::umma_desc_k_sw128(aQ + ptile);
          const uint64_t p_lo = ptx::umma_desc_k_sw128(aQ + (TB / 64) * C::P_TILE + ptile);
          // B = V: [TB keys x 64 d] tiles; 16 keys = 16 rows of 128 bytes; second d half at +TILE
          const uint32_t lbo = (p.dbg & 1) ? 1024u : (uint32_t)C::TILE, sbo = (p.dbg & 1) ? (uint32_t)C::TILE : 1024u;
          const uint64_t v_hi = umma_desc_mn_sw128(aV + (uint32_t)k * 2048, lbo, sbo);
          const uint64_t v_lo = umma_desc_mn_sw128(aV + 2 * C::TILE + (uint32_t)k * 2048, lbo, sbo);
          ptx::mma_f16_ss(accO, p_lo, v_hi, idesc_o, (kc | k) != 0);
          ptx::mma_f16_ss(accO, p_hi, v_lo, idesc_o, 1);
          ptx::mma_f16_ss(accO, p_hi, v_hi, idesc_o, 1);
        }
        ptx::tcgen05_commit(barO);
        if (kc + 1 < nkc) {
          ptx::mbar_wait(barO, kc & 1);  // V and P buffers free
          load_operand(sV, barV, 2 * DM + h * HD, (kc + 1) * TB);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue: one thread per query row
    const int q = warp & 3;
    const int r = q * 32 + lane;           // row in the 128-row tile == TMEM lane
    const int i = q0 + r;                  // query frame
    const bool row_ok = i < p.T && r < (TB == 64 ? 64 : 128);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sc = 0.08838834764831845f * 1.4426950408889634f;  // 1/sqrt(128) * log2(e)

    // all score chunks complete
    for (int kc = 0; kc < nkc; ++kc) ptx::mbar_wait(&barS[kc], 0);
    ptx::tcgen05_fence_after();
    // pass 1: exact row maximum over the causal window
    float mx = -INFINITY;
    for (int c0 = 0; c0 < nkc * TB; c0 += 32) {
      uint32_t v[32];
      __syncwarp();
      ptx::tmem_ld_32x32b_x32(lane_addr + S_COL + (uint32_t)c0, v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j <= i && c0 + j < p.T) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    if (!row_ok) mx = 0.f;
    float sum = 0.f;
    for (int kc = 0; kc < nkc; ++kc) {
      if (kc > 0) ptx::mbar_wait(barO, (kc - 1) & 1);  // previous P chunk consumed by the tensor core
      for (int c0 = 0; c0 < TB; c0 += 32) {
        uint32_t v[32];
        __syncwarp();
        ptx::tmem_ld_32x32b_x32(lane_addr + S_COL + (uint32_t)(kc * TB + c0), v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j8 = 0; j8 < 32; j8 += 8) {
          __nv_bfloat16 hi8[8], lo8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = kc * TB + c0 + j8 + e;
            float pv = 0.f;
            if (row_ok && j <= i) pv = exp2f((__uint_as_float(v[j8 + e]) - mx) * sc);
            sum += pv;
            split_bf16(pv, hi8[e], lo8[e]);
          }
          // K-major SW128 A-operand layout: 16-byte chunk index XOR (row & 7)
          const int jj = c0 + j8;  // key offset inside the chunk
          const uint32_t off = (uint32_t)(jj >> 6) * C::P_TILE + (uint32_t)r * 128 +
                               ((((uint32_t)(jj & 63) >> 3) ^ ((uint32_t)r & 7)) << 4);
          *reinterpret_cast<uint4*>(sQ + off) = *reinterpret_cast<uint4*>(hi8);
          *reinterpret_cast<uint4*>(sQ + (TB / 64) * C::P_TILE + off) = *reinterpret_cast<uint4*>(lo8);
        }
      }
      ptx::fence_proxy_async_smem();  // make the generic-proxy stores visible to the tensor core (async proxy)
      ptx::mbar_arrive(barP);
    }
    // epilogue: O / rowsum -> bf16 (hi, lo)
    ptx::mbar_wait(barO, (nkc - 1) & 1);
    ptx::tcgen05_fence_after();
    const float inv = 1.f / sum;
    const size_t grow = ((size_t)i * p.Beff + b) * DM + (size_t)h * HD;
    for (int c0 = 0; c0 < HD; c0 += 32) {
      uint32_t v[32];
      __syncwarp();
      ptx::tmem_ld_32x32b_x32(lane_addr + O_COL + (uint32_t)c0, v);
      ptx::tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int j8 = 0; j8 < 32; j8 += 8) {
          __nv_bfloat16 hi8[8], lo8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) split_bf16(__uint_as_float(v[j8 + e]) * inv, hi8[e], lo8[e]);
          *reinterpret_cast<uint4*>(p.out_hi + grow + c0 + j8) = *reinterpret_cast<uint4*>(hi8);
          *reinterpret_cast<uint4*>(p.out_lo + grow + c0 + j8) = *reinterpret_cast<uint4*>(lo8);
        }
      }
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int TB>
inline cudaError_t launch(const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, const Params& p, cudaStream_t s) {
  using C = Cfg<TB>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (p.T > (TB == 64 ? 64 : 256)) return cudaErrorInvalidValue;  // at most two key chunks of 128
  const int qblocks = (p.T + 127) / 128;
  attention_kernel<TB><<<p.Beff * 4 * qblocks, kThreads, C::SMEM_BYTES, s>>>(tm_hi, tm_lo, p);
  return cudaGetLastError();
}

}  // namespace attn
}  // namespace regen
