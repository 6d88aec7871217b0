// Self-attention on tcgen05 tensor cores (sm_100a), bf16x3 operands, fp32 softmax; causal (arch 'online') or unmasked
// (arch 'offline', Params::causal = 0).
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
// TB = 64 (T <= 64: K and V share a buffer, O reuses S's TMEM columns: ~69 KB of shared memory, 128 TMEM columns, three
// CTAs per SM) or 128 (longer sequences, one CTA per SM, key chunks of 128 with the
// scores of all chunks resident in TMEM so the row max is exact before any exponential is taken).
// A third kernel (attention_mc_kernel, below) covers 64 < T <= 256 with the compact footprint (64-query blocks, 64-key
// chunks, two CTAs per SM); model.cu picks it when the last 128-query block would be at most half full.
// Reference semantics: nn.MultiheadAttention with the additive causal mask of model/cmdm.py:168-171, 220-227.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace regen {
namespace attn {

constexpr int HD = 128;
constexpr int DM = 512;

template <int TB>
struct Cfg {
  static constexpr int TILE = TB * 128;            // TB rows x 64 bf16
  static constexpr int OPERAND = 4 * TILE;         // hi d0-63 | hi d64-127 | lo d0-63 | lo d64-127
  static constexpr int P_TILE = 128 * 128;         // 128 query rows x 64 keys
  static constexpr int P_BYTES = 2 * (TB / 64) * P_TILE;
  static_assert(P_BYTES == OPERAND, "P aliases the Q operand region");
  // TB = 64 ("compact"): K and V share one buffer (V is fetched while the softmax runs) and O reuses the TMEM columns
  // of S, so a CTA needs ~69 KB of shared memory, 128 TMEM columns and 288 threads: three CTAs per SM.
  static constexpr bool COMPACT = TB == 64;
  // softmax / epilogue warps: two per TMEM lane quarter, each owns half of a chunk's key columns (softmax) and half of
  // the head dimension (epilogue).  The compact kernel ran 4 warps (one per quarter, 64 scores per thread) in the first
  // version; two warps per quarter halve the per-thread softmax and read-out work on the CTA's latency chain.
  static constexpr int SW = 8;
  static constexpr int THREADS = 32 + 32 * SW;               // + warp 0: TMA + MMA control
  static constexpr int SMEM_BYTES = (COMPACT ? 2 : 3) * OPERAND + 4096 + 1024;  // + barriers / row statistics
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
  __nv_bfloat16* out_hi;  // [T*Beff, 512] (the kernel stores through tm_ohi / tm_olo)
  __nv_bfloat16* out_lo;
  int T, Beff;
  int causal;  // 1: additive causal mask of model/cmdm.py:168-171 (arch 'online'); 0: no mask (arch 'offline')
  int m8;   // 1: outputs in the mixed8 operand format of the fused out_proj + LayerNorm kernel (gemm_ln_sm100.cuh): tm_ohi
            // stores fp16 (same [T, Beff, 512] 16-bit map), tm_olo is the byte map [T, Beff, 1024] (box 32 frames x 128 B,
            // SWIZZLE_128B): per group of 64 columns, 64 bytes e4m3((o - fp16(o)) * 2^9) | 64 bytes e4m3(fp16(o) / 4)
  int dbg;  // bit 0 (test hook only) swaps the LBO / SBO fields of the V descriptor (bring-up A/B switch);
            // bit 2: wait for the bulk stores' global writes before exit (default; REGEN_DEBUG_EXIT_WAIT_READ=1 clears it, A/B: no measurable difference)
  unsigned long long* timeline;  // bring-up instrumentation (null in production): CTA 0 stamps clock64() at events
  unsigned long long* steplog;   // whole-step timeline (ptx::steplog_begin / steplog_end), null in production
  int steplog_slot, steplog_cta;  // steplog_cta: word offset of the per-CTA exit-time table (0 = off)
  // L2 eviction-priority hints (ptx::kL2Evict*, 0 = none): q | k | v tiles are dead once read, the output is the next
  // kernel's A operand
  unsigned long long pol_load, pol_store;
};

#define REGEN_ATL(k)                                                                      \
  do {                                                                                    \
    if (p.timeline && blockIdx.x == 0) p.timeline[(k)] = (unsigned long long)clock64();   \
  } while (0)

// 32 accumulator columns of one output row -> staged operand chunks.  bf16 pair: four 16-byte chunks of the hi tile and
// of the lo tile (32 rows x 128 B, SWIZZLE_128B, chunk index chunk0 + i).  mixed8: the same four chunks of the fp16 tile at
// st_hi, and one 16-byte chunk in each half of the byte tile at st_lo (32 rows x 128 B, SWIZZLE_128B: 64 residual bytes |
// 64 hi bytes per row).
__device__ __forceinline__ void stage_out32(const uint32_t (&v)[32], float inv, uint8_t* st_hi, uint8_t* st_lo, int lane,
                                            int c0, bool m8) {
  if (m8) {
#pragma unroll
    for (int j16 = 0; j16 < 32; j16 += 16) {
      uint32_t hw[8], l8[4], h8[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {  // 4 columns -> two fp16x2 words, one word of residual bytes, one word of hi bytes
        float lo[4], hf[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float a = __uint_as_float(v[j16 + 4 * g + 2 * e]) * inv, b2 = __uint_as_float(v[j16 + 4 * g + 2 * e + 1]) * inv;
          hw[2 * g + e] = ptx::pack_f16x2_sat(a, b2);
          const ptx::f32x2 h2 = ptx::f16x2_to_f32x2(hw[2 * g + e]);
          ptx::upk2(ptx::mul2(h2, ptx::splat2(0.25f)), hf[2 * e], hf[2 * e + 1]);
          ptx::upk2(ptx::mul2(ptx::sub2(ptx::pk2(a, b2), h2), ptx::splat2(512.f)), lo[2 * e], lo[2 * e + 1]);
        }
        l8[g] = ptx::pack_e4m3x4(lo[0], lo[1], lo[2], lo[3]);
        h8[g] = ptx::pack_e4m3x4(hf[0], hf[1], hf[2], hf[3]);
      }
      const int col = (c0 & 63) + j16;  // column inside the warp's 64-column tile
      const uint32_t off16 = (uint32_t)lane * 128;
      *reinterpret_cast<uint4*>(st_hi + off16 + ((((uint32_t)col >> 3) ^ ((uint32_t)lane & 7)) << 4)) =
          make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(st_hi + off16 + (((((uint32_t)col >> 3) + 1) ^ ((uint32_t)lane & 7)) << 4)) =
          make_uint4(hw[4], hw[5], hw[6], hw[7]);
      // byte tile: 32 rows x 128 B (SWIZZLE_128B), residual bytes in chunks 0..3, hi bytes in chunks 4..7
      *reinterpret_cast<uint4*>(st_lo + off16 + ((((uint32_t)col >> 4) ^ ((uint32_t)lane & 7)) << 4)) =
          make_uint4(l8[0], l8[1], l8[2], l8[3]);
      *reinterpret_cast<uint4*>(st_lo + off16 + (((4 + ((uint32_t)col >> 4)) ^ ((uint32_t)lane & 7)) << 4)) =
          make_uint4(h8[0], h8[1], h8[2], h8[3]);
    }
    return;
  }
#pragma unroll
  for (int j8 = 0; j8 < 32; j8 += 8) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = __uint_as_float(v[j8 + 2 * e]) * inv, b2 = __uint_as_float(v[j8 + 2 * e + 1]) * inv;
      __nv_bfloat162 hh = __floats2bfloat162_rn(a, b2);
      hw[e] = *reinterpret_cast<uint32_t*>(&hh);
      __nv_bfloat162 ll = __floats2bfloat162_rn(a - __uint_as_float(hw[e] << 16), b2 - __uint_as_float(hw[e] & 0xffff0000u));
      lw[e] = *reinterpret_cast<uint32_t*>(&ll);
    }
    const int chunk = ((c0 & 63) + j8) >> 3;  // 16-byte chunk of the 128-byte row
    const uint32_t off = (uint32_t)lane * 128 + (((uint32_t)chunk ^ ((uint32_t)lane & 7)) << 4);
    *reinterpret_cast<uint4*>(st_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(st_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

template <int TB>
__global__ void __launch_bounds__(Cfg<TB>::THREADS, TB == 64 ? 3 : 1)
attention_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                 const __grid_constant__ CUtensorMap tm_ohi, const __grid_constant__ CUtensorMap tm_olo, const Params p) {
  using C = Cfg<TB>;
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array (NOT an integer round trip): the compiler keeps the
  // shared address space and emits LDS / STS instead of generic LD / ST for every staging access below
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                  // later: P
  constexpr int NBUF = C::COMPACT ? 2 : 3;
  uint8_t* sK = smem + C::OPERAND;
  uint8_t* sV = C::COMPACT ? sK : smem + 2 * C::OPERAND;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NBUF * C::OPERAND);
  uint64_t* barQ = bars + 0;
  uint64_t* barK = bars + 1;
  uint64_t* barV = bars + 2;
  uint64_t* barS = bars + 3;   // [2] S chunk kc complete (tcgen05.commit); one barrier per chunk because the
                               //     softmax threads may arrive after several chunks have completed
  uint64_t* barP = bars + 5;   // P chunk written by the 128 softmax threads
  uint64_t* barO = bars + 6;   // P.V chunk complete (tcgen05.commit)
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 7);
  uint64_t* barK1 = bars + 8;  // TB = 128: second key chunk, loaded at kernel start into the (still free) V buffer
  uint64_t* barV1 = bars + 9;  // TB = 128: second value chunk (its load may be issued before the first has landed, so it
                               // cannot share barV: an arrive on a phase with no pending arrival is undefined)
  float* s_max = reinterpret_cast<float*>(smem + NBUF * C::OPERAND + 1024);   // [128 rows][2 key halves]
  float* s_sum = s_max + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblocks = (p.T + 127) / 128;
  const int qb = blockIdx.x % qblocks;
  const int bh = blockIdx.x / qblocks;
  const int h = bh & 3, b = bh >> 2;
  const int q0 = qb * 128;                                   // first query frame of this CTA
  const int kv_len = p.causal ? min(p.T, q0 + 128) : p.T;    // causal: keys [0, kv_len)
  const int nkc = (kv_len + TB - 1) / TB;                    // key chunks
  // O: 128 columns, S: TB columns per chunk.  Compact: O overwrites S (S is dead once P has been written).
  constexpr uint32_t TMEM_COLS = C::COMPACT ? 128 : 512;
  constexpr uint32_t O_COL = 0, S_COL = C::COMPACT ? 0 : 128;
  if (threadIdx.x == 0) REGEN_ATL(0);

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tm_hi);
      ptx::prefetch_tmap(&tm_lo);
      ptx::mbar_init(barQ, 1);
      ptx::mbar_init(barK, 1);
      ptx::mbar_init(barK1, 1);
      ptx::mbar_init(barV1, 1);
      ptx::mbar_init(barV, 1);
      ptx::mbar_init(&barS[0], 1);
      ptx::mbar_init(&barS[1], 1);
      ptx::mbar_init(barP, 32 * C::SW);
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
  ptx::griddep_wait();    // PDL: q | k | v come from the preceding QKV GEMM
  ptx::griddep_launch();  // the successor launches once every CTA of this grid has started
  ptx::steplog_begin(p.steplog, p.steplog_slot);
  if (threadIdx.x == 0) REGEN_ATL(1);

  if (warp == 0) {
    {
      // ------------------------------------------------------------------ control warp: TMA + MMA issue
      // The whole warp runs this code warp-uniformly and elects one lane per TMA / tcgen05 instruction (ptx::elect_one):
      // a loop inside `if (lane == 0)` costs more issue cycles per MMA than these small MMAs take to execute.
      auto load_operand = [&](uint8_t* dst, uint64_t* bar, int col0, int t0) {
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(bar, C::OPERAND);
          ptx::tma_load_3d(dst, &tm_hi, bar, col0, b, t0, p.pol_load);
          ptx::tma_load_3d(dst + C::TILE, &tm_hi, bar, col0 + 64, b, t0, p.pol_load);
          ptx::tma_load_3d(dst + 2 * C::TILE, &tm_lo, bar, col0, b, t0, p.pol_load);
          ptx::tma_load_3d(dst + 3 * C::TILE, &tm_lo, bar, col0 + 64, b, t0, p.pol_load);
        }
      };
      // Buffer rotation (TB = 128, two key chunks): K_0 -> sK and K_1 -> sV are both fetched at the start; V_0 takes sK as
      // soon as S_0 is complete and V_1 takes sV as soon as S_1 is, i.e. both V loads run behind the softmax passes and
      // no operand load sits between two MMA phases.  One key chunk: K_0 -> sK, V_0 -> sV at the start.
      const bool rotate = !C::COMPACT && nkc > 1;
      load_operand(sQ, barQ, h * HD, q0);
      load_operand(sK, barK, DM + h * HD, 0);
      if (rotate) load_operand(sV, barK1, DM + h * HD, TB);
      else if (!C::COMPACT) load_operand(sV, barV, 2 * DM + h * HD, 0);

      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16_f32(128, TB);
      // P.V: B operand (V) is MN-major -> b_major bit 16
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16_f32(128, HD) | (1u << 16);
      const uint32_t aQ = ptx::smem_u32(sQ), aK = ptx::smem_u32(sK), aV = ptx::smem_u32(sV);

      // ---- scores: S_kc = Q . K_kc^T for every key chunk (all chunks stay resident in TMEM)
      ptx::mbar_wait(barQ, 0);
      for (int kc = 0; kc < nkc; ++kc) {
        if (kc == 0) ptx::mbar_wait(barK, 0); else ptx::mbar_wait(barK1, 0);
        if (kc == 0) REGEN_ATL(2);
        ptx::tcgen05_fence_after();
        const uint32_t accS = tmem_base + S_COL + (uint32_t)(kc * TB);
        const uint32_t aKc = kc == 0 ? aK : aV;                      // K_1 sits in the V buffer
        // one elected lane issues the chunk's 24 MMAs back to back; descriptors: base of each operand tile built once,
        // + 2 in the address field per 32-byte k-step (the per-MMA issue cost, ~80 cycles when every MMA was elected and
        // its descriptors rebuilt, exceeded the 32 cycles such a 128 x 64 x 16 MMA executes)
        {
          const uint64_t dq_hi0 = ptx::umma_desc_k_sw128(aQ), dq_hi1 = ptx::umma_desc_k_sw128(aQ + C::TILE);
          const uint64_t dq_lo0 = ptx::umma_desc_k_sw128(aQ + 2 * C::TILE), dq_lo1 = ptx::umma_desc_k_sw128(aQ + 3 * C::TILE);
          const uint64_t dk_hi0 = ptx::umma_desc_k_sw128(aKc), dk_hi1 = ptx::umma_desc_k_sw128(aKc + C::TILE);
          const uint64_t dk_lo0 = ptx::umma_desc_k_sw128(aKc + 2 * C::TILE), dk_lo1 = ptx::umma_desc_k_sw128(aKc + 3 * C::TILE);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint64_t adv = (uint64_t)(2 * (k & 3));          // 16 bf16 = 32 bytes inside the swizzle row
              const uint64_t q_hi = ((k >> 2) ? dq_hi1 : dq_hi0) + adv, q_lo = ((k >> 2) ? dq_lo1 : dq_lo0) + adv;
              const uint64_t k_hi = ((k >> 2) ? dk_hi1 : dk_hi0) + adv, k_lo = ((k >> 2) ? dk_lo1 : dk_lo0) + adv;
              ptx::mma_f16_ss(accS, q_lo, k_hi, idesc_s, k != 0);
              ptx::mma_f16_ss(accS, q_hi, k_lo, idesc_s, 1);
              ptx::mma_f16_ss(accS, q_hi, k_hi, idesc_s, 1);
            }
          }
        }
        if (ptx::elect_one()) ptx::tcgen05_commit(&barS[kc]);
        ptx::mbar_wait(&barS[kc], 0);  // this chunk's K buffer (and finally Q) free again
        if (kc == 0) REGEN_ATL(3);
        if (rotate) load_operand(kc == 0 ? sK : sV, kc == 0 ? barV : barV1, 2 * DM + h * HD, kc * TB);  // V_kc -> K_kc's buffer
      }
      if (C::COMPACT) load_operand(sV, barV, 2 * DM + h * HD, 0);  // K is dead: V takes its buffer during the softmax
      // ---- O += P_kc . V_kc
      for (int kc = 0; kc < nkc; ++kc) {
        ptx::mbar_wait(barP, kc & 1);  // P chunk is in shared memory (written through the generic proxy + fence)
        if (kc == 0) ptx::mbar_wait(barV, 0); else ptx::mbar_wait(barV1, 0);
        if (kc == 0) REGEN_ATL(6);
        ptx::tcgen05_fence_after();
        const uint32_t accO = tmem_base + O_COL;
        const uint32_t aVc = (rotate && kc == 0) ? aK : aV;          // rotation: V_0 in sK, V_1 in sV
        {
          // A = P: [128 rows x 64 keys] tiles, K-major, k-step = 16 keys (+2 in the address field); B = V: [TB keys x 64 d]
          // tiles, 16 keys = 16 rows of 128 bytes (+128 in the address field), second d half at +TILE
          const uint32_t lbo = (p.dbg & 1) ? 1024u : (uint32_t)C::TILE, sbo = (p.dbg & 1) ? (uint32_t)C::TILE : 1024u;
          const uint64_t dv_hi = umma_desc_mn_sw128(aVc, lbo, sbo), dv_lo = umma_desc_mn_sw128(aVc + 2 * C::TILE, lbo, sbo);
          uint64_t dp_hi[TB / 64], dp_lo[TB / 64];
#pragma unroll
          for (int t = 0; t < TB / 64; ++t) {
            dp_hi[t] = ptx::umma_desc_k_sw128(aQ + (uint32_t)t * C::P_TILE);
            dp_lo[t] = ptx::umma_desc_k_sw128(aQ + (TB / 64) * C::P_TILE + (uint32_t)t * C::P_TILE);
          }
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < TB / 16; ++k) {
              const uint64_t p_hi = dp_hi[k >> 2] + (uint64_t)(2 * (k & 3)), p_lo = dp_lo[k >> 2] + (uint64_t)(2 * (k & 3));
              const uint64_t v_hi = dv_hi + (uint64_t)(128 * k), v_lo = dv_lo + (uint64_t)(128 * k);
              ptx::mma_f16_ss(accO, p_lo, v_hi, idesc_o, (kc | k) != 0);
              ptx::mma_f16_ss(accO, p_hi, v_lo, idesc_o, 1);
              ptx::mma_f16_ss(accO, p_hi, v_hi, idesc_o, 1);
            }
          }
        }
        if (ptx::elect_one()) ptx::tcgen05_commit(barO);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue
    // warps 1..8: TMEM lane quarter q = warp & 3, two warps per quarter; warp `half` owns key columns
    // [half*KH, half*KH + KH) of every chunk (softmax) and head-dim columns [half*64, half*64 + 64) (epilogue).
    constexpr int NH = C::SW / 4;          // warps per TMEM lane quarter
    constexpr int KH = TB / NH;            // keys per warp per chunk
    constexpr int DW = HD / NH;            // head-dim columns per warp in the epilogue
    const int q = warp & 3, half = (warp - 1) >> 2;
    const int r = q * 32 + lane;           // row in the 128-row tile == TMEM lane
    const int i = q0 + r;                  // query frame
    const bool row_ok = i < p.T && r < (TB == 64 ? 64 : 128);
    const int jmax = p.causal ? i : p.T - 1;  // last key this query attends to
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sc = 0.08838834764831845f * 1.4426950408889634f;  // 1/sqrt(128) * log2(e)

    // Warps whose 32 rows all lie beyond the sequence (T = 60: rows 64..127 of the 128-row MMA tile, i.e. half of the
    // softmax warps) only keep the P barrier's arrival count: their P rows stay undefined, which only feeds the unused
    // rows 64..127 of O, and they neither stage nor store anything.
    const bool warp_live = q0 + q * 32 < p.T && q * 32 < (TB == 64 ? 64 : 128);
    if (!warp_live) {
      for (int kc = 0; kc < nkc; ++kc) {
        if (kc > 0) ptx::mbar_wait(barO, (kc - 1) & 1);
        ptx::mbar_arrive(barP);
      }
    } else {
    for (int kc = 0; kc < nkc; ++kc) ptx::mbar_wait(&barS[kc], 0);  // all score chunks complete
    ptx::tcgen05_fence_after();
    float mx = -INFINITY, sum = 0.f;
    auto row_max = [&](const uint32_t (&v)[32], int j0) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j0 + j <= jmax && j0 + j < p.T) mx = fmaxf(mx, __uint_as_float(v[j]));
    };
    // P for 32 keys starting at key c0 of chunk kc: exp2, bf16 (hi, lo) split, K-major SW128 A-operand layout
    auto emit_p = [&](const uint32_t (&v)[32], int kc, int c0) {
#pragma unroll
      for (int j8 = 0; j8 < 32; j8 += 8) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = kc * TB + c0 + j8 + 2 * e;
          float p0 = 0.f, p1 = 0.f;
          if (row_ok && j <= jmax) p0 = exp2f((__uint_as_float(v[j8 + 2 * e]) - mx) * sc);
          if (row_ok && j + 1 <= jmax) p1 = exp2f((__uint_as_float(v[j8 + 2 * e + 1]) - mx) * sc);
          sum += p0 + p1;
          __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
          hw[e] = *reinterpret_cast<uint32_t*>(&hh);
          __nv_bfloat162 ll = __floats2bfloat162_rn(p0 - __uint_as_float(hw[e] << 16), p1 - __uint_as_float(hw[e] & 0xffff0000u));
          lw[e] = *reinterpret_cast<uint32_t*>(&ll);
        }
        // 16-byte chunk index XOR (row & 7)
        const int jj = c0 + j8;  // key offset inside the chunk
        const uint32_t off = (uint32_t)(jj >> 6) * C::P_TILE + (uint32_t)r * 128 +
                             ((((uint32_t)(jj & 63) >> 3) ^ ((uint32_t)r & 7)) << 4);
        *reinterpret_cast<uint4*>(sQ + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(sQ + (TB / 64) * C::P_TILE + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    };
    if constexpr (C::COMPACT) {
      // one key chunk of 64, two warps per lane quarter: this warp's 32 scores of the row are read from TMEM ONCE and
      // stay in registers for the maximum and the exponentials; the row maximum is exchanged with the sibling warp
      static_assert(KH == 32, "compact softmax: one 32-column TMEM load per warp");
      uint32_t s0[32];
      __syncwarp();
      ptx::tmem_ld_32x32b_x32(lane_addr + S_COL + (uint32_t)(half * KH), s0);
      ptx::tmem_ld_wait(s0);
      row_max(s0, half * KH);
      s_max[r * 2 + half] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      mx = fmaxf(mx, s_max[r * 2 + (half ^ 1)]);
      if (!row_ok) mx = 0.f;
      if (warp == 1 && lane == 0) REGEN_ATL(4);
      emit_p(s0, 0, half * KH);
      ptx::fence_proxy_async_smem();  // make the generic-proxy stores visible to the tensor core (async proxy)
      ptx::mbar_arrive(barP);
      if (warp == 1 && lane == 0) REGEN_ATL(5);
    } else {
    // pass 1: exact row maximum over the causal window (this warp's key half of every chunk, then exchange)
    for (int kc = 0; kc < nkc; ++kc) {
#pragma unroll 1
      for (int c0 = half * KH; c0 < half * KH + KH; c0 += 32) {
        uint32_t v[32];
        __syncwarp();
        ptx::tmem_ld_32x32b_x32(lane_addr + S_COL + (uint32_t)(kc * TB + c0), v);
        ptx::tmem_ld_wait();
        row_max(v, kc * TB + c0);
      }
    }
    if (NH == 2) {
      s_max[r * 2 + half] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      mx = fmaxf(mx, s_max[r * 2 + (half ^ 1)]);
    }
    if (!row_ok) mx = 0.f;
    if (warp == 1 && lane == 0) REGEN_ATL(4);
    for (int kc = 0; kc < nkc; ++kc) {
      if (kc > 0) ptx::mbar_wait(barO, (kc - 1) & 1);  // previous P chunk consumed by the tensor core
#pragma unroll 1
      for (int c0 = half * KH; c0 < half * KH + KH; c0 += 32) {
        uint32_t v[32];
        __syncwarp();
        ptx::tmem_ld_32x32b_x32(lane_addr + S_COL + (uint32_t)(kc * TB + c0), v);
        ptx::tmem_ld_wait();
        emit_p(v, kc, c0);
      }
      ptx::fence_proxy_async_smem();  // make the generic-proxy stores visible to the tensor core (async proxy)
      ptx::mbar_arrive(barP);
      if (warp == 1 && lane == 0 && kc == 0) REGEN_ATL(5);
    }
    }
    if (NH == 2) {
      s_sum[r * 2 + half] = sum;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      sum += s_sum[r * 2 + (half ^ 1)];
    }
    // epilogue: O / rowsum -> bf16 (hi, lo), staged as [32 rows x 64 d] tiles (128-byte rows, SWIZZLE_128B) in the
    // dead K / V operand buffers and written with TMA stores (3-D box {64 d, 1 sample, 32 frames}; frames >= T clipped)
    ptx::mbar_wait(barO, (nkc - 1) & 1);
    if (warp == 1 && lane == 0) REGEN_ATL(7);
    ptx::tcgen05_fence_after();
    const float inv = 1.f / sum;
    // staging: (DW / 64) hi tiles + as many lo tiles of 4 KB per warp, 64 KB per CTA: the dead K/V buffers, or in
    // compact mode the dead P (= Q) and K/V buffers, which are contiguous
    uint8_t* st = (C::COMPACT ? sQ : sK) + (warp - 1) * (DW / 64) * 8192;
    auto stage_o = [&](const uint32_t (&v)[32], int c0) {
      uint8_t* st_hi = st + (c0 >> 6) * 8192;
      stage_out32(v, inv, st_hi, st_hi + 4096, lane, c0, p.m8 != 0);
    };
    {  // the next 32 columns of O are in flight while the current ones are scaled, split and staged
      static_assert((DW / 32) % 2 == 0, "O columns are processed in (va, vb) pairs");
      uint32_t va[32], vb[32];
      const uint32_t o_addr = lane_addr + O_COL + (uint32_t)(half * DW);
      __syncwarp();
      ptx::tmem_ld_32x32b_x32(o_addr, va);
#pragma unroll 1
      for (int c0 = 0; c0 < DW; c0 += 64) {
        ptx::tmem_ld_wait(va);
        ptx::tmem_ld_32x32b_x32(o_addr + (uint32_t)(c0 + 32), vb);
        stage_o(va, c0);
        ptx::tmem_ld_wait(vb);
        if (c0 + 64 < DW) ptx::tmem_ld_32x32b_x32(o_addr + (uint32_t)(c0 + 64), va);
        stage_o(vb, c0 + 32);
      }
    }
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && q0 + q * 32 < p.T) {
#pragma unroll
      for (int tile = 0; tile < DW / 64; ++tile) {
        const int col = h * HD + half * DW + tile * 64;
        ptx::tma_store_3d(&tm_ohi, st + tile * 8192, col, b, q0 + q * 32, p.pol_store);
        // mixed8: the byte tile (residual bytes | hi bytes of these 64 columns) lands at byte column 2 col of the [.., 1024] rows
        ptx::tma_store_3d(&tm_olo, st + tile * 8192 + 4096, p.m8 ? 2 * col : col, b, q0 + q * 32, p.pol_store);
      }
      ptx::bulk_commit();
      // the staging tiles must have been read before the CTA exits; the kernel boundary orders the global writes
      if (p.dbg & 4) ptx::bulk_wait<0>(); else ptx::bulk_wait_read<0>();
    }
    }  // warp_live
  }

  if (warp == 1 && lane == 0) REGEN_ATL(8);
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) REGEN_ATL(9);
  ptx::steplog_end(p.steplog, p.steplog_slot, p.steplog_cta);
  if (warp == 0) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =====================================================================================================
// Multi-chunk compact kernel for 64 < T <= 256 ("MC"): the compact kernel's footprint generalised to several key chunks.
//
// One CTA per (sample b, head h, 64-query block).  Keys / values stream in chunks of 64 through TWO 32 KB operand buffers;
// the operand sequence K_0 .. K_{n-1}, V_0 .. V_{n-1} alternates between them (operand j -> buffer j & 1) and operand j is
// fetched as soon as the MMAs that read operand j - 2 have completed, so every load runs behind an MMA / softmax phase.
// The scores of all (<= 4) chunks stay resident in TMEM columns [0, 64 n) so the row maximum is exact before any
// exponential; P chunks are double buffered inside the dead Q buffer (valid rows 0..63 only: hi | lo | hi | lo, 8 KB each)
// and O (128 columns) overwrites the score columns of chunks 0 and 1 once BOTH have been turned into P.
// Footprint: 3 x 32 KB + barriers of shared memory, 256 TMEM columns, 288 threads -> two CTAs per SM (the 128-key-chunk
// kernel above runs one), and a 64-row query block wastes fewer MMA rows on the ragged last block (T = 150: 22 of 64
// rows instead of 22 of 128).  Every barrier is used exactly once per CTA (parity 0): no phase bookkeeping.
// =====================================================================================================
struct CfgMC {
  static constexpr int TILE = 64 * 128;          // 64 rows x 64 bf16
  static constexpr int OPERAND = 4 * TILE;       // hi d0-63 | hi d64-127 | lo d0-63 | lo d64-127  (32 KB)
  static constexpr int MAXC = 4;                 // key chunks of 64: T <= 256
  static constexpr int SW = 8;
  static constexpr int THREADS = 32 + 32 * SW;
  static constexpr int SMEM_BYTES = 3 * OPERAND + 4096 + 1024;
  static constexpr uint32_t TMEM_COLS = 256;
};

__global__ void __launch_bounds__(CfgMC::THREADS, 2)
attention_mc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                    const __grid_constant__ CUtensorMap tm_ohi, const __grid_constant__ CUtensorMap tm_olo, const Params p) {
  using C = CfgMC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                       // Q, later the two P buffers
  uint8_t* sB0 = smem + C::OPERAND;         // operand buffers (K / V chunks)
  uint8_t* sB1 = smem + 2 * C::OPERAND;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * C::OPERAND);
  uint64_t* barQ = bars + 0;
  uint64_t* barOp = bars + 1;               // [8]  operand j landed (K_0..K_{n-1}, V_0..V_{n-1})
  uint64_t* barS = bars + 9;                // [4]  S chunk complete (tcgen05.commit)
  uint64_t* barP = bars + 13;               // [4]  P chunk written by all softmax threads
  uint64_t* barPV = bars + 17;              // [4]  P.V chunk complete (tcgen05.commit)
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 21);
  float* s_max = reinterpret_cast<float*>(smem + 3 * C::OPERAND + 1024);   // [64 rows][2 key halves]
  float* s_sum = s_max + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblocks = (p.T + 63) / 64;
  const int qb = blockIdx.x % qblocks;
  const int bh = blockIdx.x / qblocks;
  const int h = bh & 3, b = bh >> 2;
  const int q0 = qb * 64;
  const int kv_len = p.causal ? min(p.T, q0 + 64) : p.T;
  const int n = (kv_len + 63) / 64;         // key chunks, 1..4

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tm_hi);
      ptx::prefetch_tmap(&tm_lo);
      ptx::mbar_init(barQ, 1);
      for (int i = 0; i < 2 * C::MAXC; ++i) ptx::mbar_init(&barOp[i], 1);
      for (int i = 0; i < C::MAXC; ++i) {
        ptx::mbar_init(&barS[i], 1);
        ptx::mbar_init(&barP[i], 32 * C::SW);
        ptx::mbar_init(&barPV[i], 1);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_base_smem, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  ptx::griddep_wait();
  ptx::griddep_launch();
  ptx::steplog_begin(p.steplog, p.steplog_slot);

  if (warp == 0) {
    {
      // ------------------------------------------------------------------ control warp: TMA + MMA issue (warp-uniform,
      // one elected lane per TMA / tcgen05 instruction, see attention_kernel)
      auto load_operand = [&](uint8_t* dst, uint64_t* bar, int col0, int t0) {
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(bar, C::OPERAND);
          ptx::tma_load_3d(dst, &tm_hi, bar, col0, b, t0, p.pol_load);
          ptx::tma_load_3d(dst + C::TILE, &tm_hi, bar, col0 + 64, b, t0, p.pol_load);
          ptx::tma_load_3d(dst + 2 * C::TILE, &tm_lo, bar, col0, b, t0, p.pol_load);
          ptx::tma_load_3d(dst + 3 * C::TILE, &tm_lo, bar, col0 + 64, b, t0, p.pol_load);
        }
      };
      // operand j of the sequence K_0..K_{n-1}, V_0..V_{n-1} -> buffer j & 1
      auto load_op = [&](int j) {
        uint8_t* dst = (j & 1) ? sB1 : sB0;
        if (j < n) load_operand(dst, &barOp[j], DM + h * HD, j * 64);
        else load_operand(dst, &barOp[j], 2 * DM + h * HD, (j - n) * 64);
      };
      load_operand(sQ, barQ, h * HD, q0);
      load_op(0);
      load_op(1);  // K_1, or V_0 when there is a single key chunk (2 n >= 2 always)

      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16_f32(128, 64);
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16_f32(128, HD) | (1u << 16);  // B (V) is MN-major
      const uint32_t aQ = ptx::smem_u32(sQ), aB0 = ptx::smem_u32(sB0), aB1 = ptx::smem_u32(sB1);

      // ---- scores: S_kc = Q . K_kc^T into TMEM columns [64 kc, 64 kc + 64)
      ptx::mbar_wait(barQ, 0);
      for (int kc = 0; kc < n; ++kc) {
        ptx::mbar_wait(&barOp[kc], 0);
        ptx::tcgen05_fence_after();
        const uint32_t accS = tmem_base + (uint32_t)(kc * 64);
        const uint32_t aK = (kc & 1) ? aB1 : aB0;
        {  // one elected lane issues the chunk's 24 MMAs back to back (descriptor bases built once, see attention_kernel)
          const uint64_t dq_hi0 = ptx::umma_desc_k_sw128(aQ), dq_hi1 = ptx::umma_desc_k_sw128(aQ + C::TILE);
          const uint64_t dq_lo0 = ptx::umma_desc_k_sw128(aQ + 2 * C::TILE), dq_lo1 = ptx::umma_desc_k_sw128(aQ + 3 * C::TILE);
          const uint64_t dk_hi0 = ptx::umma_desc_k_sw128(aK), dk_hi1 = ptx::umma_desc_k_sw128(aK + C::TILE);
          const uint64_t dk_lo0 = ptx::umma_desc_k_sw128(aK + 2 * C::TILE), dk_lo1 = ptx::umma_desc_k_sw128(aK + 3 * C::TILE);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint64_t adv = (uint64_t)(2 * (k & 3));
              const uint64_t q_hi = ((k >> 2) ? dq_hi1 : dq_hi0) + adv, q_lo = ((k >> 2) ? dq_lo1 : dq_lo0) + adv;
              const uint64_t k_hi = ((k >> 2) ? dk_hi1 : dk_hi0) + adv, k_lo = ((k >> 2) ? dk_lo1 : dk_lo0) + adv;
              ptx::mma_f16_ss(accS, q_lo, k_hi, idesc_s, k != 0);
              ptx::mma_f16_ss(accS, q_hi, k_lo, idesc_s, 1);
              ptx::mma_f16_ss(accS, q_hi, k_hi, idesc_s, 1);
            }
          }
        }
        if (ptx::elect_one()) ptx::tcgen05_commit(&barS[kc]);
        ptx::mbar_wait(&barS[kc], 0);              // K_kc consumed: its buffer takes operand kc + 2
        if (kc + 2 < 2 * n) load_op(kc + 2);
      }
      // ---- O += P_kc . V_kc  (O in TMEM columns [0, 128) = the dead scores of chunks 0 and 1)
      for (int kc = 0; kc < n; ++kc) {
        ptx::mbar_wait(&barP[kc], 0);
        if (kc == 0 && n > 1) ptx::mbar_wait(&barP[1], 0);   // S_1 has been read too before O overwrites it
        ptx::mbar_wait(&barOp[n + kc], 0);
        ptx::tcgen05_fence_after();
        const uint32_t accO = tmem_base;
        const uint32_t aV = ((n + kc) & 1) ? aB1 : aB0;
        const uint32_t aP = aQ + (uint32_t)(kc & 1) * 16384u;   // P buffer kc & 1: hi at +0, lo at +8 KB
        {
          const uint64_t dp_hi = ptx::umma_desc_k_sw128(aP), dp_lo = ptx::umma_desc_k_sw128(aP + 8192u);
          const uint64_t dv_hi = umma_desc_mn_sw128(aV, (uint32_t)C::TILE, 1024u);
          const uint64_t dv_lo = umma_desc_mn_sw128(aV + 2 * C::TILE, (uint32_t)C::TILE, 1024u);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // 16 keys per step: +32 bytes in P (K-major), +2048 bytes in V (MN-major)
              ptx::mma_f16_ss(accO, dp_lo + (uint64_t)(2 * k), dv_hi + (uint64_t)(128 * k), idesc_o, (kc | k) != 0);
              ptx::mma_f16_ss(accO, dp_hi + (uint64_t)(2 * k), dv_lo + (uint64_t)(128 * k), idesc_o, 1);
              ptx::mma_f16_ss(accO, dp_hi + (uint64_t)(2 * k), dv_hi + (uint64_t)(128 * k), idesc_o, 1);
            }
          }
        }
        if (ptx::elect_one()) ptx::tcgen05_commit(&barPV[kc]);
        if (n + kc + 2 < 2 * n) {                   // V_kc consumed: its buffer takes V_{kc + 2}
          ptx::mbar_wait(&barPV[kc], 0);
          load_op(n + kc + 2);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (warps 1..8)
    const int q = warp & 3, half = (warp - 1) >> 2;   // TMEM lane quarter; key half (softmax) / head-dim half (epilogue)
    const int r = q * 32 + lane;                      // row of the 128-row MMA tile == TMEM lane; rows >= 64 are padding
    const int i = q0 + r;                             // query frame
    const bool row_ok = r < 64 && i < p.T;
    const int jmax = p.causal ? i : p.T - 1;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sc = 0.08838834764831845f * 1.4426950408889634f;  // 1/sqrt(128) * log2(e)
    const bool warp_live = q * 32 < 64 && q0 + q * 32 < p.T;
    if (!warp_live) {
      for (int kc = 0; kc < n; ++kc) ptx::mbar_arrive(&barP[kc]);   // keep the arrival counts only
    } else {
      for (int kc = 0; kc < n; ++kc) ptx::mbar_wait(&barS[kc], 0);
      ptx::tcgen05_fence_after();
      float mx = -INFINITY, sum = 0.f;
      // pass 1: exact row maximum over this warp's 32 keys of every chunk, then exchange with the sibling warp
      for (int kc = 0; kc < n; ++kc) {
        uint32_t v[32];
        __syncwarp();
        ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(kc * 64 + half * 32), v);
        ptx::tmem_ld_wait(v);
        const int j0 = kc * 64 + half * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j0 + j <= jmax && j0 + j < p.T) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      s_max[r * 2 + half] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      mx = fmaxf(mx, s_max[r * 2 + (half ^ 1)]);
      if (!row_ok) mx = 0.f;
      // pass 2: P chunks (exp2, bf16 hi / lo split, K-major SW128 A-operand layout) into the P buffer kc & 1
      for (int kc = 0; kc < n; ++kc) {
        if (kc >= 2) ptx::mbar_wait(&barPV[kc - 2], 0);   // the tensor core is done with this P buffer
        uint32_t v[32];
        __syncwarp();
        ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)(kc * 64 + half * 32), v);
        ptx::tmem_ld_wait(v);
        uint8_t* pb = sQ + (kc & 1) * 16384;
#pragma unroll
        for (int j8 = 0; j8 < 32; j8 += 8) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = kc * 64 + half * 32 + j8 + 2 * e;
            float p0 = 0.f, p1 = 0.f;
            if (row_ok && j <= jmax) p0 = exp2f((__uint_as_float(v[j8 + 2 * e]) - mx) * sc);
            if (row_ok && j + 1 <= jmax) p1 = exp2f((__uint_as_float(v[j8 + 2 * e + 1]) - mx) * sc);
            sum += p0 + p1;
            __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
            hw[e] = *reinterpret_cast<uint32_t*>(&hh);
            __nv_bfloat162 ll = __floats2bfloat162_rn(p0 - __uint_as_float(hw[e] << 16), p1 - __uint_as_float(hw[e] & 0xffff0000u));
            lw[e] = *reinterpret_cast<uint32_t*>(&ll);
          }
          const int jj = half * 32 + j8;  // key offset inside the chunk
          const uint32_t off = (uint32_t)r * 128 + ((((uint32_t)jj >> 3) ^ ((uint32_t)r & 7)) << 4);
          *reinterpret_cast<uint4*>(pb + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(pb + 8192 + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        ptx::tcgen05_fence_before();    // this warp's tcgen05.ld of S_kc is ordered before the MMAs that overwrite it
        ptx::fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
        ptx::mbar_arrive(&barP[kc]);
      }
      s_sum[r * 2 + half] = sum;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      sum += s_sum[r * 2 + (half ^ 1)];
      // epilogue: O / rowsum -> bf16 (hi, lo); this warp owns head-dim columns [64 half, 64 half + 64); staged as one
      // 32 x 64 hi tile + one lo tile (4 KB each) in the dead Q / operand buffers, written with TMA stores
      ptx::mbar_wait(&barPV[n - 1], 0);
      ptx::tcgen05_fence_after();
      const float inv = 1.f / sum;
      uint8_t* st_hi = sQ + (warp - 1) * 8192;
      uint8_t* st_lo = st_hi + 4096;
      const uint32_t o_addr = lane_addr + (uint32_t)(half * 64);
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        __syncwarp();
        ptx::tmem_ld_32x32b_x32(o_addr + (uint32_t)c0, v);
        ptx::tmem_ld_wait(v);
        stage_out32(v, inv, st_hi, st_lo, lane, c0, p.m8 != 0);
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_3d(&tm_ohi, st_hi, h * HD + half * 64, b, q0 + q * 32, p.pol_store);
        ptx::tma_store_3d(&tm_olo, st_lo, (p.m8 ? 2 : 1) * (h * HD + half * 64), b, q0 + q * 32, p.pol_store);
        ptx::bulk_commit();
        ptx::bulk_wait<0>();
      }
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::steplog_end(p.steplog, p.steplog_slot, p.steplog_cta);
  if (warp == 0) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// tm_hi / tm_lo must be the 3-D q|k|v maps with 64-FRAME boxes (like the compact kernel's)
inline cudaError_t launch_mc(const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, const CUtensorMap& tm_ohi,
                             const CUtensorMap& tm_olo, const Params& p, cudaStream_t s) {
  using C = CfgMC;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (p.T > 64 * C::MAXC) return cudaErrorInvalidValue;
  const int qblocks = (p.T + 63) / 64;
  return launch_pdl(attention_mc_kernel, dim3(p.Beff * 4 * qblocks), dim3(C::THREADS), C::SMEM_BYTES, s, tm_hi, tm_lo,
                    tm_ohi, tm_olo, p);
}

template <int TB>
inline cudaError_t launch(const CUtensorMap& tm_hi, const CUtensorMap& tm_lo, const CUtensorMap& tm_ohi,
                          const CUtensorMap& tm_olo, const Params& p, cudaStream_t s) {
  using C = Cfg<TB>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (p.T > (TB == 64 ? 64 : 256)) return cudaErrorInvalidValue;  // at most two key chunks of 128
  const int qblocks = (p.T + 127) / 128;
  return launch_pdl(attention_kernel<TB>, dim3(p.Beff * 4 * qblocks), dim3(C::THREADS), C::SMEM_BYTES, s, tm_hi, tm_lo,
                    tm_ohi, tm_olo, p);
}

// =====================================================================================================
// Long sequences (more than 256 tokens per sample): streaming attention on the CUDA cores, exact fp32.
//
// The tcgen05 kernels above keep the scores of a whole key row resident in tensor memory, which bounds them at 256 keys;
// the reference's positional table allows 5000 frames (model/cmdm.py:265-281), so longer sequences must not be an
// error.  No dataset of the reference is longer than 196 frames, so this path is about completeness, not speed: one warp
// per query row, keys / values streamed through shared memory in tiles of 32 (hi + lo recombined to fp32 while staging),
// online softmax (running maximum and sum, fp32), output re-split into the bf16 (hi, lo) pair the output projection reads.
// Reference semantics as above: nn.MultiheadAttention, additive causal mask (model/cmdm.py:168-171) or none.
// =====================================================================================================
constexpr int kLongWarps = 8, kLongTile = 32;

__global__ void __launch_bounds__(kLongWarps * 32) attention_long_kernel(const __nv_bfloat16* __restrict__ qkv_hi,
                                                                         const __nv_bfloat16* __restrict__ qkv_lo,
                                                                         __nv_bfloat16* __restrict__ out_hi,
                                                                         __nv_bfloat16* __restrict__ out_lo, int T, int Beff,
                                                                         int causal) {
  __shared__ float sK[kLongTile][HD + 1];
  __shared__ float sV[kLongTile][HD];
  __shared__ float sQ[kLongWarps][HD];
  ptx::griddep_wait();
  ptx::griddep_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y >> 2, h = blockIdx.y & 3;
  const int i0 = blockIdx.x * kLongWarps;                  // first query frame of this block
  const int i = i0 + warp;                                 // this warp's query frame
  const bool live = i < T;
  const int last_q = min(i0 + kLongWarps, T) - 1;
  const int kv_len = causal ? last_q + 1 : T;              // keys [0, kv_len) are needed by some row of the block
  auto ld = [&](int t, int col) {                          // fp32 value of q | k | v element (frame t, column col)
    const size_t off = ((size_t)t * Beff + b) * (3 * DM) + col;
    return __bfloat162float(qkv_hi[off]) + __bfloat162float(qkv_lo[off]);
  };
  if (live)
    for (int d = lane; d < HD; d += 32) sQ[warp][d] = ld(i, h * HD + d);
  const float sc = 0.08838834764831845f;                   // 1/sqrt(128)
  float m = -INFINITY, l = 0.f;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};                     // output dims lane, lane + 32, lane + 64, lane + 96
  for (int j0 = 0; j0 < kv_len; j0 += kLongTile) {
    __syncthreads();                                       // previous tile fully consumed (and sQ visible)
    for (int e = threadIdx.x; e < kLongTile * HD; e += kLongWarps * 32) {
      const int r = e / HD, d = e - r * HD;
      const bool ok = j0 + r < kv_len;
      sK[r][d] = ok ? ld(j0 + r, DM + h * HD + d) : 0.f;
      sV[r][d] = ok ? ld(j0 + r, 2 * DM + h * HD + d) : 0.f;
    }
    __syncthreads();
    if (!live) continue;
    const int j = j0 + lane;                               // lane = key of the tile
    const bool ok = j < T && (!causal || j <= i);
    float s = -INFINITY;
    if (ok) {
      float a = 0.f;
#pragma unroll 8
      for (int d = 0; d < HD; ++d) a = fmaf(sQ[warp][d], sK[lane][d], a);
      s = a * sc;
    }
    const float mt = warp_max(s);
    if (mt == -INFINITY) continue;                         // every key of this tile is masked for this row
    const float m_new = fmaxf(m, mt);
    const float corr = expf(m - m_new);                    // 0 at the first tile (m = -inf)
    const float pj = ok ? expf(s - m_new) : 0.f;
    l = l * corr + warp_sum(pj);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] *= corr;
    for (int r = 0; r < kLongTile; ++r) {
      const float pr = __shfl_sync(0xffffffffu, pj, r);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = fmaf(pr, sV[r][lane + 32 * k], acc[k]);
    }
    m = m_new;
  }
  if (live) {
    const float inv = 1.f / l;
    const size_t row = ((size_t)i * Beff + b) * DM + h * HD;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat16 hi, lo;
      split_bf16(acc[k] * inv, hi, lo);
      out_hi[row + lane + 32 * k] = hi;
      out_lo[row + lane + 32 * k] = lo;
    }
  }
}

inline cudaError_t launch_long(const __nv_bfloat16* qkv_hi, const __nv_bfloat16* qkv_lo, const Params& p, cudaStream_t s) {
  if (p.Beff * 4 > 65535) return cudaErrorInvalidValue;
  return launch_pdl(attention_long_kernel, dim3((unsigned)((p.T + kLongWarps - 1) / kLongWarps), (unsigned)(p.Beff * 4)),
                    dim3(kLongWarps * 32), 0, s, qkv_hi, qkv_lo, p.out_hi, p.out_lo, p.T, p.Beff, p.causal);
}

}  // namespace attn
}  // namespace regen
