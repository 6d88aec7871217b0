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
//
// R16 = true (default): the residual stream lives in HBM only as its bf16 (hi, lo) pair.  The residual tile is loaded
// as two 32 x 64 bf16 boxes (128-byte rows) and rebuilt as hi + lo, and the final pass stores hi / lo only: the final
// pass is bound by the TMA row rate (one box row per ~2.6 cycles), so dropping the fp32 copy halves its 128 row requests
// per 64 columns and the bytes written.  The residual then carries 16 significand bits (forward error ~3e-5 instead of
// ~1.5e-5, tolerance 1e-3).  R16 = false (REGEN_DEBUG_F32_RESIDUAL=1) keeps the fp32 copy of h.
//
// M8 = true (precision 'mixed8': attention out_proj + norm1 + norm2 and linear2 + norm3): the product runs as ONE fp16 MMA plus two e4m3 correction MMAs at twice
// the rate -- 2 bf16-MMA equivalents per product instead of 3 -- at the same 4 operand bytes per element:
//   D  = (A_lo * 2^9) . (W_hi * 2^6)^T + (A_hi * 2^-2) . (W_lo * 2^17)^T   kind::f8f6f4, all four operands e4m3, K = 32 per MMA
//   D  = A16 . W16^T + D * 2^-15                                         first kind::f16 MMA (scale-input-d = 15)
//   D += A16 . W16^T                                                     remaining kind::f16 MMAs
// with A16 = fp16(a), A_lo = a - A16, A_hi = A16 (same for W).  The correction terms are 2^-11 of the result, so
// 4 significand bits suffice for them (measured on B200: 4.0e-5 against 8.6e-6 for fp16 x3, tools/mixed8_probe.cu).
// The power-of-two operand scales (product 2^15 in both terms, undone by scale-input-d) place |a| <= 1792 and |w| <= 7 inside
// e4m3's range; beyond that the fp8 copies saturate and the element falls back to plain fp16 accuracy.
// Operands: tm_a_hi = A16 [M, K] (16-bit), tm_a_lo = A8 [M, 2K] bytes, tm_w_hi = W16 [512, K], tm_w_lo = W8 [512, 2K] bytes.
// The byte rows interleave the two correction operands per group of 64 K elements: A8 = [A_lo8 (64) | A_hi8 (64)] ...,
// W8 = [W_hi8 (64) | W_lo8 (64)] ..., so that one 128-byte row segment of A8 against the same segment of W8 is the sum of
// both correction products of those 64 elements (and the producing epilogues emit ONE 32 x 128-byte box per 64 columns);
// the 8-bit boxes are 128 rows x 128 bytes.  The ring is then 4 stages of
// 48 KB (A tile | two W half tiles of 16 KB), every stage feeds 8 MMAs (1 k cycles): first the 2 K / 128 correction stages,
// then the K / 64 main-term stages.
#pragma once
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "ptx.cuh"

namespace regen {
namespace gemmln {

constexpr int BM = 128, BK = 64, UMMA_K = 16, ND = 512;
constexpr int STAGE_BYTES = 6 * 16384;  // A_hi | A_lo | W_hi[0] | W_hi[1] | W_lo[0] | W_lo[1]
constexpr int STAGES = 2;
constexpr int RING_BYTES = STAGES * STAGE_BYTES;   // operand ring; re-used as epilogue staging
constexpr int M8_STAGES = 4, M8_STAGE_BYTES = 3 * 16384;  // mixed8 ring: A | W[0] | W[1], same 192 KB
static_assert(M8_STAGES * M8_STAGE_BYTES == RING_BYTES, "both ring shapes fill the same bytes");
constexpr int SC = 32;                  // columns per epilogue sub-chunk (one tcgen05.ld/st x32, one TMA box)
constexpr int SLOT = 32 * SC * 4;       // 4 KB: 32 rows x 32 fp32, 128-byte rows (SWIZZLE_128B)
// Epilogue geometry for EW epilogue warps per CTA (EW / 4 warps share a TMEM lane quarter and split the 512 columns).
// The operand ring (192 KB) is re-used as staging, WARP_BYTES per warp: residual ring | c ring during passes 1 / 2, and
// the output staging of the final pass aliases the same bytes (both rings are drained by then):
//   EW =  8: 24 KB per warp: res 3 x 4 KB + c 3 x 4 KB (res 6 x 4 KB without CHAIN);
//            output: fp32 ring 2 x 4 KB | bf16 hi 2 x 4 KB | bf16 lo 2 x 4 KB  (bf16 tiles are 32 rows x 64 columns)
//   EW = 16: 12 KB per warp: res 2 x 4 KB, c 1 x 4 KB; output: fp32 1 x 4 KB | hi 4 KB | lo 4 KB  (A/B variant only)
template <int EW, bool CHAIN = true>
struct Epi {
  static constexpr int PARTS = EW / 4;
  static constexpr int WCOLS = ND / PARTS;      // columns per warp
  static constexpr int NSC = WCOLS / SC;        // sub-chunks per warp
  // without CHAIN there is no c ring: its slots deepen the residual ring (more TMA loads in flight per warp)
  static constexpr int RING_R = CHAIN ? (EW == 16 ? 2 : 3) : (EW == 16 ? 3 : 6);
  static constexpr int RING_C = EW == 16 ? 1 : 3;
  static constexpr int NOB = EW == 16 ? 1 : 2;  // output staging depth (fp32 slots, and bf16 hi / lo tile pairs)
  static constexpr int WARP_BYTES = RING_BYTES / EW;
  static constexpr int THREADS = 64 + 32 * EW;
  static_assert((RING_R + (CHAIN ? RING_C : 0)) * SLOT <= WARP_BYTES && 3 * NOB * SLOT <= WARP_BYTES,
                "epilogue staging budget");
  static_assert(RING_R <= 8 && (!CHAIN || RING_R <= 4), "ring barriers: residual slots 0.., c slots 4..");
  static_assert(NSC % 2 == 0, "sub-chunks are processed in pairs (register double buffer, 64-column bf16 tiles)");
};
constexpr int PARAM_BYTES = 5 * ND * 4; // bias, g1, b1, g2, b2
constexpr int STATS_BYTES = 2 * 128 * 4 * 8;   // [2 exchanges][128 rows][<= 4 column parts] float2
constexpr int BAR_BYTES = 2048;                // 16 pipeline barrier slots + 16 warps x 8 ring barriers
constexpr int SMEM_BYTES = RING_BYTES + PARAM_BYTES + STATS_BYTES + BAR_BYTES + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct Params {
  int M, K, Beff;
  const float *bias, *g1, *b1, *g2, *b2;  // [512] each (g2/b2 unused without CHAIN)
  float ln_eps;
  int exit_wait_full;            // 1: wait for the bulk stores' global writes before exit (default; REGEN_DEBUG_EXIT_WAIT_READ=1 clears it, A/B: no measurable difference)
  int nob16;                     // R16 only: output staging depth of the final pass (2 or 3 hi / lo tile pairs per warp)
  int store_f32;                 // LN = false only: 1 = also store the fp32 copy of h through tm_c (0 with R16 consumers)
  int prefetch_res;              // 1: L2-prefetch the residual tile during the main loop (REGEN_DEBUG_NO_RES_PREFETCH=1 -> 0)
  unsigned long long* timeline;  // bring-up instrumentation (null in production), see tools/ln_timeline.py
  unsigned long long* steplog;   // whole-step timeline (ptx::steplog_begin / steplog_end), null in production
  int steplog_slot, steplog_cta;  // steplog_cta: word offset of the per-CTA exit-time table (0 = off)
  // L2 eviction-priority hints (ptx::kL2Evict*, 0 = none): A tiles (read once: attention output / FFN activations), W tiles
  // (re-read by every row block), bf16 (hi, lo) outputs (operands of the next kernel)
  unsigned long long pol_a, pol_w, pol_store;
};

#define REGEN_LTL(k)                                                                                   \
  do {                                                                                                 \
    if (p.timeline && blockIdx.x == 0) p.timeline[(CHAIN ? 0 : 20) + (k)] = (unsigned long long)clock64(); \
  } while (0)

// byte offset of 16-byte chunk c of row r in a staged tile: fp32 32 x 32 (128-byte rows, SWIZZLE_128B) and
// bf16 32 x 32 (64-byte rows, SWIZZLE_64B)
__device__ __forceinline__ int off_f32(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
__device__ __forceinline__ int off_bf16(int r, int c) { return r * 64 + ((c ^ ((r >> 1) & 3)) << 4); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// LN = false (input projection): h <- A.W^T + b + residual, no LayerNorm, ONE sweep; tm_res is then only the residual
// LOAD map (the loop-invariant conditioning bias) and tm_c the fp32 STORE map of h.
// tm_res: fp32 [M, 512] residual stream h (box 32 x 32, SWIZZLE_128B) -- used for the residual LOAD and the h STORE
// tm_c  : fp32 [Beff + 32, 512] cyclic per-sample constant (row r = c[r % Beff]); only read with CHAIN
// tm_ohi / tm_olo: bf16 [M, 512] split of h (store, box 32 rows x 64 columns, SWIZZLE_128B)
template <bool SPLIT, bool CHAIN, int EW, bool LN = true, bool R16 = false, bool M8 = false, bool H8 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Epi<EW, CHAIN>::THREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
               const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_c,
               const __grid_constant__ CUtensorMap tm_ohi, const __grid_constant__ CUtensorMap tm_olo, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array (NOT an integer round trip): the compiler keeps the
  // shared address space and emits LDS / STS instead of generic LD / ST for every staging access below
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  static_assert(!M8 || (SPLIT && LN && R16), "mixed8 operands: built for the (hi, lo)-residual LayerNorm variants");
  // H8 (precision 'mixed8h'): the residual stream h itself lives in HBM as the mixed8 operand pack -- tm_ohi = fp16 [M, 512]
  // (box 32 x 64), tm_olo = bytes [M, 1024] (box 32 rows x 128 B: per 64 columns 64 bytes e4m3((h - fp16(h)) * 2^9) | 64 bytes
  // e4m3(fp16(h) / 4)), loaded as the residual (fp16 + residual byte * 2^-9: ~15 significand bits) and stored by the final
  // pass, so that the QKV / FFN1 / output GEMMs read h with the mixed8 main loop too.  Same box sizes and staging as R16.
  static_assert(!H8 || (EW == 8 && (R16 || !LN)), "the mixed8 residual-stream format shares the R16 staging layout");
  constexpr int OLO = H8 ? 2 : 1;   // column coordinate scale of tm_olo (bytes: two per column)
  constexpr int NST = M8 ? M8_STAGES : STAGES;              // ring stages
  constexpr int STB = M8 ? M8_STAGE_BYTES : STAGE_BYTES;    // bytes per stage
  float* s_par = reinterpret_cast<float*>(smem + RING_BYTES);             // bias | g1 | b1 | g2 | b2
  float2* s_stats = reinterpret_cast<float2*>(smem + RING_BYTES + PARAM_BYTES);  // [2 exchanges][128 rows][2 halves]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING_BYTES + PARAM_BYTES + STATS_BYTES);
  uint64_t* full_bar = bars;          // [NST <= 4]
  uint64_t* empty_bar = bars + 4;     // [NST <= 4]
  uint64_t* tmem_full_bar = bars + 8;
  uint64_t* tmem_empty_bar = bars + 9;   // leader's copy: 16 arrivals (epilogue warps of both CTAs)
  uint64_t* epi_done_bar = bars + 10;    // local: 8 arrivals, operand ring free again for the producer
  using E = Epi<EW, CHAIN>;
  // R16: residual ring of RING_P pair slots (hi box | lo box, 8 KB, two sub-chunks each), c ring of RC fp32 slots
  static_assert(!R16 || (EW == 8 && LN), "the bf16 (hi, lo) residual variant is built for 8 epilogue warps");
  constexpr int RING_P = CHAIN ? 2 : 3;
  constexpr int RC = R16 ? 2 : E::RING_C;
  constexpr int NP = E::NSC / 2;         // sub-chunk pairs per warp
  static_assert(!R16 || (RING_P * 2 + (CHAIN ? RC : 0)) * SLOT <= E::WARP_BYTES, "R16 staging budget");
  uint64_t* ring_bar = bars + 16;        // [EW warps][8]: res slots 0..3, c slots 4..7
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 16 + 8 * EW);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_kb = p.K / BK;
  const int num_tiles = (p.M + 2 * BM - 1) / (2 * BM);
  if (threadIdx.x == 0) REGEN_LTL(0);

  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tm_a_hi);
      ptx::prefetch_tmap(&tm_w_hi);
      ptx::prefetch_tmap(&tm_res);
      for (int s = 0; s < NST; ++s) {
        ptx::mbar_init(&full_bar[s], 1);
        ptx::mbar_init(&empty_bar[s], 1);
      }
      ptx::mbar_init(tmem_full_bar, 1);
      ptx::mbar_init(tmem_empty_bar, 2 * EW);
      ptx::mbar_init(epi_done_bar, EW);
    }
    for (int i = lane; i < 8 * EW; i += 32) ptx::mbar_init(&ring_bar[i], 1);
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
  ptx::griddep_wait();    // PDL: the setup above overlapped the previous kernel's tail
  ptx::griddep_launch();
  ptx::steplog_begin(p.steplog, p.steplog_slot);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int m0 = tile * (2 * BM) + (int)rank * BM;
        if (it > 0) ptx::mbar_wait(epi_done_bar, (it - 1) & 1);  // the epilogue staged in the operand ring
        // pipeline iterations of one tile: k-blocks of 64, or (mixed8) 2 K / 128 correction stages + K / 64 main stages
        const int n1 = M8 ? 2 * (p.K / 128) : 0;
        const int nit = M8 ? n1 + num_kb : num_kb;
        for (int kb = 0; kb < nit; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * STB;
          if constexpr (M8) {
            if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * STB);
            // correction stage i: bytes [128 i, 128 i + 128) of the byte rows = the residual-times-hi and hi-times-residual
            // operand bytes of 64 K elements (see the header comment: the two correction products are one dot product over 2 K)
            const bool corr = kb < n1;
            const CUtensorMap* ma = corr ? &tm_a_lo : &tm_a_hi;
            const CUtensorMap* mw = corr ? &tm_w_lo : &tm_w_hi;
            const int c0 = corr ? kb * 128 : (kb - n1) * BK;
            ptx::tma_load_2d_2sm(st, ma, &full_bar[stage], c0, m0, p.pol_a);
            ptx::tma_load_2d_2sm(st + 16384, mw, &full_bar[stage], c0, (int)rank * 128, p.pol_w);
            ptx::tma_load_2d_2sm(st + 2 * 16384, mw, &full_bar[stage], c0, 256 + (int)rank * 128, p.pol_w);
          } else {
          if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * (SPLIT ? STAGE_BYTES : STAGE_BYTES / 2));
          ptx::tma_load_2d_2sm(st, &tm_a_hi, &full_bar[stage], kb * BK, m0, p.pol_a);
          ptx::tma_load_2d_2sm(st + 2 * 16384, &tm_w_hi, &full_bar[stage], kb * BK, (int)rank * 128, p.pol_w);
          ptx::tma_load_2d_2sm(st + 3 * 16384, &tm_w_hi, &full_bar[stage], kb * BK, 256 + (int)rank * 128, p.pol_w);
          if (SPLIT) {
            ptx::tma_load_2d_2sm(st + 16384, &tm_a_lo, &full_bar[stage], kb * BK, m0, p.pol_a);
            ptx::tma_load_2d_2sm(st + 4 * 16384, &tm_w_lo, &full_bar[stage], kb * BK, (int)rank * 128, p.pol_w);
            ptx::tma_load_2d_2sm(st + 5 * 16384, &tm_w_lo, &full_bar[stage], kb * BK, 256 + (int)rank * 128, p.pol_w);
          }
          }
          // The epilogue's first pass reads this CTA's 128 x 512 fp32 residual tile in one burst; by then h has been
          // evicted from L2 by the operand stream (ncu: the whole tile came from DRAM, pass 1 was HBM-bound).  Pull it
          // into L2 during the main loop, which leaves DRAM bandwidth unused: 64 boxes of 32 x 32 spread over the k-blocks.
          if (p.prefetch_res) {
            const int per_kb = (64 + nit - 1) / nit;
            for (int i = kb * per_kb; i < (kb + 1) * per_kb && i < 64; ++i) {
              if constexpr (R16) {  // 32 hi + 32 lo boxes of 32 rows x 64 columns
                const int k = i >> 1;
                ptx::tma_prefetch_l2_2d((i & 1) ? &tm_olo : &tm_ohi, ((i & 1) ? OLO : 1) * (k & 7) * 2 * SC, m0 + (k >> 3) * 32);
              } else {
                ptx::tma_prefetch_l2_2d(&tm_res, (i & 15) * SC, m0 + (i >> 4) * 32);
              }
            }
          }
          if (++stage == NST) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    // The whole warp runs the loop (warp-uniform control flow and operands); one elected lane issues each tcgen05
    // instruction (ptx::elect_one: a single-lane loop costs more cycles per MMA in issue than the MMA takes to execute).
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(2 * BM, 256);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        ptx::mbar_wait(tmem_empty_bar, (it & 1) ^ 1);  // both CTAs' epilogues are done with the accumulator
        ptx::tcgen05_fence_after();
        if constexpr (M8) {
          // Three straight-line loops (no per-MMA branches in the issuing thread): correction stages (e4m3 x e4m3), the
          // first main-term stage (its two k = 0 MMAs fold the scaled corrections in: D = A.B + D * 2^-15), the remaining
          // main-term stages.  e4m3 x e4m3 and fp16 x fp16 share the descriptor bits (formats 0); 32 bytes per k-step.
          constexpr uint32_t idesc0 = ptx::umma_idesc_fmt0_f32(2 * BM, 256);
          const int n1 = 2 * (p.K / 128);
          int kb_log = 0;
          auto stage_begin = [&]() -> uint32_t {
            ptx::mbar_wait(&full_bar[stage], phase);
            if (kb_log == 0 && it == 0 && lane == 0) REGEN_LTL(1);
            if (!CHAIN && p.timeline && blockIdx.x == 0 && it == 0 && kb_log < 40 && lane == 0)
              p.timeline[48 + kb_log] = (unsigned long long)clock64();
            ++kb_log;
            ptx::tcgen05_fence_after();
            return ptx::smem_u32(smem + stage * STB);
          };
          auto stage_end = [&]() {
            if (ptx::elect_one()) ptx::tcgen05_commit_2sm(&empty_bar[stage]);
            if (++stage == NST) {
              stage = 0;
              phase ^= 1;
            }
          };
          for (int kb = 0; kb < n1; ++kb) {
            const uint32_t st = stage_begin();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a = ptx::umma_desc_k_sw128(st + k * 32);
              if (ptx::elect_one()) ptx::mma_f8_ss_2sm(tmem_base, a, ptx::umma_desc_k_sw128(st + 16384 + k * 32), idesc0, (kb | k) != 0);
              if (ptx::elect_one()) ptx::mma_f8_ss_2sm(tmem_base + 256, a, ptx::umma_desc_k_sw128(st + 2 * 16384 + k * 32), idesc0, (kb | k) != 0);
            }
            stage_end();
          }
          {
            const uint32_t st = stage_begin();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a = ptx::umma_desc_k_sw128(st + k * 32);
              const uint64_t w0 = ptx::umma_desc_k_sw128(st + 16384 + k * 32);
              const uint64_t w1 = ptx::umma_desc_k_sw128(st + 2 * 16384 + k * 32);
              if (k == 0) {
                if (ptx::elect_one()) ptx::mma_f16_ss_2sm_scale15(tmem_base, a, w0, idesc0);
                if (ptx::elect_one()) ptx::mma_f16_ss_2sm_scale15(tmem_base + 256, a, w1, idesc0);
              } else {
                if (ptx::elect_one()) ptx::mma_f16_ss_2sm(tmem_base, a, w0, idesc0, 1);
                if (ptx::elect_one()) ptx::mma_f16_ss_2sm(tmem_base + 256, a, w1, idesc0, 1);
              }
            }
            stage_end();
          }
          for (int kb = 1; kb < num_kb; ++kb) {
            const uint32_t st = stage_begin();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a = ptx::umma_desc_k_sw128(st + k * 32);
              if (ptx::elect_one()) ptx::mma_f16_ss_2sm(tmem_base, a, ptx::umma_desc_k_sw128(st + 16384 + k * 32), idesc0, 1);
              if (ptx::elect_one()) ptx::mma_f16_ss_2sm(tmem_base + 256, a, ptx::umma_desc_k_sw128(st + 2 * 16384 + k * 32), idesc0, 1);
            }
            stage_end();
          }
        } else {
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          if (kb == 0 && it == 0 && lane == 0) REGEN_LTL(1);
          // bring-up: arrival time of every stage of the first tile (linear2 instance: slots 48..)
          if (!CHAIN && LN && p.timeline && blockIdx.x == 0 && it == 0 && kb < 40 && lane == 0) p.timeline[48 + kb] = (unsigned long long)clock64();
          ptx::tcgen05_fence_after();
          const uint32_t st = ptx::smem_u32(smem + stage * STB);
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
                if (ptx::elect_one()) ptx::mma_f16_ss_2sm(acc, a_lo, w_hi, idesc, (kb | k) != 0);
                if (ptx::elect_one()) ptx::mma_f16_ss_2sm(acc, a_hi, w_lo, idesc, 1);
                if (ptx::elect_one()) ptx::mma_f16_ss_2sm(acc, a_hi, w_hi, idesc, 1);
              } else {
                if (ptx::elect_one()) ptx::mma_f16_ss_2sm(acc, a_hi, w_hi, idesc, (kb | k) != 0);
              }
            }
          }
          if (ptx::elect_one()) ptx::tcgen05_commit_2sm(&empty_bar[stage]);
          if (++stage == NST) {
            stage = 0;
            phase ^= 1;
          }
        }
        }
        if (ptx::elect_one()) ptx::tcgen05_commit_2sm(tmem_full_bar);
        if (it == 0 && lane == 0) REGEN_LTL(2);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: warps 2..9, one thread = one row
    const int ew = warp - 2;
    const int q = warp & 3, hf = ew >> 2;  // hf: which WCOLS-wide column part of the 512 this warp owns
    const int r_local = q * 32 + lane;                 // row inside this CTA's 128 rows == TMEM lane
    // LayerNorm / bias vectors -> shared memory while the main loop runs (global loads are L2 round trips here)
    for (int i = threadIdx.x - 64; i < 5 * ND / 4; i += 32 * EW) {
      const int which = i / (ND / 4), j = i % (ND / 4);
      const float* src = which == 0 ? p.bias : which == 1 ? p.g1 : which == 2 ? p.b1 : which == 3 ? p.g2 : p.b2;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src) v = __ldg(reinterpret_cast<const float4*>(src) + j);
      reinterpret_cast<float4*>(s_par)[i] = v;
    }
    named_bar_sync(5, 32 * EW);
    uint8_t* my = smem + ew * E::WARP_BYTES;
    uint8_t* res_ring = my;
    uint8_t* c_ring = my + (R16 ? RING_P * 2 : E::RING_R) * SLOT;
    uint8_t* out_buf = my;                             // 3 x NOB x 4 KB, used after both rings are drained
    uint64_t* rbar = ring_bar + ew * 8;                // [0..2] residual slots, [4..6] c slots
    const float* s_bias = s_par + hf * E::WCOLS;
    const float* s_g1 = s_par + ND + hf * E::WCOLS;
    const float* s_b1 = s_par + 2 * ND + hf * E::WCOLS;
    const float* s_g2 = s_par + 3 * ND + hf * E::WCOLS;
    const float* s_b2 = s_par + 4 * ND + hf * E::WCOLS;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * E::WCOLS);
    const float inv_n = 1.0f / ND;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int row0 = tile * (2 * BM) + (int)rank * BM + q * 32;  // global row of lane 0
      const int n_base = hf * E::WCOLS;
      const int crow0 = row0 % p.Beff;                              // first row in the cyclic c table
      ptx::mbar_wait(tmem_full_bar, it & 1);                        // accumulator complete => operand ring is idle
      const bool tr = warp == 2 && lane == 0 && it == 0;
      if (tr) REGEN_LTL(3);
      ptx::tcgen05_fence_after();
      if constexpr (!LN) {
        // ---- no LayerNorm: single sweep  h = acc + bias + residual -> fp32 + bf16 (hi, lo).
        // staging slots: fp32 F[2] = 0, 1 | hi tile = 2 | lo tile = 3 | residual ring (2 deep) = 4, 5
        static_assert(!CHAIN, "the no-LayerNorm sweep has no c table");
        uint8_t* rring = my + 4 * SLOT;
        if (lane == 0) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            ptx::mbar_expect_tx(&rbar[s], SLOT);
            ptx::tma_load_2d(rring + s * SLOT, &tm_res, &rbar[s], n_base + SC * s, row0);
          }
        }
        uint32_t ra[32], rb[32];
        auto sweep = [&](uint32_t (&r)[32], int sc) {
          const int slot = sc & 1, u = sc >> 1, half = sc & 1;
          uint8_t* fb = my + (sc & 1) * SLOT;
          uint8_t* hb = my + 2 * SLOT;
          uint8_t* lb = my + 3 * SLOT;
          // F[sc & 1] was stored by group sc - 2, the single hi / lo tile pair by group sc - 1 (sc even)
          if (sc >= 2 && lane == 0) {
            if (half == 0) ptx::bulk_wait_read<0>(); else ptx::bulk_wait_read<1>();
          }
          ptx::mbar_wait(&rbar[slot], (uint32_t)(it * (E::NSC / 2) + sc / 2) & 1);
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = 2 * c + jj;
              const float4 b4 = *reinterpret_cast<const float4*>(s_bias + sc * SC + 4 * j);
              const float4 rj = *reinterpret_cast<const float4*>(rring + slot * SLOT + off_f32(lane, j));
              const float z0 = __uint_as_float(r[4 * j]) + b4.x + rj.x, z1 = __uint_as_float(r[4 * j + 1]) + b4.y + rj.y;
              const float z2 = __uint_as_float(r[4 * j + 2]) + b4.z + rj.z, z3 = __uint_as_float(r[4 * j + 3]) + b4.w + rj.w;
              if (p.store_f32) *reinterpret_cast<float4*>(fb + off_f32(lane, j)) = make_float4(z0, z1, z2, z3);
              const uint32_t h0 = gemm::pack_bf16x2(z0, z1), h1 = gemm::pack_bf16x2(z2, z3);
              hw[2 * jj] = h0;
              hw[2 * jj + 1] = h1;
              lw[2 * jj] = gemm::pack_bf16x2(z0 - __uint_as_float(h0 << 16), z1 - __uint_as_float(h0 & 0xffff0000u));
              lw[2 * jj + 1] = gemm::pack_bf16x2(z2 - __uint_as_float(h1 << 16), z3 - __uint_as_float(h1 & 0xffff0000u));
            }
            if constexpr (H8) {
              // re-encode the 8 values of this chunk as fp16 + residual / hi bytes (z is rebuilt from the bf16 pair exactly)
              uint32_t fw[4];
              float lo8[8], hi8[8];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const ptx::f32x2 z = ptx::add2(ptx::bf16x2_to_f32x2(hw[e]), ptx::bf16x2_to_f32x2(lw[e]));
                float z0, z1;
                ptx::upk2(z, z0, z1);
                fw[e] = ptx::pack_f16x2_sat(z0, z1);
                const ptx::f32x2 h2 = ptx::f16x2_to_f32x2(fw[e]);
                ptx::upk2(ptx::mul2(ptx::sub2(z, h2), ptx::splat2(512.f)), lo8[2 * e], lo8[2 * e + 1]);
                ptx::upk2(ptx::mul2(h2, ptx::splat2(0.25f)), hi8[2 * e], hi8[2 * e + 1]);
              }
              *reinterpret_cast<uint4*>(hb + off_f32(lane, half * 4 + c)) = make_uint4(fw[0], fw[1], fw[2], fw[3]);
              *reinterpret_cast<uint2*>(lb + off_f32(lane, half * 2 + (c >> 1)) + (c & 1) * 8) =
                  make_uint2(ptx::pack_e4m3x4(lo8[0], lo8[1], lo8[2], lo8[3]), ptx::pack_e4m3x4(lo8[4], lo8[5], lo8[6], lo8[7]));
              *reinterpret_cast<uint2*>(lb + off_f32(lane, 4 + half * 2 + (c >> 1)) + (c & 1) * 8) =
                  make_uint2(ptx::pack_e4m3x4(hi8[0], hi8[1], hi8[2], hi8[3]), ptx::pack_e4m3x4(hi8[4], hi8[5], hi8[6], hi8[7]));
            } else {
              *reinterpret_cast<uint4*>(hb + off_f32(lane, half * 4 + c)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(lb + off_f32(lane, half * 4 + c)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();  // staging complete, every lane has read the residual slot
          if (lane == 0) {
            if (sc + 2 < E::NSC) {
              ptx::mbar_expect_tx(&rbar[slot], SLOT);
              ptx::tma_load_2d(rring + slot * SLOT, &tm_res, &rbar[slot], n_base + SC * (sc + 2), row0);
            }
            if (p.store_f32) ptx::tma_store_2d(&tm_c, fb, n_base + SC * sc, row0);
            if (half) {
              ptx::tma_store_2d(&tm_ohi, hb, n_base + 2 * SC * u, row0, p.pol_store);
              ptx::tma_store_2d(&tm_olo, lb, OLO * (n_base + 2 * SC * u), row0, p.pol_store);
            }
            ptx::bulk_commit();
          }
        };
        __syncwarp();
        ptx::tmem_ld_32x32b_x32(lane_addr, ra);
#pragma unroll 1
        for (int sc = 0; sc < E::NSC; sc += 2) {
          ptx::tmem_ld_wait(ra);
          ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)((sc + 1) * SC), rb);
          sweep(ra, sc);
          ptx::tmem_ld_wait(rb);
          if (sc + 2 < E::NSC) ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)((sc + 2) * SC), ra);
          sweep(rb, sc + 1);
        }
      } else {
      // prime the residual (and c) rings
      if (lane == 0) {
        if constexpr (R16) {
#pragma unroll
          for (int s = 0; s < RING_P; ++s) {
            ptx::mbar_expect_tx(&rbar[s], 2 * SLOT);
            ptx::tma_load_2d(res_ring + s * 2 * SLOT, &tm_ohi, &rbar[s], n_base + 2 * SC * s, row0);
            ptx::tma_load_2d(res_ring + s * 2 * SLOT + SLOT, &tm_olo, &rbar[s], OLO * (n_base + 2 * SC * s), row0);
          }
        } else {
#pragma unroll
          for (int s = 0; s < E::RING_R; ++s) {
            ptx::mbar_expect_tx(&rbar[s], SLOT);
            ptx::tma_load_2d(res_ring + s * SLOT, &tm_res, &rbar[s], n_base + SC * s, row0);
          }
        }
        if (CHAIN) {
#pragma unroll
          for (int s = 0; s < RC; ++s) {
            ptx::mbar_expect_tx(&rbar[4 + s], SLOT);
            ptx::tma_load_2d(c_ring + s * SLOT, &tm_c, &rbar[4 + s], n_base + SC * s, crow0);
          }
        }
      }
      // The three passes are software pipelined over the 32-column sub-chunks: the tcgen05.ld of sub-chunk sc + 1 is in
      // flight (second register buffer) while sub-chunk sc is processed, so the TMEM read latency (~500 cycles with
      // 8 warps sharing the 64 B/clk read port) is off each warp's critical path.
      uint32_t ra[32], rb[32];
      // ---- pass 1: v = acc + bias + residual, row statistics, v -> TMEM
      float sum = 0.f, sq = 0.f;
      ptx::f32x2 psum = ptx::splat2(0.f), psq = ptx::splat2(0.f);   // (even, odd) column partial sums of the packed passes
      auto pass1 = [&](uint32_t (&r)[32], int sc) {
        if constexpr (R16) {
          // residual = hi + lo from the pair slot of sub-chunks (2u, 2u + 1): 32 rows x 64 bf16, 128-byte rows
          const int u = sc >> 1, half = sc & 1, slot = u % RING_P;
          if (half == 0)
            ptx::mbar_wait(&rbar[slot], (uint32_t)(it * ((NP - slot + RING_P - 1) / RING_P) + u / RING_P) & 1);
          const uint8_t* hs = res_ring + slot * 2 * SLOT;
          const uint8_t* ls = hs + SLOT;
#pragma unroll
          for (int c = 0; c < 4; ++c) {  // 8 columns per 16-byte chunk of hi / lo; packed fp32 pairs (even, odd column)
            const uint4 h4 = *reinterpret_cast<const uint4*>(hs + off_f32(lane, half * 4 + c));
            if constexpr (H8) {
              // residual = fp16 + residual byte * 2^-9 (8 residual bytes of these columns: first half of the byte row)
              const uint2 l2 = *reinterpret_cast<const uint2*>(ls + off_f32(lane, half * 2 + (c >> 1)) + (c & 1) * 8);
              const float4 bA = *reinterpret_cast<const float4*>(s_bias + sc * SC + 8 * c);
              const float4 bB = *reinterpret_cast<const float4*>(s_bias + sc * SC + 8 * c + 4);
              const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, l8[4] = {l2.x, l2.x >> 16, l2.y, l2.y >> 16};
              const ptx::f32x2 bp[4] = {ptx::pk2(bA.x, bA.y), ptx::pk2(bA.z, bA.w), ptx::pk2(bB.x, bB.y), ptx::pk2(bB.z, bB.w)};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const ptx::f32x2 res = ptx::fma2(ptx::e4m3x2_to_f32x2(l8[e]), ptx::splat2(0.001953125f), ptx::f16x2_to_f32x2(hw[e]));
                const ptx::f32x2 v = ptx::add2(ptx::add2(ptx::pk2u(r[8 * c + 2 * e], r[8 * c + 2 * e + 1]), bp[e]), res);
                psum = ptx::add2(psum, v);
                psq = ptx::fma2(v, v, psq);
                ptx::upk2u(v, r[8 * c + 2 * e], r[8 * c + 2 * e + 1]);
              }
              continue;
            }
            const uint4 l4 = *reinterpret_cast<const uint4*>(ls + off_f32(lane, half * 4 + c));
            const float4 bA = *reinterpret_cast<const float4*>(s_bias + sc * SC + 8 * c);
            const float4 bB = *reinterpret_cast<const float4*>(s_bias + sc * SC + 8 * c + 4);
            const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
            const ptx::f32x2 bp[4] = {ptx::pk2(bA.x, bA.y), ptx::pk2(bA.z, bA.w), ptx::pk2(bB.x, bB.y), ptx::pk2(bB.z, bB.w)};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const ptx::f32x2 res = ptx::add2(ptx::bf16x2_to_f32x2(hw[e]), ptx::bf16x2_to_f32x2(lw[e]));
              const ptx::f32x2 v = ptx::add2(ptx::add2(ptx::pk2u(r[8 * c + 2 * e], r[8 * c + 2 * e + 1]), bp[e]), res);
              psum = ptx::add2(psum, v);
              psq = ptx::fma2(v, v, psq);
              ptx::upk2u(v, r[8 * c + 2 * e], r[8 * c + 2 * e + 1]);
            }
          }
          ptx::tmem_st_32x32b_x32(lane_addr + (uint32_t)(sc * SC), r);
          __syncwarp();  // every lane has read its half of the pair slot
          if (half == 1 && lane == 0 && u + RING_P < NP) {
            ptx::mbar_expect_tx(&rbar[slot], 2 * SLOT);
            ptx::tma_load_2d(res_ring + slot * 2 * SLOT, &tm_ohi, &rbar[slot], n_base + 2 * SC * (u + RING_P), row0);
            ptx::tma_load_2d(res_ring + slot * 2 * SLOT + SLOT, &tm_olo, &rbar[slot], OLO * (n_base + 2 * SC * (u + RING_P)), row0);
          }
        } else {
        constexpr int RING = E::RING_R;
        const int slot = sc % RING;
        // slot `slot` is filled ceil((NSC - slot) / RING) times per tile; this is fill number sc / RING of this tile
        ptx::mbar_wait(&rbar[slot], (uint32_t)(it * ((E::NSC - slot + RING - 1) / RING) + sc / RING) & 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + sc * SC + 4 * j);
          const float4 rj = *reinterpret_cast<const float4*>(res_ring + slot * SLOT + off_f32(lane, j));
          float v0 = __uint_as_float(r[4 * j]) + b4.x + rj.x, v1 = __uint_as_float(r[4 * j + 1]) + b4.y + rj.y;
          float v2 = __uint_as_float(r[4 * j + 2]) + b4.z + rj.z, v3 = __uint_as_float(r[4 * j + 3]) + b4.w + rj.w;
          sum += (v0 + v1) + (v2 + v3);
          sq = fmaf(v0, v0, sq); sq = fmaf(v1, v1, sq); sq = fmaf(v2, v2, sq); sq = fmaf(v3, v3, sq);
          r[4 * j] = __float_as_uint(v0); r[4 * j + 1] = __float_as_uint(v1);
          r[4 * j + 2] = __float_as_uint(v2); r[4 * j + 3] = __float_as_uint(v3);
        }
        ptx::tmem_st_32x32b_x32(lane_addr + (uint32_t)(sc * SC), r);
        __syncwarp();  // every lane has read the slot
        if (lane == 0 && sc + RING < E::NSC) {
          ptx::mbar_expect_tx(&rbar[slot], SLOT);
          ptx::tma_load_2d(res_ring + slot * SLOT, &tm_res, &rbar[slot], n_base + SC * (sc + RING), row0);
        }
        }
      };
      __syncwarp();
      ptx::tmem_ld_32x32b_x32(lane_addr, ra);
#pragma unroll 1
      for (int sc = 0; sc < E::NSC; sc += 2) {
        ptx::tmem_ld_wait(ra);
        ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)((sc + 1) * SC), rb);
        pass1(ra, sc);
        ptx::tmem_ld_wait(rb);
        if (sc + 2 < E::NSC) ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)((sc + 2) * SC), ra);
        pass1(rb, sc + 1);
      }
      ptx::tmem_st_wait();
      if (tr) REGEN_LTL(4);
      if constexpr (R16) {
        float a, b;
        ptx::upk2(psum, a, b);
        sum = a + b;
        ptx::upk2(psq, a, b);
        sq = a + b;
      }
      // exchange the half-row statistics with the warp that owns the other 256 columns of the same rows
      s_stats[r_local * 4 + hf] = make_float2(sum, sq);
      named_bar_sync(1 + q, 32 * E::PARTS);
      if (tr) REGEN_LTL(5);
      sum = 0.f;
      sq = 0.f;
#pragma unroll
      for (int pp = 0; pp < E::PARTS; ++pp) {  // same order in every warp: bit-identical statistics across the row
        const float2 o = s_stats[r_local * 4 + pp];
        sum += o.x;
        sq += o.y;
      }
      float mean = sum * inv_n;
      float rstd = 1.0f / sqrtf(fmaxf(sq * inv_n - mean * mean, 0.f) + p.ln_eps);
      // (x - mean) * rstd * g + b as two FMAs: x * rstd + (-mean * rstd), then * g + b (the passes are issue-bound)
      float nmr = -mean * rstd;

      if (CHAIN) {
        // ---- pass 2: y = LN1(v) + c, statistics of y, y -> TMEM
        float sum2 = 0.f, sq2 = 0.f;
        ptx::f32x2 sum2b = ptx::splat2(0.f), sq2b = ptx::splat2(0.f);
        auto pass2 = [&](uint32_t (&r)[32], int sc) {
          constexpr int RING = RC;
          const int slot = sc % RING;
          ptx::mbar_wait(&rbar[4 + slot], (uint32_t)(it * ((E::NSC - slot + RING - 1) / RING) + sc / RING) & 1);
          if constexpr (R16) {
            const ptx::f32x2 rs2 = ptx::splat2(rstd), nm2 = ptx::splat2(nmr);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 g4 = *reinterpret_cast<const float4*>(s_g1 + sc * SC + 4 * j);
              const float4 b4 = *reinterpret_cast<const float4*>(s_b1 + sc * SC + 4 * j);
              const float4 cj = *reinterpret_cast<const float4*>(c_ring + slot * SLOT + off_f32(lane, j));
              const ptx::f32x2 gp[2] = {ptx::pk2(g4.x, g4.y), ptx::pk2(g4.z, g4.w)};
              const ptx::f32x2 bp[2] = {ptx::pk2(b4.x, b4.y), ptx::pk2(b4.z, b4.w)};
              const ptx::f32x2 cp[2] = {ptx::pk2(cj.x, cj.y), ptx::pk2(cj.z, cj.w)};
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const ptx::f32x2 v = ptx::pk2u(r[4 * j + 2 * e], r[4 * j + 2 * e + 1]);
                const ptx::f32x2 y = ptx::add2(ptx::fma2(ptx::fma2(v, rs2, nm2), gp[e], bp[e]), cp[e]);
                sum2b = ptx::add2(sum2b, y);
                sq2b = ptx::fma2(y, y, sq2b);
                ptx::upk2u(y, r[4 * j + 2 * e], r[4 * j + 2 * e + 1]);
              }
            }
          } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g4 = *reinterpret_cast<const float4*>(s_g1 + sc * SC + 4 * j);
            const float4 b4 = *reinterpret_cast<const float4*>(s_b1 + sc * SC + 4 * j);
            const float4 cj = *reinterpret_cast<const float4*>(c_ring + slot * SLOT + off_f32(lane, j));
            float y0 = fmaf(fmaf(__uint_as_float(r[4 * j]), rstd, nmr), g4.x, b4.x) + cj.x;
            float y1 = fmaf(fmaf(__uint_as_float(r[4 * j + 1]), rstd, nmr), g4.y, b4.y) + cj.y;
            float y2 = fmaf(fmaf(__uint_as_float(r[4 * j + 2]), rstd, nmr), g4.z, b4.z) + cj.z;
            float y3 = fmaf(fmaf(__uint_as_float(r[4 * j + 3]), rstd, nmr), g4.w, b4.w) + cj.w;
            sum2 += (y0 + y1) + (y2 + y3);
            sq2 = fmaf(y0, y0, sq2); sq2 = fmaf(y1, y1, sq2); sq2 = fmaf(y2, y2, sq2); sq2 = fmaf(y3, y3, sq2);
            r[4 * j] = __float_as_uint(y0); r[4 * j + 1] = __float_as_uint(y1);
            r[4 * j + 2] = __float_as_uint(y2); r[4 * j + 3] = __float_as_uint(y3);
          }
          }
          ptx::tmem_st_32x32b_x32(lane_addr + (uint32_t)(sc * SC), r);
          __syncwarp();  // every lane has read the slot
          if (lane == 0 && sc + RING < E::NSC) {
            ptx::mbar_expect_tx(&rbar[4 + slot], SLOT);
            ptx::tma_load_2d(c_ring + slot * SLOT, &tm_c, &rbar[4 + slot], n_base + SC * (sc + RING), crow0);
          }
        };
        __syncwarp();
        ptx::tmem_ld_32x32b_x32(lane_addr, ra);
#pragma unroll 1
        for (int sc = 0; sc < E::NSC; sc += 2) {
          ptx::tmem_ld_wait(ra);
          ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)((sc + 1) * SC), rb);
          pass2(ra, sc);
          ptx::tmem_ld_wait(rb);
          if (sc + 2 < E::NSC) ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)((sc + 2) * SC), ra);
          pass2(rb, sc + 1);
        }
        ptx::tmem_st_wait();
        if (tr) REGEN_LTL(6);
        if constexpr (R16) {
          float a, b;
          ptx::upk2(sum2b, a, b);
          sum2 = a + b;
          ptx::upk2(sq2b, a, b);
          sq2 = a + b;
        }
        s_stats[512 + r_local * 4 + hf] = make_float2(sum2, sq2);
        named_bar_sync(1 + q, 32 * E::PARTS);
        if (tr) REGEN_LTL(7);
        sum2 = 0.f;
        sq2 = 0.f;
#pragma unroll
        for (int pp = 0; pp < E::PARTS; ++pp) {
          const float2 o = s_stats[512 + r_local * 4 + pp];
          sum2 += o.x;
          sq2 += o.y;
        }
        mean = sum2 * inv_n;
        rstd = 1.0f / sqrtf(fmaxf(sq2 * inv_n - mean * mean, 0.f) + p.ln_eps);
        nmr = -mean * rstd;
      }

      // ---- final pass: z = LN(.), fp32 + bf16 (hi, lo) out through TMA stores.  Measured on B200: the TMA unit retires
      // about one box row per 2.6 cycles whatever its length (<= 128 B), and that row rate -- not shared-memory or TMEM
      // bandwidth -- bounds this pass, so the bf16 halves are staged as 32 x 64 tiles (128-byte rows, one store per
      // PAIR of sub-chunks): 128 instead of 192 row requests per 64 columns.  Staging per warp: fp32 slots F[NOB],
      // hi tiles H[NOB], lo tiles L[NOB] (4 KB each); every sub-chunk commits one bulk group (the odd ones carry the
      // bf16 tiles of their pair).
      const float* gg = CHAIN ? s_g2 : s_g1;
      const float* bb = CHAIN ? s_b2 : s_b1;
      auto pass3 = [&](uint32_t (&r)[32], int sc) {
        const int u = sc >> 1, half = sc & 1;
        const int nob = R16 ? p.nob16 : E::NOB;
        uint8_t* fb = out_buf + (sc % E::NOB) * SLOT;
        uint8_t* hb = out_buf + (R16 ? 2 * (u % nob) : E::NOB + (u % E::NOB)) * SLOT;
        uint8_t* lb = out_buf + (R16 ? 2 * (u % nob) + 1 : 2 * E::NOB + (u % E::NOB)) * SLOT;
        if (tr && sc == 2) REGEN_LTL(11);
        if constexpr (R16) {
          // one bulk group per sub-chunk PAIR: H/L[u % nob] were last stored by the group of pair u - nob
          if (half == 0 && u >= nob && lane == 0) {
            if (nob == 3) ptx::bulk_wait_read<2>(); else ptx::bulk_wait_read<1>();
          }
        } else {
        // F[sc % NOB] was last stored by group sc - NOB; H/L[u % NOB] by group 2 (u - NOB) + 1 = sc - 2 NOB + 1 (sc even):
        // both are complete once at most NOB - 1 groups are pending
        if (sc >= E::NOB && lane == 0) ptx::bulk_wait_read<E::NOB - 1>();
        }
        __syncwarp();
        if (tr && sc == 2) REGEN_LTL(12);
        if constexpr (R16) {
          const ptx::f32x2 rs2 = ptx::splat2(rstd), nm2 = ptx::splat2(nmr);
#pragma unroll
          for (int c = 0; c < 4; ++c) {  // 8 columns -> one bf16 hi chunk, one bf16 lo chunk; packed fp32 pairs
            uint32_t hw[4], lw[4];
            float lo8[8], hi8[8];   // H8 only
            const float4 gA = *reinterpret_cast<const float4*>(gg + sc * SC + 8 * c);
            const float4 gB = *reinterpret_cast<const float4*>(gg + sc * SC + 8 * c + 4);
            const float4 bA = *reinterpret_cast<const float4*>(bb + sc * SC + 8 * c);
            const float4 bB = *reinterpret_cast<const float4*>(bb + sc * SC + 8 * c + 4);
            const ptx::f32x2 gp[4] = {ptx::pk2(gA.x, gA.y), ptx::pk2(gA.z, gA.w), ptx::pk2(gB.x, gB.y), ptx::pk2(gB.z, gB.w)};
            const ptx::f32x2 bp[4] = {ptx::pk2(bA.x, bA.y), ptx::pk2(bA.z, bA.w), ptx::pk2(bB.x, bB.y), ptx::pk2(bB.z, bB.w)};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const ptx::f32x2 z = ptx::fma2(ptx::fma2(ptx::pk2u(r[8 * c + 2 * e], r[8 * c + 2 * e + 1]), rs2, nm2), gp[e], bp[e]);
              float z0, z1, l0, l1;
              ptx::upk2(z, z0, z1);
              hw[e] = gemm::pack_bf16x2(z0, z1);
              ptx::upk2(ptx::sub2(z, ptx::bf16x2_to_f32x2(hw[e])), l0, l1);   // float(hi) = the bf16 bits in the upper half
              lw[e] = gemm::pack_bf16x2(l0, l1);
              if constexpr (H8) {   // fp16 + residual / hi bytes instead of the bf16 pair
                hw[e] = ptx::pack_f16x2_sat(z0, z1);
                const ptx::f32x2 h2 = ptx::f16x2_to_f32x2(hw[e]);
                ptx::upk2(ptx::mul2(ptx::sub2(z, h2), ptx::splat2(512.f)), lo8[2 * e], lo8[2 * e + 1]);
                ptx::upk2(ptx::mul2(h2, ptx::splat2(0.25f)), hi8[2 * e], hi8[2 * e + 1]);
              }
            }
            // 16-byte chunk (half * 4 + c) of this row's 128-byte bf16 row (same SWIZZLE_128B pattern as the fp32 tile)
            *reinterpret_cast<uint4*>(hb + off_f32(lane, half * 4 + c)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            if constexpr (H8) {  // byte tile: residual bytes in chunks 0..3, hi bytes in chunks 4..7, 8 bytes per 8 columns
              *reinterpret_cast<uint2*>(lb + off_f32(lane, half * 2 + (c >> 1)) + (c & 1) * 8) =
                  make_uint2(ptx::pack_e4m3x4(lo8[0], lo8[1], lo8[2], lo8[3]), ptx::pack_e4m3x4(lo8[4], lo8[5], lo8[6], lo8[7]));
              *reinterpret_cast<uint2*>(lb + off_f32(lane, 4 + half * 2 + (c >> 1)) + (c & 1) * 8) =
                  make_uint2(ptx::pack_e4m3x4(hi8[0], hi8[1], hi8[2], hi8[3]), ptx::pack_e4m3x4(hi8[4], hi8[5], hi8[6], hi8[7]));
            } else {
              *reinterpret_cast<uint4*>(lb + off_f32(lane, half * 4 + c)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
        } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {  // 8 columns = two fp32 chunks, one bf16 hi chunk, one bf16 lo chunk
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * c + jj;
            const float4 g4 = *reinterpret_cast<const float4*>(gg + sc * SC + 4 * j);
            const float4 b4 = *reinterpret_cast<const float4*>(bb + sc * SC + 4 * j);
            const float z0 = fmaf(fmaf(__uint_as_float(r[4 * j]), rstd, nmr), g4.x, b4.x);
            const float z1 = fmaf(fmaf(__uint_as_float(r[4 * j + 1]), rstd, nmr), g4.y, b4.y);
            const float z2 = fmaf(fmaf(__uint_as_float(r[4 * j + 2]), rstd, nmr), g4.z, b4.z);
            const float z3 = fmaf(fmaf(__uint_as_float(r[4 * j + 3]), rstd, nmr), g4.w, b4.w);
            if constexpr (!R16) *reinterpret_cast<float4*>(fb + off_f32(lane, j)) = make_float4(z0, z1, z2, z3);
            const uint32_t h0 = gemm::pack_bf16x2(z0, z1), h1 = gemm::pack_bf16x2(z2, z3);
            hw[2 * jj] = h0;
            hw[2 * jj + 1] = h1;
            lw[2 * jj] = gemm::pack_bf16x2(z0 - __uint_as_float(h0 << 16), z1 - __uint_as_float(h0 & 0xffff0000u));
            lw[2 * jj + 1] = gemm::pack_bf16x2(z2 - __uint_as_float(h1 << 16), z3 - __uint_as_float(h1 & 0xffff0000u));
          }
          // 16-byte chunk (half * 4 + c) of this row's 128-byte bf16 row (same SWIZZLE_128B pattern as the fp32 tile)
          *reinterpret_cast<uint4*>(hb + off_f32(lane, half * 4 + c)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(lb + off_f32(lane, half * 4 + c)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        }
        if (tr && sc == 2) REGEN_LTL(13);
        if constexpr (R16) {
          if (half) {  // the pair's hi / lo tiles are complete: one fence, two stores, one group
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_2d(&tm_ohi, hb, n_base + 2 * SC * u, row0, p.pol_store);
              ptx::tma_store_2d(&tm_olo, lb, OLO * (n_base + 2 * SC * u), row0, p.pol_store);
              ptx::bulk_commit();
            }
          }
        } else {
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (tr && sc == 2) REGEN_LTL(14);
        if (lane == 0) {
          ptx::tma_store_2d(&tm_res, fb, n_base + SC * sc, row0);
          if (half) {
            ptx::tma_store_2d(&tm_ohi, hb, n_base + 2 * SC * u, row0, p.pol_store);
            ptx::tma_store_2d(&tm_olo, lb, n_base + 2 * SC * u, row0, p.pol_store);
          }
          ptx::bulk_commit();
        }
        }
      };
      __syncwarp();  // both rings fully consumed by every lane: their memory becomes the output buffers
      ptx::tmem_ld_32x32b_x32(lane_addr, ra);
#pragma unroll 1
      for (int sc = 0; sc < E::NSC; sc += 2) {
        ptx::tmem_ld_wait(ra);
        ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)((sc + 1) * SC), rb);
        pass3(ra, sc);
        if (tr && sc == 2) REGEN_LTL(15);
        ptx::tmem_ld_wait(rb);
        if (tr && sc == 2) REGEN_LTL(16);
        if (sc + 2 < E::NSC) ptx::tmem_ld_32x32b_x32(lane_addr + (uint32_t)((sc + 2) * SC), ra);
        pass3(rb, sc + 1);
      }
      // accumulator and operand ring are free again
      }
      if (tr) REGEN_LTL(8);
      if (lane == 0) ptx::bulk_wait_read<0>();
      if (tr) REGEN_LTL(9);
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive_remote(tmem_empty_bar, 0);
        ptx::mbar_arrive(epi_done_bar);
      }
      named_bar_sync(1 + q, 32 * E::PARTS);  // the statistics slots are rewritten by the next tile
    }
    // every tile already waited until its stores had READ the staging tiles; the kernel boundary (griddepcontrol.wait in
    // the dependent grid) orders the global writes, so waiting for their completion here only lengthens the exit
    if (lane == 0 && p.exit_wait_full) ptx::bulk_wait<0>();
  }

  ptx::tcgen05_fence_before();
  ptx::cluster_sync();
  ptx::steplog_end(p.steplog, p.steplog_slot, p.steplog_cta);
  if (threadIdx.x == 0) REGEN_LTL(10);
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <bool SPLIT, bool CHAIN, int EW, bool LN = true, bool R16 = false, bool M8 = false, bool H8 = false>
inline cudaError_t launch_impl(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                               const CUtensorMap& w_lo, const CUtensorMap& res, const CUtensorMap& c,
                               const CUtensorMap& ohi, const CUtensorMap& olo, const Params& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ln_kernel<SPLIT, CHAIN, EW, LN, R16, M8, H8>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int64_t tiles = ceil_div(p.M, 2 * BM);
  const int clusters = (int)(tiles < kNumSMs / 2 ? tiles : kNumSMs / 2);
  return launch_pdl(gemm_ln_kernel<SPLIT, CHAIN, EW, LN, R16, M8, H8>, dim3(2 * clusters), dim3(Epi<EW, CHAIN>::THREADS), SMEM_BYTES,
                    stream, a_hi, a_lo, w_hi, w_lo, res, c, ohi, olo, p);
}

// 8 epilogue warps per CTA by default.  Measured on B200 (round 1): 16 warps change the three epilogue passes by < 10 %
// (10.2k/8.9k/16.8k vs 11.2k/9.8k/18.6k cycles) -- the passes are bound by TMEM and shared-memory traffic (256 KB read +
// 256 KB written per pass per CTA), not by per-warp latency -- and cost register spills at the 96-register cap of 576
// threads, so 16 stays an A/B switch (REGEN_DEBUG_LN_EW=16, fp32-residual variant only).
// r16: residual stream as bf16 (hi, lo) only (see the header comment); `res` is then unused, ohi / olo are loaded AND stored.
template <bool SPLIT, bool CHAIN>
inline cudaError_t launch(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                          const CUtensorMap& w_lo, const CUtensorMap& res, const CUtensorMap& c, const CUtensorMap& ohi,
                          const CUtensorMap& olo, const Params& p, cudaStream_t stream, bool r16) {
  static int ew = 0;
  if (!ew) {
    const char* e = getenv("REGEN_DEBUG_LN_EW");
    ew = (e && atoi(e) == 16) ? 16 : 8;
  }
  if (r16) return launch_impl<SPLIT, CHAIN, 8, true, true>(a_hi, a_lo, w_hi, w_lo, res, c, ohi, olo, p, stream);
  return ew == 16 ? launch_impl<SPLIT, CHAIN, 16>(a_hi, a_lo, w_hi, w_lo, res, c, ohi, olo, p, stream)
                  : launch_impl<SPLIT, CHAIN, 8>(a_hi, a_lo, w_hi, w_lo, res, c, ohi, olo, p, stream);
}

// mixed8 operands (see the header comment): a16 / a8 / w16 / w8 maps, (hi, lo) residual stream; K must be a multiple of 128
// h8: the residual stream is the mixed8 pack too (ohi = fp16 map, olo = byte map [M, 1024] with 32 x 128-byte boxes)
template <bool CHAIN>
inline cudaError_t launch_m8(const CUtensorMap& a16, const CUtensorMap& a8, const CUtensorMap& w16, const CUtensorMap& w8,
                             const CUtensorMap& c, const CUtensorMap& ohi, const CUtensorMap& olo, const Params& p,
                             cudaStream_t stream, bool h8 = false) {
  if (h8) return launch_impl<true, CHAIN, 8, true, true, true, true>(a16, a8, w16, w8, ohi, c, ohi, olo, p, stream);
  return launch_impl<true, CHAIN, 8, true, true, true>(a16, a8, w16, w8, ohi, c, ohi, olo, p, stream);
}

// h <- A.W^T + bias + residual without LayerNorm (input projection): `res` = residual LOAD map, `out32` = fp32 STORE map
// (written only with Params::store_f32)
template <bool SPLIT>
inline cudaError_t launch_noln(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                               const CUtensorMap& w_lo, const CUtensorMap& res, const CUtensorMap& out32,
                               const CUtensorMap& ohi, const CUtensorMap& olo, const Params& p, cudaStream_t stream) {
  return launch_impl<SPLIT, false, 8, false>(a_hi, a_lo, w_hi, w_lo, res, out32, ohi, olo, p, stream);
}
// the same with h written as the mixed8 pack (ohi = fp16 map, olo = byte map): input projection under precision 'mixed8h'
inline cudaError_t launch_noln_h8(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                                  const CUtensorMap& w_lo, const CUtensorMap& res, const CUtensorMap& out32,
                                  const CUtensorMap& ohi, const CUtensorMap& olo, const Params& p, cudaStream_t stream) {
  return launch_impl<true, false, 8, false, false, false, true>(a_hi, a_lo, w_hi, w_lo, res, out32, ohi, olo, p, stream);
}

}  // namespace gemmln
}  // namespace regen
