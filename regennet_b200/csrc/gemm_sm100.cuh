// Warp-specialised tcgen05 GEMM for sm_100a:   C[M,N] = epilogue( A[M,K] . W[N,K]^T )
//
//   * operands are bf16, K-major, staged global -> shared by TMA (128-byte swizzle) into a
//     multi-stage mbarrier ring; one elected thread issues tcgen05.mma (UMMA 128 x BN x 16,
//     cta_group::1) accumulating fp32 in tensor memory;
//   * SPLIT mode ("bf16x3"): every fp32 operand is carried as a (hi, lo) bf16 pair and each k-step
//     issues three MMAs  A_hi.W_hi + A_hi.W_lo + A_lo.W_hi  into the same TMEM accumulator, which
//     restores ~16 significand bits (max abs error ~2e-5 on the CMDM forward instead of 1e-2);
//   * epilogue: 8 warps (16 in the pair kernel's no-residual variant) read the accumulator with tcgen05.ld (32x32b:
//     one thread = one output row), add bias / residual, optionally apply exact-erf GELU, and write fp32 and/or a bf16
//     (hi, lo) pair for the next GEMM's A operand -- staged in shared memory as swizzled TMA boxes and stored with
//     cp.async.bulk.tensor.
//
// Two kernels: gemm_tn_kernel (one CTA per 128 x BN tile; tiny batches, M <= 128) and gemm2_tn_kernel (CTA pair,
// cta_group::2, 256 x BN tiles; everything else).  Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2.. = epilogue.  Both are persistent (one CTA per SM looping over tiles); the fp32 accumulator is double buffered
// in TMEM so the epilogue of one tile overlaps the main loop of the next; every kernel is launched with programmatic
// stream serialization and waits (griddepcontrol.wait) after its prologue.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace regen {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 320;  // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue

template <int BN, bool SPLIT>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int W_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_BYTES + W_BYTES);
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;  // 2 for SPLIT @ BN=256, 4 for plain bf16
  static constexpr int STG_BYTES = 8 * 4096;  // kEpiWarps * STG_WARP_BYTES (epilogue staging, see epilogue_slice)
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment
  static constexpr int TMEM_COLS = 2 * BN;  // double-buffered accumulator
  static_assert(STAGES >= 2, "need at least a double buffer");
  static_assert(BN == 64 || BN == 128 || BN == 256, "BN must be 64, 128 or 256 (TMEM_COLS a power of two <= 512)");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct Params {
  int M, N, K;              // K must be a multiple of 64; rows >= M / cols >= N are masked
  const float* bias;        // [N] or null
  const float* residual;    // [M, ld_res] fp32 or null (added before the activation)
  int ld_res;
  float* out_f32;           // [M, ld_out] or null
  int ld_out;
  __nv_bfloat16* out_hi;    // [M, ld_split] or null: bf16 (hi, lo) split of the result
  __nv_bfloat16* out_lo;
  int ld_split;
  int gelu;                 // exact-erf GELU after bias/residual
  int tma_store;            // 1: outputs are written with TMA stores through the tm_o* tensor maps (needs N % 4 == 0)
  int exit_wait_full;       // 1: wait for the bulk stores' global writes before exit (default; REGEN_DEBUG_EXIT_WAIT_READ=1 clears it, A/B: no measurable difference)
  int pair64;               // 16-warp pair kernel, bf16 (hi, lo) outputs only: 64-column store boxes through warp pairs (tm_ohi /
                            // tm_olo must then be the 32 x 64 box maps)
  int m8;                   // pair64 epilogue only: outputs in the mixed8 operand format of the fused linear2 + LayerNorm kernel
                            // (gemm_ln_sm100.cuh): tm_ohi = fp16 [M, N] (box 32 x 64), tm_olo = bytes [M, 2 N] (box 32 rows x 128 B,
                            // SWIZZLE_128B): per group of 64 columns, 64 bytes e4m3((v - fp16(v)) * 2^9) | 64 bytes e4m3(fp16(v) / 4)
  int slice_w_rows;         // pair kernel: rows of the W box the slice maps (tm_ws_*) load per CTA (set by launch2)
  int tail_split;           // pair kernel: the tiles of the last, partial wave are cut into 1 / 2 / 4 column slices (set by launch2)
  // bring-up instrumentation (test hook only, null in production): CTA 0 records clock64() at pipeline events
  //   [0] kernel entry  [1] setup done  [2] kernel exit  [8+2i] MMA of tile i: operands of first k-block landed
  //   [9+2i] MMA of tile i: last instruction issued  [40+2i] epilogue of tile i: accumulator ready  [41+2i] done
  //   [80..] fine-grained stamps inside the first epilogue sub-chunks of tile 0 (warp 2)
  unsigned long long* timeline;
  unsigned long long* steplog;  // whole-step timeline (ptx::steplog_begin / steplog_end), null in production
  int steplog_slot, steplog_cta;  // steplog_cta: word offset of the per-CTA exit-time table (0 = off)
  // L2 eviction-priority hints of the TMA traffic (ptx::kL2Evict*, 0 = none): W tiles (re-read by every row block) and
  // bf16 (hi, lo) outputs (the next kernel's operands)
  unsigned long long pol_w, pol_store;
  int a_rows;               // single-CTA kernel: rows of the A box the tm_a_* maps load (0 = 128).  With at most 64 token rows
                            // (one sample of T <= 64 frames) the 64-row box halves the A bytes per k-block; rows 64..127 of the
                            // staged tile are then never written, and the accumulator rows they feed are never read
};

#define REGEN_TL(slot)                                                        \
  do {                                                                        \
    if (p.timeline && blockIdx.x == 0) p.timeline[(slot)] = (unsigned long long)clock64(); \
  } while (0)

// GELU(v) = v/2 (1 + erf(v/sqrt 2)), the exact-erf form torch's activation='gelu' uses.  erf is evaluated with the
// Abramowitz-Stegun 7.1.26 rational/exponential form (|erf error| <= 5.5e-7 in fp32, |GELU error| <= 4.7e-7 measured
// over [-8, 8]) instead of libdevice's erff: ~15 instructions instead of ~40, which matters because the FFN1
// epilogue evaluates 15.7 M GELUs per layer on 8 warps per SM next to the tensor-core main loop.
__device__ __forceinline__ float gelu_erf(float v) {
  const float z = fabsf(v) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));  // MUFU.RCP (an IEEE __frcp_rn costs ~10x more)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = 1.0f - poly * t * __expf(-z * z);  // erf(|v|/sqrt 2)
  return 0.5f * v + 0.5f * fabsf(v) * e;             // v/2 (1 + sign(v) e)
}

// The same GELU on a packed (even, odd column) fp32 pair: FFMA2 / FMUL2 halve the issue slots of the polynomial, the two
// MUFU (reciprocal, exp2) stay scalar.  Bit-for-bit the scalar formula per half (each packed op is an IEEE fp32 op).
__device__ __forceinline__ ptx::f32x2 gelu_erf2(ptx::f32x2 v) {
  uint32_t v0, v1;
  ptx::upk2u(v, v0, v1);
  const ptx::f32x2 av = ptx::pk2u(v0 & 0x7fffffffu, v1 & 0x7fffffffu);                       // |v|
  const ptx::f32x2 z = ptx::mul2(av, ptx::splat2(0.70710678118654752440f));
  float d0, d1;
  ptx::upk2(ptx::fma2(ptx::splat2(0.3275911f), z, ptx::splat2(1.0f)), d0, d1);
  const ptx::f32x2 t = ptx::pk2(__fdividef(1.0f, d0), __fdividef(1.0f, d1));
  ptx::f32x2 poly = ptx::fma2(ptx::splat2(1.061405429f), t, ptx::splat2(-1.453152027f));
  poly = ptx::fma2(poly, t, ptx::splat2(1.421413741f));
  poly = ptx::fma2(poly, t, ptx::splat2(-0.284496736f));
  poly = ptx::fma2(poly, t, ptx::splat2(0.254829592f));
  float q0, q1;
  ptx::upk2(ptx::mul2(z, z), q0, q1);
  const ptx::f32x2 nex = ptx::pk2(-__expf(-q0), -__expf(-q1));
  const ptx::f32x2 e = ptx::fma2(ptx::mul2(poly, t), nex, ptx::splat2(1.0f));                // erf(|v| / sqrt 2)
  const ptx::f32x2 half = ptx::splat2(0.5f);
  return ptx::fma2(ptx::mul2(half, av), e, ptx::mul2(half, v));                              // v/2 (1 + sign(v) e)
}

// Epilogue of a 32-row x ncols slice of an accumulator tile, executed by one warp (lane = TMEM lane = output row).
//   acc   TMEM address of (first lane of this warp's quarter, first column of the slice)
//   row0  global row of lane 0, n0 global column of the slice's first column
//   stg   this warp's private shared-memory staging area (STG_WARP_BYTES = two 2 KB buffers, 1 KB aligned)
// The accumulator is consumed in 16-column sub-chunks.  Each sub-chunk is staged in shared memory as a dense
// 32-row tile in exactly the layout of a TMA box (fp32: 64-byte rows, SWIZZLE_64B; bf16: 32-byte rows,
// SWIZZLE_32B), which makes the row-per-thread accesses bank-conflict free, and is then written by ONE
// cp.async.bulk.tensor store (clipped at M / N by the tensor map) -- no LSU global stores on the path.  Without
// output tensor maps (tma_store == 0) the staged tile is read back in a coalesced mapping and stored with st.global.
constexpr int STG_WARP_BYTES = 4096;               // 2 buffers x 2 KB
constexpr int kEpiWarps = 8;                       // 2 warps per TMEM lane quarter, each takes half of the columns

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of 16-byte chunk `c` of row `r` inside a staged tile
__device__ __forceinline__ int stg_off_f32(int r, int c) { return r * 64 + ((c ^ ((r >> 1) & 3)) << 4); }   // SWIZZLE_64B
__device__ __forceinline__ int stg_off_128(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }          // SWIZZLE_128B
__device__ __forceinline__ int stg_off_bf16(int r, int c) { return r * 32 + ((c ^ ((r >> 2) & 1)) << 4); }  // SWIZZLE_32B

#define REGEN_TLF(k)                                                                              \
  do {                                                                                            \
    if (trace && lane == 0 && c0 < 48) p.timeline[80 + (c0 / 16) * 8 + (k)] = (unsigned long long)clock64(); \
  } while (0)

template <bool RES = true, bool DBUF = true>
__device__ __forceinline__ void epilogue_slice(const Params& p, const CUtensorMap* tm_o32, const CUtensorMap* tm_ohi,
                                               const CUtensorMap* tm_olo, uint8_t* stg_raw, uint32_t acc, int row0,
                                               int n0, int ncols, int lane, bool trace_req = false) {
  const bool trace = trace_req && p.timeline && blockIdx.x == 0;
  const bool vec_ok = (p.N & 3) == 0;
  const int fr = lane >> 2, fch = lane & 3;        // fp32 coalesced mapping: row i*8 + fr, 16-byte chunk fch
  const int hr = lane >> 1, hch = lane & 1;        // bf16 coalesced mapping: row i*16 + hr, 16-byte chunk hch
  if (vec_ok) {
    // buffers may still be the source of bulk stores issued for the previous tile
    if (p.tma_store) {
      if (lane == 0) ptx::bulk_wait_read<0>();
      __syncwarp();
    }
    float4 rpre[4];
    auto prefetch_res = [&](int c0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = row0 + i * 8 + fr, n = n0 + c0 + fch * 4;
        rpre[i] = (r < p.M && n < p.N) ? *reinterpret_cast<const float4*>(p.residual + (size_t)r * p.ld_res + n)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    // bias of the next sub-chunk is prefetched as well: with ~220 KB of the SM's unified L1/shared memory carved out
    // as shared memory there is practically no L1, so every global load is an L2 round trip (~800 cycles under load)
    float4 bpre[4];
    auto prefetch_bias = [&](int c0) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        bpre[j] = (n0 + c0 + 4 * j < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + 4 * j))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    const bool has_res = RES && p.residual != nullptr;
    if (has_res) prefetch_res(0);
    if (p.bias) prefetch_bias(0);
    int sc = 0;
#pragma unroll 1
    for (int c0 = 0; c0 < ncols; c0 += 16, ++sc) {
      if (n0 + c0 >= p.N) break;  // warp-uniform
      uint8_t* buf = stg_raw + (DBUF ? (sc & 1) * 2048 : 0);
      if (p.tma_store && sc >= (DBUF ? 2 : 1)) {  // the previous store from this buffer has finished reading it
        if (lane == 0) {
          if (DBUF) ptx::bulk_wait_read<1>(); else ptx::bulk_wait_read<0>();
        }
      }
      uint32_t r[16];
      __syncwarp();
      REGEN_TLF(0);
      ptx::tmem_ld_32x32b_x16(acc + (uint32_t)c0, r);
      if (has_res) {
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(buf + stg_off_f32(i * 8 + fr, fch)) = rpre[i];
        __syncwarp();
        REGEN_TLF(1);
        if (c0 + 16 < ncols && n0 + c0 + 16 < p.N) prefetch_res(c0 + 16);
      }
      ptx::tmem_ld_wait();
      REGEN_TLF(2);
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[4 * j] += bpre[j].x; v[4 * j + 1] += bpre[j].y; v[4 * j + 2] += bpre[j].z; v[4 * j + 3] += bpre[j].w;
        }
        if (c0 + 16 < ncols && n0 + c0 + 16 < p.N) prefetch_bias(c0 + 16);
      }
      if (has_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 r4 = *reinterpret_cast<const float4*>(buf + stg_off_f32(lane, j));
          v[4 * j] += r4.x; v[4 * j + 1] += r4.y; v[4 * j + 2] += r4.z; v[4 * j + 3] += r4.w;
        }
        __syncwarp();
      }
      REGEN_TLF(3);
      if (p.gelu) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
      }
      REGEN_TLF(4);
      if (p.out_f32) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(buf + stg_off_f32(lane, j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        if (p.tma_store) {
          ptx::fence_proxy_async_smem();
          __syncwarp();
          REGEN_TLF(5);
          if (lane == 0) {
            ptx::tma_store_2d(tm_o32, buf, n0 + c0, row0);
            ptx::bulk_commit();
            if (p.out_hi) ptx::bulk_wait_read<0>();  // the bf16 pair is staged in the same buffer next
          }
          __syncwarp();
        } else {
          __syncwarp();
          REGEN_TLF(5);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = row0 + i * 8 + fr, n = n0 + c0 + fch * 4;
            if (rr < p.M && n < p.N)
              *reinterpret_cast<float4*>(p.out_f32 + (size_t)rr * p.ld_out + n) =
                  *reinterpret_cast<const float4*>(buf + stg_off_f32(i * 8 + fr, fch));
          }
          __syncwarp();
        }
        REGEN_TLF(6);
      }
      if (p.out_hi) {
        uint32_t hw[8], lw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = v[2 * j], b = v[2 * j + 1];
          hw[j] = pack_bf16x2(a, b);
          // float(hi) is the bf16 bit pattern in the upper half of the word
          lw[j] = pack_bf16x2(a - __uint_as_float(hw[j] << 16), b - __uint_as_float(hw[j] & 0xffff0000u));
        }
        uint8_t* bh = buf;
        uint8_t* bl = buf + 1024;
        *reinterpret_cast<uint4*>(bh + stg_off_bf16(lane, 0)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(bh + stg_off_bf16(lane, 1)) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
        *reinterpret_cast<uint4*>(bl + stg_off_bf16(lane, 0)) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        *reinterpret_cast<uint4*>(bl + stg_off_bf16(lane, 1)) = make_uint4(lw[4], lw[5], lw[6], lw[7]);
        if (p.tma_store) {
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(tm_ohi, bh, n0 + c0, row0, p.pol_store);
            ptx::tma_store_2d(tm_olo, bl, n0 + c0, row0, p.pol_store);
            ptx::bulk_commit();
          }
        } else {
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int rr = row0 + i * 16 + hr, n = n0 + c0 + hch * 8;
            if (rr < p.M && n < p.N) {
              *reinterpret_cast<uint4*>(p.out_hi + (size_t)rr * p.ld_split + n) =
                  *reinterpret_cast<const uint4*>(bh + stg_off_bf16(i * 16 + hr, hch));
              *reinterpret_cast<uint4*>(p.out_lo + (size_t)rr * p.ld_split + n) =
                  *reinterpret_cast<const uint4*>(bl + stg_off_bf16(i * 16 + hr, hch));
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // N % 4 != 0 (hml_vec output projection, N = 263): scalar row-per-thread epilogue
    const int row = row0 + lane;
    const bool row_ok = row < p.M;
    const float* res_row = p.residual ? p.residual + (size_t)row * p.ld_res : nullptr;
    float* out_row = p.out_f32 ? p.out_f32 + (size_t)row * p.ld_out : nullptr;
    __nv_bfloat16* hi_row = p.out_hi ? p.out_hi + (size_t)row * p.ld_split : nullptr;
    __nv_bfloat16* lo_row = p.out_hi ? p.out_lo + (size_t)row * p.ld_split : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      if (n0 + c0 >= p.N) break;
      uint32_t r[16];
      __syncwarp();
      ptx::tmem_ld_32x32b_x16(acc + (uint32_t)c0, r);
      ptx::tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = n0 + c0 + j;
          if (n < p.N) {
            float w = __uint_as_float(r[j]);
            if (p.bias) w += __ldg(p.bias + n);
            if (res_row) w += res_row[n];
            if (p.gelu) w = gelu_erf(w);
            if (out_row) out_row[n] = w;
            if (hi_row) split_bf16(w, hi_row[n], lo_row[n]);
          }
        }
      }
    }
  }
}

// Epilogue variant of the 16-warp kernels (no residual): 32-column TMEM loads; the bf16 outputs leave as 32-column
// TMA boxes (64-byte rows, SWIZZLE_64B: half as many row requests for the TMA unit as 16-column boxes), hi then lo
// through the warp's single 2 KB staging tile; fp32 outputs as two 16-column boxes.  The warp's 64-column bias slice
// lives in registers (one float2 per lane, one L2 round trip per tile) and is broadcast with shuffles.
template <int PCOLS>
__device__ __forceinline__ void epilogue_slice32(const Params& p, const CUtensorMap* tm_o32, const CUtensorMap* tm_ohi,
                                                 const CUtensorMap* tm_olo, uint8_t* stg, uint32_t acc, int row0, int n0,
                                                 int lane) {
  static_assert(PCOLS == 64, "one float2 of bias per lane");
  if (n0 >= p.N) return;  // warp-uniform
  float2 b2 = make_float2(0.f, 0.f);
  if (p.bias && n0 + 2 * lane < p.N) b2 = __ldg(reinterpret_cast<const float2*>(p.bias + n0) + lane);
#pragma unroll 1
  for (int c0 = 0; c0 < PCOLS; c0 += 32) {
    if (n0 + c0 >= p.N) break;
    uint32_t r[32];
    __syncwarp();
    ptx::tmem_ld_32x32b_x32(acc + (uint32_t)c0, r);
    ptx::tmem_ld_wait();
    // + bias (lane (c0 + j) / 2 holds columns c0 + j - (j & 1) .. + 1), optional GELU
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const int src = (c0 >> 1) + (j >> 1);
      float v0 = __uint_as_float(r[j]) + __shfl_sync(0xffffffffu, b2.x, src);
      float v1 = __uint_as_float(r[j + 1]) + __shfl_sync(0xffffffffu, b2.y, src);
      if (p.gelu) {
        v0 = gelu_erf(v0);
        v1 = gelu_erf(v1);
      }
      r[j] = __float_as_uint(v0);
      r[j + 1] = __float_as_uint(v1);
    }
    if (p.out_hi) {
      uint32_t lw[16];
      if (lane == 0) ptx::bulk_wait_read<0>();  // previous store from the staging tile has been read
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t hw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = __uint_as_float(r[8 * c + 2 * k]), b = __uint_as_float(r[8 * c + 2 * k + 1]);
          hw[k] = pack_bf16x2(a, b);
          lw[4 * c + k] = pack_bf16x2(a - __uint_as_float(hw[k] << 16), b - __uint_as_float(hw[k] & 0xffff0000u));
        }
        *reinterpret_cast<uint4*>(stg + stg_off_f32(lane, c)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);  // SWIZZLE_64B
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_2d(tm_ohi, stg, n0 + c0, row0, p.pol_store);
        ptx::bulk_commit();
        ptx::bulk_wait_read<0>();
      }
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(stg + stg_off_f32(lane, c)) = make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_2d(tm_olo, stg, n0 + c0, row0, p.pol_store);
        ptx::bulk_commit();
      }
    } else {
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {  // two 16-column fp32 boxes (64-byte rows, SWIZZLE_64B)
        if (n0 + c0 + 16 * hlf >= p.N) break;
        if (lane == 0) ptx::bulk_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(stg + stg_off_f32(lane, c)) =
              make_uint4(r[16 * hlf + 4 * c], r[16 * hlf + 4 * c + 1], r[16 * hlf + 4 * c + 2], r[16 * hlf + 4 * c + 3]);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_2d(tm_o32, stg, n0 + c0 + 16 * hlf, row0);
          ptx::bulk_commit();
        }
      }
    }
  }
}

// 16-warp kernels, bf16 (hi, lo) outputs only (QKV, FFN1): the TMA unit retires about one box row per ~2.6 cycles whatever
// its length (measured in the fused GEMM+LN kernel), so 64-byte rows cost twice the row requests of 128-byte rows.  Two
// warps of the same TMEM lane quarter (column parts 2k and 2k + 1) therefore own 128 columns TOGETHER and emit them as
// two 64-column groups: in group g warp `sub` converts columns 64 g + 32 sub .. + 32 and writes the left / right 64
// bytes of every 128-byte row of ONE shared 4 KB staging tile (32 rows x 64 bf16, SWIZZLE_128B = the two warps' 2 KB
// budgets), and warp sub = 0 issues the store.  Per tile and CTA: 1024 instead of 2048 store row requests.
__device__ __forceinline__ void epilogue_pair64(const Params& p, const CUtensorMap* tm_ohi, const CUtensorMap* tm_olo,
                                                uint8_t* stg, uint32_t acc, int row0, int n0, int width, int sub,
                                                int bar_id, int lane) {
  auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); };
  // bias of this warp's 2 x 32 columns: lane holds columns n0 + 64 (lane / 16) + 32 sub + 2 (lane % 16) .. + 1
  float2 b2 = make_float2(0.f, 0.f);
  {
    const int col = n0 + 64 * (lane >> 4) + 32 * sub + 2 * (lane & 15);
    if (p.bias && col + 1 < p.N) b2 = __ldg(reinterpret_cast<const float2*>(p.bias + col));
  }
  const bool issuer = sub == 0 && lane == 0;
#pragma unroll 1
  for (int g = 0; g < 2; ++g) {
    if (64 * g >= width || n0 + 64 * g >= p.N) break;  // uniform across the pair
    uint32_t r[32];
    __syncwarp();
    ptx::tmem_ld_32x32b_x32(acc + (uint32_t)(64 * g + 32 * sub), r);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const int src = 16 * g + (j >> 1);
      ptx::f32x2 v = ptx::add2(ptx::pk2u(r[j], r[j + 1]),
                               ptx::pk2(__shfl_sync(0xffffffffu, b2.x, src), __shfl_sync(0xffffffffu, b2.y, src)));
      if (p.gelu) v = gelu_erf2(v);
      ptx::upk2u(v, r[j], r[j + 1]);
    }
    uint32_t lw[16];  // bf16 lo words, or (mixed8) 8 words of scaled-residual e4m3 bytes | 8 words of e4m3(hi) bytes
    if (issuer) ptx::bulk_wait_read<0>();  // the previous store from the shared tile has been read
    pair_sync();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t hw[4];
      if (p.m8) {
        float lo[8], hf[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = __uint_as_float(r[8 * c + 2 * k]), b = __uint_as_float(r[8 * c + 2 * k + 1]);
          hw[k] = ptx::pack_f16x2_sat(a, b);
          const ptx::f32x2 h2 = ptx::f16x2_to_f32x2(hw[k]);
          ptx::upk2(ptx::mul2(h2, ptx::splat2(0.25f)), hf[2 * k], hf[2 * k + 1]);
          ptx::upk2(ptx::mul2(ptx::sub2(ptx::pk2(a, b), h2), ptx::splat2(512.f)), lo[2 * k], lo[2 * k + 1]);
        }
        lw[2 * c] = ptx::pack_e4m3x4(lo[0], lo[1], lo[2], lo[3]);
        lw[2 * c + 1] = ptx::pack_e4m3x4(lo[4], lo[5], lo[6], lo[7]);
        lw[8 + 2 * c] = ptx::pack_e4m3x4(hf[0], hf[1], hf[2], hf[3]);
        lw[8 + 2 * c + 1] = ptx::pack_e4m3x4(hf[4], hf[5], hf[6], hf[7]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = __uint_as_float(r[8 * c + 2 * k]), b = __uint_as_float(r[8 * c + 2 * k + 1]);
          hw[k] = pack_bf16x2(a, b);
          float l0, l1;   // float(hi) is the bf16 bit pattern in the upper half of the word
          ptx::upk2(ptx::sub2(ptx::pk2(a, b), ptx::bf16x2_to_f32x2(hw[k])), l0, l1);
          lw[4 * c + k] = pack_bf16x2(l0, l1);
        }
      }
      *reinterpret_cast<uint4*>(stg + stg_off_128(lane, 4 * sub + c)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    }
    ptx::fence_proxy_async_smem();
    pair_sync();  // both halves of the hi tile are staged
    if (issuer) {
      ptx::tma_store_2d(tm_ohi, stg, n0 + 64 * g, row0, p.pol_store);
      ptx::bulk_commit();
      ptx::bulk_wait_read<0>();
    }
    pair_sync();  // hi tile read: the staging tile takes the lo halves
    if (p.m8) {
      // one byte tile of 32 rows x 128 B (SWIZZLE_128B): residual bytes of the 64 columns | e4m3(hi) bytes of the 64 columns;
      // this warp owns bytes [32 sub, 32 sub + 32) of each half
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int c = 0; c < 2; ++c)
          *reinterpret_cast<uint4*>(stg + stg_off_128(lane, 4 * t + 2 * sub + c)) =
              make_uint4(lw[8 * t + 4 * c], lw[8 * t + 4 * c + 1], lw[8 * t + 4 * c + 2], lw[8 * t + 4 * c + 3]);
      ptx::fence_proxy_async_smem();
      pair_sync();
      if (issuer) {
        ptx::tma_store_2d(tm_olo, stg, 2 * (n0 + 64 * g), row0, p.pol_store);
        ptx::bulk_commit();
      }
      continue;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<uint4*>(stg + stg_off_128(lane, 4 * sub + c)) =
          make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
    ptx::fence_proxy_async_smem();
    pair_sync();
    if (issuer) {
      ptx::tma_store_2d(tm_olo, stg, n0 + 64 * g, row0, p.pol_store);
      ptx::bulk_commit();
    }
  }
}

// Persistent kernel: grid = min(#tiles, #SMs); CTA c processes tiles c, c + grid, ... (n fastest, so
// CTAs running concurrently share A tiles through L2).  The accumulator is double buffered in TMEM
// (2 x BN columns): the epilogue of tile i overlaps the TMA/MMA main loop of tile i + 1.
template <int BN, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
               const __grid_constant__ CUtensorMap tm_o32, const __grid_constant__ CUtensorMap tm_ohi,
               const __grid_constant__ CUtensorMap tm_olo, const Params p) {
  using C = Cfg<BN, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array (NOT an integer round trip): the compiler keeps the
  // shared address space and emits LDS / STS instead of generic LD / ST for every staging access below
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* staging = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::STG_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.K / BK;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + BM - 1) / BM;
  const int num_tiles = tiles_n * tiles_m;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a_hi);
    ptx::prefetch_tmap(&tm_w_hi);
    if (SPLIT) {
      ptx::prefetch_tmap(&tm_a_lo);
      ptx::prefetch_tmap(&tm_w_lo);
    }
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tmem_full_bar[b], 1);
      ptx::mbar_init(&tmem_empty_bar[b], kEpiWarps);  // one arrival per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_base_smem, C::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  // PDL: everything above overlapped the previous kernel's tail; its results are needed from here on
  ptx::griddep_wait();
  ptx::griddep_launch();
  ptx::steplog_begin(p.steplog, p.steplog_slot);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * C::STAGE_BYTES;
          ptx::mbar_expect_tx(&full_bar[stage],
                              (SPLIT ? 2u : 1u) * (uint32_t)((p.a_rows ? p.a_rows : BM) * BK * 2 + C::W_BYTES));
          ptx::tma_load_2d(st, &tm_a_hi, &full_bar[stage], kb * BK, m0);
          ptx::tma_load_2d(st + C::A_BYTES, &tm_w_hi, &full_bar[stage], kb * BK, n0);
          if (SPLIT) {
            ptx::tma_load_2d(st + C::A_BYTES + C::W_BYTES, &tm_a_lo, &full_bar[stage], kb * BK, m0);
            ptx::tma_load_2d(st + 2 * C::A_BYTES + C::W_BYTES, &tm_w_lo, &full_bar[stage], kb * BK, n0);
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-uniform; one elected lane per
    // tcgen05 instruction, see ptx::elect_one)
    {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        ptx::mbar_wait(&tmem_empty_bar[buf], ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        ptx::tcgen05_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tcgen05_fence_after();
          const uint32_t st = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t a_hi = ptx::umma_desc_k_sw128(st);
          const uint64_t w_hi = ptx::umma_desc_k_sw128(st + C::A_BYTES);
          const uint64_t a_lo = ptx::umma_desc_k_sw128(st + C::A_BYTES + C::W_BYTES);
          const uint64_t w_lo = ptx::umma_desc_k_sw128(st + 2 * C::A_BYTES + C::W_BYTES);
          // one elected lane issues the k-block's MMAs back to back: with BN = 64 tiles (small batches) an MMA executes in
          // 32 cycles, less than electing a lane for each costs
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
              const uint64_t adv = (uint64_t)(k * UMMA_K * 2 >> 4);
              if (SPLIT) {
                // small cross terms first, the dominant hi.hi product last
                ptx::mma_f16_ss(acc, a_lo + adv, w_hi + adv, idesc, (kb | k) != 0);
                ptx::mma_f16_ss(acc, a_hi + adv, w_lo + adv, idesc, 1);
                ptx::mma_f16_ss(acc, a_hi + adv, w_hi + adv, idesc, 1);
              } else {
                ptx::mma_f16_ss(acc, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
              }
            }
          }
          if (ptx::elect_one()) ptx::tcgen05_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (ptx::elect_one()) ptx::tcgen05_commit(&tmem_full_bar[buf]);  // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which half of the tile's columns
    uint8_t* stg = staging + (warp - 2) * STG_WARP_BYTES;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      const int buf = it & 1;
      ptx::mbar_wait(&tmem_full_bar[buf], (it >> 1) & 1);
      ptx::tcgen05_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * BN + half * (BN / 2)) + ((uint32_t)(q * 32) << 16);
      epilogue_slice(p, &tm_o32, &tm_ohi, &tm_olo, stg, acc, m0 + q * 32, n0 + half * (BN / 2), BN / 2, lane);
      // release the accumulator buffer to the MMA warp
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[buf]);
    }
  }

  // ------------------------------------------------------------------ teardown
  // the staging tiles must have been READ before the CTA exits; completion of the global writes is ordered by the
  // kernel boundary (exit_wait_full = 1 waits for it here, A/B switch)
  if (p.tma_store && warp >= 2 && lane == 0) {
    if (p.exit_wait_full) ptx::bulk_wait<0>(); else ptx::bulk_wait_read<0>();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::steplog_end(p.steplog, p.steplog_slot, p.steplog_cta);
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

struct OutMaps {  // store-side tensor maps (only read when Params::tma_store != 0)
  CUtensorMap f32, hi, lo;
};

// Host launcher.  Tensor maps: A maps have box {64, 128}; W maps have box {64, BN}.  BN = 64 is the small-batch shape:
// a GEMM with at most 128 rows is bound by streaming W, so it is cut into N / 64 column tiles to use N / 64 SMs.
template <int BN, bool SPLIT>
inline cudaError_t launch(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                          const CUtensorMap& w_lo, const OutMaps& o, const Params& p, cudaStream_t stream) {
  using C = Cfg<BN, SPLIT>;
  static bool configured = false;  // per template instantiation; attribute is per-function, set once
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<BN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int64_t tiles = ceil_div(p.N, BN) * ceil_div(p.M, BM);
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  return launch_pdl(gemm_tn_kernel<BN, SPLIT>, dim3(grid), dim3(kThreads), C::SMEM_BYTES, stream, a_hi, a_lo, w_hi, w_lo,
                    o.f32, o.hi, o.lo, p);
}

// =====================================================================================================
// CTA-pair version (tcgen05 cta_group::2): two CTAs of a cluster -- two SMs of one TPC -- cooperate on a
// 256 x BN tile.  CTA r owns accumulator rows [128 r, 128 r + 128) of the tile and loads its own A rows plus HALF of
// the W tile (rows [BN/2 r, BN/2 (r+1))); the UMMA (M = 256) reads B from both CTAs' shared memory, so per SM the
// TMA traffic drops from 96 KB to 64 KB per k-block and the tensor core's shared-memory operand reads from 12 KB to
// 8 KB per instruction -- the single-CTA kernel is bound by exactly that bandwidth.  The leader CTA (rank 0)
// issues all MMAs; full barriers live in the leader, empty / accumulator-full barriers are multicast to both CTAs.
// =====================================================================================================
template <int BN, bool SPLIT, int EW>
struct Cfg2 {
  static constexpr int A_BYTES = BM * BK * 2;          // this CTA's 128 A rows
  static constexpr int W_BYTES = (BN / 2) * BK * 2;    // this CTA's half of the W tile
  static constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_BYTES + W_BYTES);
  // staging: 8 epilogue warps x 4 KB (double buffered) or 16 x 2 KB
  static constexpr int STG_BYTES = 32768;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;  // 3 for SPLIT @ BN=256 (2 stages starve the MMA)
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
  static_assert(BN == 256, "CTA-pair kernel is written for BN = 256");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// EW = 8: general epilogue (residual, fp32 and/or bf16-pair outputs), double-buffered staging.
// EW = 16 (RES = false): 4 epilogue warps per scheduler for the bias/GELU/bf16-split epilogues, which are bound by
// instruction issue and latency, not bandwidth; 576 threads -> at most 112 registers, so the residual path is compiled out.
// M8 = true (precision 'mixed8h', operand pack of gemm_ln_sm100.cuh): tm_a_hi / tm_w_hi are the fp16 halves, tm_a_lo / tm_w_lo
// the byte rows [.., 2K] (per 64 K elements: A residual bytes | A hi bytes against W hi bytes | W residual bytes), tm_ws_hi /
// tm_ws_lo the narrow W boxes of both for the tail slices.  Ring, barriers and transaction bytes are those of the bf16x3 loop
// (3 stages of 64 KB, four 128-byte-row boxes per CTA and stage: A | W | A' | W'); a stage holds TWO consecutive 128-byte
// blocks of the current phase -- first the 2K/128 correction blocks (kind::f8f6f4, 32 K elements per MMA), then the K/64
// main-term blocks (kind::f16; the very first one folds the 2^15-scaled corrections in through scale-input-d) -- i.e. 8 MMAs
// per stage instead of 12 over the same K/64 stages per tile: two thirds of the tensor time at the same bytes.
template <int BN, bool SPLIT, int EW, bool RES, bool M8 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
gemm2_tn_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                const __grid_constant__ CUtensorMap tm_o32, const __grid_constant__ CUtensorMap tm_ohi,
                const __grid_constant__ CUtensorMap tm_olo, const __grid_constant__ CUtensorMap tm_ws_hi,
                const __grid_constant__ CUtensorMap tm_ws_lo, const Params p) {
  using C = Cfg2<BN, SPLIT, EW>;
  extern __shared__ uint8_t smem_raw[];
  // 1 KB alignment by pointer arithmetic on the __shared__ array (NOT an integer round trip): the compiler keeps the
  // shared address space and emits LDS / STS instead of generic LD / ST for every staging access below
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* staging = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::STG_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2] (the leader's copy is the one in use)
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_kb = p.K / BK;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int tiles_m = (p.M + 2 * BM - 1) / (2 * BM);
  const int num_tiles = tiles_n * tiles_m;
  // Work units.  The tiles of all complete waves are 256 x BN; the tiles of the last, partial wave (R = num_tiles mod
  // num_clusters of them, which would keep R clusters busy for a whole tile time while the others idle) are cut into
  // `split` column slices of BN / split columns, R * split <= num_clusters, so the tail costs 1 / split of a tile.
  const int split = p.tail_split > 1 ? p.tail_split : 1;
  const int full_tiles = split > 1 ? (num_tiles / num_clusters) * num_clusters : num_tiles;
  const int num_units = full_tiles + (num_tiles - full_tiles) * split;
  auto decode = [&](int u, int& tile, int& ncol0, int& width) {
    if (u < full_tiles) {
      tile = u; ncol0 = 0; width = BN;
    } else {
      const int k = u - full_tiles;
      tile = full_tiles + k / split; width = BN / split; ncol0 = (k % split) * width;
    }
  };
  if (threadIdx.x == 0) REGEN_TL(0);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a_hi);
    ptx::prefetch_tmap(&tm_w_hi);
    if (SPLIT) {
      ptx::prefetch_tmap(&tm_a_lo);
      ptx::prefetch_tmap(&tm_w_lo);
    }
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);   // leader's producer arrives (expect_tx for both CTAs' bytes)
      ptx::mbar_init(&empty_bar[s], 1);  // one multicast tcgen05.commit
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tmem_full_bar[b], 1);
      ptx::mbar_init(&tmem_empty_bar[b], 2 * EW);  // epilogue warps of both CTAs
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(tmem_base_smem, C::TMEM_COLS);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();      // CTA-level ordering of the TMEM base address (written by tcgen05.alloc) before it is read below;
                        // the cluster barrier alone orders it too, but compute-sanitizer racecheck does not model that
  ptx::cluster_sync();  // barriers of both CTAs initialised before any remote arrive / multicast commit
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  ptx::griddep_wait();    // PDL: the setup above overlapped the previous kernel's tail
  ptx::griddep_launch();  // the next kernel's CTAs may take over each SM as soon as this CTA exits
  ptx::steplog_begin(p.steplog, p.steplog_slot);
  if (threadIdx.x == 0) REGEN_TL(1);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = cluster_id; u < num_units; u += num_clusters) {
        int tile, ncol0, width;
        decode(u, tile, ncol0, width);
        const int m0 = (tile / tiles_n) * (2 * BM) + (int)rank * BM;
        // this CTA's half of the W slice.  Whole tiles load the BN / 2-row box; column slices of the tail load the narrow
        // slice box (slice_w_rows rows: a slice is bound by its operand loads, not by its MMAs), of which the first
        // width / 2 rows are used
        const int nw = (tile % tiles_n) * BN + ncol0 + (int)rank * (width / 2);
        const bool slice = width < BN;
        const CUtensorMap* wmap_hi = slice ? &tm_ws_hi : &tm_w_hi;
        const CUtensorMap* wmap_lo = slice ? &tm_ws_lo : &tm_w_lo;
        const uint32_t stage_tx = (SPLIT ? 2u : 1u) * (uint32_t)(C::A_BYTES + (slice ? p.slice_w_rows * BK * 2 : C::W_BYTES));
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * C::STAGE_BYTES;
          if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * stage_tx);
          if constexpr (M8) {
            // stage kb carries blocks 2 kb and 2 kb + 1 of the sequence [2K/128 correction blocks | K/64 main-term blocks]
            // (num_kb = K/64 is even: K % 128 == 0); every box row is 128 bytes: 128 e4m3 bytes or 64 fp16
            const int n1 = 2 * (p.K / 128);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int blk = 2 * kb + h;
              const bool corr = blk < n1;
              const int c0 = corr ? blk * 128 : (blk - n1) * BK;
              ptx::tma_load_2d_2sm(st + h * (C::A_BYTES + C::W_BYTES), corr ? &tm_a_lo : &tm_a_hi, &full_bar[stage], c0, m0);
              ptx::tma_load_2d_2sm(st + h * (C::A_BYTES + C::W_BYTES) + C::A_BYTES, corr ? wmap_lo : wmap_hi, &full_bar[stage], c0,
                                   nw, p.pol_w);
            }
          } else {
          ptx::tma_load_2d_2sm(st, &tm_a_hi, &full_bar[stage], kb * BK, m0);
          ptx::tma_load_2d_2sm(st + C::A_BYTES, wmap_hi, &full_bar[stage], kb * BK, nw, p.pol_w);
          if (SPLIT) {
            ptx::tma_load_2d_2sm(st + C::A_BYTES + C::W_BYTES, &tm_a_lo, &full_bar[stage], kb * BK, m0);
            ptx::tma_load_2d_2sm(st + 2 * C::A_BYTES + C::W_BYTES, wmap_lo, &full_bar[stage], kb * BK, nw, p.pol_w);
          }
          }
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // The whole warp runs the loop (warp-uniform control flow and operands); one elected lane issues each tcgen05
    // instruction (ptx::elect_one: a single-lane loop costs more cycles per MMA in issue than the MMA takes to execute).
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = cluster_id; u < num_units; u += num_clusters, ++it) {
        int tile, ncol0, width;
        decode(u, tile, ncol0, width);
        const uint32_t idesc = ptx::umma_idesc_bf16_f32(2 * BM, width);
        const int buf = it & 1;
        ptx::mbar_wait(&tmem_empty_bar[buf], ((it >> 1) & 1) ^ 1);  // both CTAs drained this accumulator
        ptx::tcgen05_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          if (kb == 0 && it < 16 && lane == 0) REGEN_TL(8 + 2 * it);
          if (it < 3 && kb < 8 && lane == 0) REGEN_TL(104 + 8 * it + kb);  // bring-up: k-block arrival times of the first tiles
          ptx::tcgen05_fence_after();
          const uint32_t st = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
          const uint64_t a_hi = ptx::umma_desc_k_sw128(st);
          const uint64_t w_hi = ptx::umma_desc_k_sw128(st + C::A_BYTES);
          const uint64_t a_lo = ptx::umma_desc_k_sw128(st + C::A_BYTES + C::W_BYTES);
          const uint64_t w_lo = ptx::umma_desc_k_sw128(st + 2 * C::A_BYTES + C::W_BYTES);
          if constexpr (M8) {
            // (a_hi, w_hi) = first block of the stage, (a_lo, w_lo) = second; formats 0 = e4m3 x e4m3 / fp16 x fp16
            const uint32_t idesc0 = ptx::umma_idesc_fmt0_f32(2 * BM, width);
            const int ncs = p.K / 128;   // correction stages; stage ncs starts the main term (straight-line code per stage kind)
            if (kb < ncs) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint64_t adv = (uint64_t)(2 * (i & 3));   // 32 bytes per k-step
                if (ptx::elect_one())
                  ptx::mma_f8_ss_2sm(acc, ((i >> 2) ? a_lo : a_hi) + adv, ((i >> 2) ? w_lo : w_hi) + adv, idesc0, (kb | i) != 0);
              }
            } else {
              if (kb == ncs) {
                if (ptx::elect_one()) ptx::mma_f16_ss_2sm_scale15(acc, a_hi, w_hi, idesc0);   // D = A.B + D * 2^-15
              } else {
                if (ptx::elect_one()) ptx::mma_f16_ss_2sm(acc, a_hi, w_hi, idesc0, 1);
              }
#pragma unroll
              for (int i = 1; i < 8; ++i) {
                const uint64_t adv = (uint64_t)(2 * (i & 3));
                if (ptx::elect_one())
                  ptx::mma_f16_ss_2sm(acc, ((i >> 2) ? a_lo : a_hi) + adv, ((i >> 2) ? w_lo : w_hi) + adv, idesc0, 1);
              }
            }
          } else {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)(k * UMMA_K * 2 >> 4);
            if (SPLIT) {
              if (ptx::elect_one()) ptx::mma_f16_ss_2sm(acc, a_lo + adv, w_hi + adv, idesc, (kb | k) != 0);
              if (ptx::elect_one()) ptx::mma_f16_ss_2sm(acc, a_hi + adv, w_lo + adv, idesc, 1);
              if (ptx::elect_one()) ptx::mma_f16_ss_2sm(acc, a_hi + adv, w_hi + adv, idesc, 1);
            } else {
              if (ptx::elect_one()) ptx::mma_f16_ss_2sm(acc, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
            }
          }
          }
          if (ptx::elect_one()) ptx::tcgen05_commit_2sm(&empty_bar[stage]);  // frees this stage in BOTH CTAs
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (ptx::elect_one()) ptx::tcgen05_commit_2sm(&tmem_full_bar[buf]);  // accumulator complete, signalled to both epilogues
        if (it < 16 && lane == 0) REGEN_TL(9 + 2 * it);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..EW+1 of both CTAs)
    constexpr int PARTS = EW / 4;          // warps per TMEM lane quarter; each takes BN / PARTS columns
    constexpr int PCOLS = BN / PARTS;
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    uint8_t* stg = staging + (warp - 2) * (C::STG_BYTES / EW);
    int it = 0;
    for (int u = cluster_id; u < num_units; u += num_clusters, ++it) {
      int tile, ncol0, width;
      decode(u, tile, ncol0, width);
      const int m0 = (tile / tiles_n) * (2 * BM) + (int)rank * BM, n0 = (tile % tiles_n) * BN + ncol0;
      const int buf = it & 1;
      ptx::mbar_wait(&tmem_full_bar[buf], (it >> 1) & 1);
      if (warp == 2 && lane == 0 && it < 16) REGEN_TL(40 + 2 * it);
      ptx::tcgen05_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)(buf * BN + part * PCOLS) + ((uint32_t)(q * 32) << 16);
      // a column slice narrower than the tile only has accumulator columns [0, width): the other parts just release
      if (EW == 16 && p.pair64) {
        // warp pair (parts 2k, 2k + 1) of lane quarter q: 128 columns through one shared 4 KB staging tile
        const int k = part >> 1, pidx = k * 4 + ((warp - 2) & 3);
        if (k * 128 < width)
          epilogue_pair64(p, &tm_ohi, &tm_olo, staging + pidx * 4096,
                          tmem_base + (uint32_t)(buf * BN + k * 128) + ((uint32_t)(q * 32) << 16), m0 + q * 32,
                          n0 + k * 128, width - k * 128, part & 1, 1 + pidx, lane);
      } else if (part * PCOLS < width) {
        if constexpr (EW == 16)
          epilogue_slice32<PCOLS>(p, &tm_o32, &tm_ohi, &tm_olo, stg, acc, m0 + q * 32, n0 + part * PCOLS, lane);
        else
          epilogue_slice<RES, true>(p, &tm_o32, &tm_ohi, &tm_olo, stg, acc, m0 + q * 32, n0 + part * PCOLS,
                                    width - part * PCOLS < PCOLS ? width - part * PCOLS : PCOLS, lane,
                                    warp == 2 && it == 0);
      }
      if (warp == 2 && lane == 0 && it < 16) REGEN_TL(41 + 2 * it);
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_remote(&tmem_empty_bar[buf], 0);  // leader's barrier
    }
  }

  // ------------------------------------------------------------------ teardown
  // the staging tiles must have been READ before the CTA exits; completion of the global writes is ordered by the
  // kernel boundary (exit_wait_full = 1 waits for it here, A/B switch)
  if (p.tma_store && warp >= 2 && lane == 0) {
    if (p.exit_wait_full) ptx::bulk_wait<0>(); else ptx::bulk_wait_read<0>();
  }
  ptx::tcgen05_fence_before();
  ptx::cluster_sync();  // no CTA may exit (or free TMEM) while its peer can still touch its smem / barriers
  ptx::steplog_end(p.steplog, p.steplog_slot, p.steplog_cta);
  if (threadIdx.x == 0) REGEN_TL(2);
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
  }
}

// W tensor maps for the pair kernel need box {64, BN/2}.
// narrow W boxes for the column slices of the tail (null -> the slices load the full BN / 2-row box)
struct SliceMaps {
  const CUtensorMap *hi32 = nullptr, *lo32 = nullptr;  // box {64, 32}: quarter slices (tail_split = 4)
  const CUtensorMap *hi64 = nullptr, *lo64 = nullptr;  // box {64, 64}: half slices (tail_split = 2)
};

template <int BN, bool SPLIT, int EW, bool RES, bool M8 = false>
inline cudaError_t launch2_impl(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                                const CUtensorMap& w_lo, const OutMaps& o, const Params& p, cudaStream_t stream,
                                const SliceMaps& sm) {
  using C = Cfg2<BN, SPLIT, EW>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_tn_kernel<BN, SPLIT, EW, RES, M8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int64_t tiles = ceil_div(p.N, BN) * ceil_div(p.M, 2 * BM);
  int clusters = (int)(tiles < kNumSMs / 2 ? tiles : kNumSMs / 2);
  static int cap = -1, no_split = -1;  // REGEN_DEBUG_GEMM_MAXCLUSTERS / REGEN_DEBUG_NO_TAIL_SPLIT: bring-up A/B switches
  if (cap < 0) {
    const char* e = getenv("REGEN_DEBUG_GEMM_MAXCLUSTERS");
    cap = e ? atoi(e) : 0;
    const char* e2 = getenv("REGEN_DEBUG_NO_TAIL_SPLIT");
    no_split = (e2 && e2[0] == '1') ? 1 : 0;
  }
  if (cap > 0 && clusters > cap) clusters = cap;
  // tail of the persistent schedule: R tiles left for the last wave -> cut them into 2 or 4 column slices if that still
  // fits one wave (FFN1 at config 2: 240 tiles on 74 clusters, R = 18 -> 72 slices of 64 columns; a lone small GEMM
  // with fewer tiles than clusters is spread over up to 4x as many SMs the same way)
  Params q = p;
  q.tail_split = 1;
  if (!no_split && cap <= 0) {
    const int maxc = kNumSMs / 2;
    const int rem = tiles >= maxc ? (int)(tiles % maxc) : (int)tiles;
    if (rem > 0 && 4 * rem <= maxc) q.tail_split = 4;
    else if (rem > 0 && 2 * rem <= maxc) q.tail_split = 2;
    if (tiles < maxc) clusters = (int)tiles * q.tail_split;  // every slice gets its own cluster
  }
  const CUtensorMap *ws_hi = &w_hi, *ws_lo = &w_lo;
  q.slice_w_rows = BN / 2;
  if (q.tail_split == 4 && sm.hi32 && sm.lo32) {
    ws_hi = sm.hi32; ws_lo = sm.lo32; q.slice_w_rows = 32;
  } else if (q.tail_split == 2 && sm.hi64 && sm.lo64) {
    ws_hi = sm.hi64; ws_lo = sm.lo64; q.slice_w_rows = 64;
  }
  return launch_pdl(gemm2_tn_kernel<BN, SPLIT, EW, RES, M8>, dim3(2 * clusters), dim3(64 + 32 * EW), C::SMEM_BYTES, stream, a_hi,
                    a_lo, w_hi, w_lo, o.f32, o.hi, o.lo, *ws_hi, *ws_lo, q);
}

// mixed8 operands (K % 128 == 0, no residual, TMA-store epilogue): a16 / a8 / w16 / w8 maps; sm: narrow W boxes of the tail
// slices (hi32 / hi64 = 16-bit halves, lo32 / lo64 = byte rows), else the slices load the full-height boxes
template <int BN>
inline cudaError_t launch2_m8(const CUtensorMap& a16, const CUtensorMap& a8, const CUtensorMap& w16, const CUtensorMap& w8,
                              const OutMaps& o, const Params& p, cudaStream_t stream, const SliceMaps& sm) {
  if (p.residual || !p.tma_store || (p.out_f32 && p.out_hi) || (p.K & 127)) return cudaErrorInvalidValue;
  return launch2_impl<BN, true, 16, false, true>(a16, a8, w16, w8, o, p, stream, sm);
}

template <int BN, bool SPLIT>
inline cudaError_t launch2(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                           const CUtensorMap& w_lo, const OutMaps& o, const Params& p, cudaStream_t stream,
                           const SliceMaps& sm = SliceMaps()) {
  // 16 epilogue warps whenever no residual is added and the outputs leave through TMA stores
  if (!p.residual && p.tma_store && !(p.out_f32 && p.out_hi))
    return launch2_impl<BN, SPLIT, 16, false>(a_hi, a_lo, w_hi, w_lo, o, p, stream, sm);
  return launch2_impl<BN, SPLIT, 8, true>(a_hi, a_lo, w_hi, w_lo, o, p, stream, sm);
}

}  // namespace gemm
}  // namespace regen
