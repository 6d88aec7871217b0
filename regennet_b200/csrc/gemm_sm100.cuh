// Warp-specialised tcgen05 GEMM for sm_100a:   C[M,N] = epilogue( A[M,K] . W[N,K]^T )
//
//   * operands are bf16, K-major, staged global -> shared by TMA (128-byte swizzle) into a
//     multi-stage mbarrier ring; one elected thread issues tcgen05.mma (UMMA 128 x BN x 16,
//     cta_group::1) accumulating fp32 in tensor memory;
//   * SPLIT mode ("bf16x3"): every fp32 operand is carried as a (hi, lo) bf16 pair and each k-step
//     issues three MMAs  A_hi.W_hi + A_hi.W_lo + A_lo.W_hi  into the same TMEM accumulator, which
//     restores ~16 significand bits (max abs error ~2e-5 on the CMDM forward instead of 1e-2);
//   * epilogue: 4 warps read the accumulator with tcgen05.ld (32x32b: one thread = one output
//     row), add bias / residual, optionally apply exact-erf GELU, and store fp32 and/or a bf16
//     (hi, lo) pair for the next GEMM's A operand.
//
// Roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
// One CTA per 128 x BN output tile.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace regen {
namespace gemm {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;

template <int BN, bool SPLIT>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int W_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_BYTES + W_BYTES);
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;  // 2 for SPLIT @ BN=256, 4 for plain bf16
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;                         // power of two for BN in {32..256}
  static_assert(STAGES >= 2, "need at least a double buffer");
  static_assert((BN & (BN - 1)) == 0 && BN >= 32 && BN <= 256, "BN must be a power of two in [32,256]");
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
};

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
               const Params p) {
  using C = Cfg<BN, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int num_kb = p.K / BK;

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
    ptx::mbar_init(tmem_full_bar, 1);
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

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * C::STAGE_BYTES;
        ptx::mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
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
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tcgen05_fence_after();
        const uint32_t st = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
        const uint64_t a_hi = ptx::umma_desc_k_sw128(st);
        const uint64_t w_hi = ptx::umma_desc_k_sw128(st + C::A_BYTES);
        const uint64_t a_lo = ptx::umma_desc_k_sw128(st + C::A_BYTES + C::W_BYTES);
        const uint64_t w_lo = ptx::umma_desc_k_sw128(st + 2 * C::A_BYTES + C::W_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
          const uint64_t adv = (uint64_t)(k * UMMA_K * 2 >> 4);
          if (SPLIT) {
            // small cross terms first, the dominant hi.hi product last
            ptx::mma_f16_ss(tmem_base, a_lo + adv, w_hi + adv, idesc, (kb | k) != 0);
            ptx::mma_f16_ss(tmem_base, a_hi + adv, w_lo + adv, idesc, 1);
            ptx::mma_f16_ss(tmem_base, a_hi + adv, w_hi + adv, idesc, 1);
          } else {
            ptx::mma_f16_ss(tmem_base, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
          }
        }
        ptx::tcgen05_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs retire
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      ptx::tcgen05_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tcgen05_fence_after();
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.M;
    const bool vec_ok = (p.N & 3) == 0;
    const float* res_row = p.residual ? p.residual + (size_t)row * p.ld_res : nullptr;
    float* out_row = p.out_f32 ? p.out_f32 + (size_t)row * p.ld_out : nullptr;
    __nv_bfloat16* hi_row = p.out_hi ? p.out_hi + (size_t)row * p.ld_split : nullptr;
    __nv_bfloat16* lo_row = p.out_hi ? p.out_lo + (size_t)row * p.ld_split : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: the warp must be converged
      ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      ptx::tmem_ld_wait();
      if (row_ok && n0 + c0 < p.N) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int n = n0 + c0 + j;
        if (n >= p.N) break;
        float v[4] = {__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                      __uint_as_float(r[j + 3])};
        if (vec_ok) {
          if (p.bias) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            v[0] += b4.x; v[1] += b4.y; v[2] += b4.z; v[3] += b4.w;
          }
          if (res_row) {
            const float4 r4 = *reinterpret_cast<const float4*>(res_row + n);
            v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
          }
          if (p.gelu) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = gelu_erf(v[e]);
          }
          if (out_row) *reinterpret_cast<float4*>(out_row + n) = make_float4(v[0], v[1], v[2], v[3]);
          if (hi_row) {
            __nv_bfloat16 h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_bf16(v[e], h[e], l[e]);
            *reinterpret_cast<uint2*>(hi_row + n) = *reinterpret_cast<uint2*>(h);
            *reinterpret_cast<uint2*>(lo_row + n) = *reinterpret_cast<uint2*>(l);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (n + e >= p.N) break;
            float w = v[e];
            if (p.bias) w += __ldg(p.bias + n + e);
            if (res_row) w += res_row[n + e];
            if (p.gelu) w = gelu_erf(w);
            if (out_row) out_row[n + e] = w;
            if (hi_row) split_bf16(w, hi_row[n + e], lo_row[n + e]);
          }
        }
      }
      }
    }
  }

  // ------------------------------------------------------------------ teardown
  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// Host launcher.  Tensor maps: A maps have box {64, 128}; W maps have box {64, BN}.
template <int BN, bool SPLIT>
inline cudaError_t launch(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                          const CUtensorMap& w_lo, const Params& p, cudaStream_t stream) {
  using C = Cfg<BN, SPLIT>;
  static bool configured = false;  // per template instantiation; attribute is per-function, set once
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<BN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(p.N, BN), (unsigned)ceil_div(p.M, BM));
  gemm_tn_kernel<BN, SPLIT><<<grid, kThreads, C::SMEM_BYTES, stream>>>(a_hi, a_lo, w_hi, w_lo, p);
  return cudaGetLastError();
}

}  // namespace gemm
}  // namespace regen
